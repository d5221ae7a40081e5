// TEST INFRASTRUCTURE (oracle/shim): the handful of protobuf accessors the reference reads from
// caffe::LayerParameter (/root/reference/src/caffe/proto/caffe.proto:2-12 for data_generation_param;
// data_param{batch_size, prefetch, verbose, block_size, sample} are fields of the LMB Caffe fork, used at
// /root/reference/src/caffe/layers/data_generation_layer.cpp:44-46,109-113,185). Plain structs instead of
// generated protobuf code: Caffe and protoc are not installed here.
#ifndef OFDG_ORACLE_CAFFE_PB_SHIM_H_
#define OFDG_ORACLE_CAFFE_PB_SHIM_H_
#include <string>
#include <vector>

namespace caffe {

class DataGenerationParameter {
 public:
  int mode() const { return mode_; }
  const std::string& texture_dbases(int i) const { return texture_dbases_.at(i); }
  int texture_dbases_size() const { return (int)texture_dbases_.size(); }
  int first_level_threads() const { return first_level_threads_; }
  int second_level_threads() const { return second_level_threads_; }
  bool use_antialiasing() const { return use_antialiasing_; }
  void set_mode(int v) { mode_ = v; }
  void add_texture_dbases(const std::string& s) { texture_dbases_.push_back(s); }
  void set_first_level_threads(int v) { first_level_threads_ = v; }
  void set_second_level_threads(int v) { second_level_threads_ = v; }
  void set_use_antialiasing(bool v) { use_antialiasing_ = v; }
 private:
  int mode_ = 1;                       // [default = 1]
  std::vector<std::string> texture_dbases_;
  int first_level_threads_ = 16;       // [default = 16]
  int second_level_threads_ = 1;       // [default = 1]
  bool use_antialiasing_ = true;       // [default = true]
};

class DataParameter {
 public:
  int batch_size() const { return batch_size_; }
  int prefetch() const { return prefetch_; }
  bool verbose() const { return verbose_; }
  int block_size() const { return block_size_; }
  const std::vector<std::string>& sample() const { return sample_; }
  void set_batch_size(int v) { batch_size_ = v; }
  void set_prefetch(int v) { prefetch_ = v; }
 private:
  int batch_size_ = 1, prefetch_ = 4, block_size_ = 0;
  bool verbose_ = false;
  std::vector<std::string> sample_;
};

class LayerParameter {
 public:
  const DataGenerationParameter& data_generation_param() const { return dgp_; }
  DataGenerationParameter* mutable_data_generation_param() { return &dgp_; }
  const DataParameter& data_param() const { return dp_; }
  DataParameter* mutable_data_param() { return &dp_; }
  int top_size() const { return (int)top_.size(); }
  void add_top(const std::string& s) { top_.push_back(s); }
  const std::string& type() const { return type_; }
  void set_type(const std::string& s) { type_ = s; }
 private:
  DataGenerationParameter dgp_;
  DataParameter dp_;
  std::vector<std::string> top_;
  std::string type_ = "DataGeneration";
};

}  // namespace caffe
#endif
