// Minimal stand-ins for the pieces of Caffe v1 (LMB fork) the DataGeneration layer touches, so the
// layer boundary can be compiled, driven and tested without Caffe (which the reference does not
// vendor: SURVEY 2 #18). Names, signatures and semantics follow Caffe's so that the layer source
// reads like the reference's (/root/reference/src/caffe/layers/data_generation_layer.cpp); in a real
// Caffe build this header is replaced by <caffe/blob.hpp>, <caffe/layer.hpp> and caffe.pb.h
// (INTEGRATION.md).
#pragma once
#include <cstddef>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace caffe {

// Blob with lazily synchronised host / device copies (caffe::SyncedMemory semantics).
template <typename Dtype>
class Blob {
 public:
  Blob() {}
  ~Blob();
  Blob(const Blob&) = delete;
  Blob& operator=(const Blob&) = delete;
  void Reshape(const std::vector<int>& shape);
  void ReshapeLike(const Blob& other) { Reshape(other.shape_); }
  const std::vector<int>& shape() const { return shape_; }
  int shape(int i) const { return shape_[i]; }
  int num_axes() const { return (int)shape_.size(); }
  int count() const { return count_; }
  int offset(int n, int c = 0, int h = 0, int w = 0) const {
    return ((n * shape_[1] + c) * shape_[2] + h) * shape_[3] + w;
  }
  const Dtype* cpu_data();
  Dtype* mutable_cpu_data();
  const Dtype* gpu_data();
  Dtype* mutable_gpu_data();

 private:
  enum Head { UNINITIALIZED, HEAD_AT_CPU, HEAD_AT_GPU, SYNCED };
  void to_cpu();
  void to_gpu();
  std::vector<int> shape_;
  int count_ = 0;
  size_t capacity_ = 0;
  Dtype* cpu_ = nullptr;
  Dtype* gpu_ = nullptr;
  Head head_ = UNINITIALIZED;
};

// caffe.proto messages, reduced to the fields the layer reads
// (/root/reference/src/caffe/proto/caffe.proto:1-13 and data_param of the LMB fork).
struct DataParameter {
  int batch_size_ = 1, prefetch_ = 4, block_size_ = 0;
  bool verbose_ = false;
  std::vector<std::string> sample_;
  int batch_size() const { return batch_size_; }
  int prefetch() const { return prefetch_; }
  int block_size() const { return block_size_; }
  bool verbose() const { return verbose_; }
  const std::vector<std::string>& sample() const { return sample_; }
};
struct DataGenerationParameter {
  int mode_ = 1;                        // required int32 mode = 9003 [default = 1]
  std::vector<std::string> texture_dbases_;
  int first_level_threads_ = 16;        // accepted, unused: the CPU worker pools are gone
  int second_level_threads_ = 1;
  bool use_antialiasing_ = true;
  // extensions of this implementation (optional, default off => the reference's behaviour):
  bool device_params_ = false;          // draw the scene parameters on the GPU (Philox production mode)
  unsigned long long seed_ = 0;         // seed of the device-side stream
  bool device_params() const { return device_params_; }
  unsigned long long seed() const { return seed_; }
  int mode() const { return mode_; }
  const std::string& texture_dbases(int i) const { return texture_dbases_.at(i); }
  int texture_dbases_size() const { return (int)texture_dbases_.size(); }
  int first_level_threads() const { return first_level_threads_; }
  int second_level_threads() const { return second_level_threads_; }
  bool use_antialiasing() const { return use_antialiasing_; }
};
struct LayerParameter {
  std::string name_, type_;
  std::vector<std::string> top_, bottom_;
  DataParameter data_param_;
  DataGenerationParameter data_generation_param_;
  const std::string& name() const { return name_; }
  const std::string& type() const { return type_; }
  int top_size() const { return (int)top_.size(); }
  int bottom_size() const { return (int)bottom_.size(); }
  const DataParameter& data_param() const { return data_param_; }
  const DataGenerationParameter& data_generation_param() const { return data_generation_param_; }
};

// Parses the text-format `layer { ... }` block of a prototxt (protoc is not available offline).
// Accepts /root/reference/example-prototxt/train.prototxt verbatim: comments, quoted strings,
// nested messages; unknown fields are an error, like protobuf's TextFormat.
LayerParameter ParseLayerPrototxt(const std::string& text);

template <typename Dtype>
class Layer {
 public:
  explicit Layer(const LayerParameter& param) : layer_param_(param) {}
  virtual ~Layer() {}
  void SetUp(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top) {
    LayerSetUp(bottom, top);
    Reshape(bottom, top);
  }
  virtual void LayerSetUp(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top) {}
  virtual void Reshape(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top) = 0;
  virtual const char* type() const { return ""; }
  virtual int ExactNumBottomBlobs() const { return -1; }
  virtual int MinTopBlobs() const { return -1; }
  virtual bool ShareInParallel() const { return false; }
  virtual void Forward_cpu(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top) = 0;
  virtual void Forward_gpu(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top) {
    Forward_cpu(bottom, top);
  }
  const LayerParameter& layer_param() const { return layer_param_; }

 protected:
  LayerParameter layer_param_;
};

// caffe/layer_factory.hpp: layers register a creator under their type string; Net::Init builds every layer of a
// prototxt with LayerRegistry<Dtype>::CreateLayer(param). The reference registers itself with
// REGISTER_LAYER_CLASS(DataGeneration) (src/caffe/layers/data_generation_layer.cpp:298-299).
template <typename Dtype>
class LayerRegistry {
 public:
  typedef std::shared_ptr<Layer<Dtype> > (*Creator)(const LayerParameter&);
  typedef std::map<std::string, Creator> CreatorRegistry;
  static CreatorRegistry& Registry() {
    static CreatorRegistry* g_registry_ = new CreatorRegistry();
    return *g_registry_;
  }
  static void AddCreator(const std::string& type, Creator creator) {
    CreatorRegistry& registry = Registry();
    if (registry.count(type)) throw std::runtime_error("Layer type " + type + " already registered.");
    registry[type] = creator;
  }
  static std::shared_ptr<Layer<Dtype> > CreateLayer(const LayerParameter& param) {
    const std::string& type = param.type();
    CreatorRegistry& registry = Registry();
    if (registry.count(type) != 1) throw std::runtime_error("Unknown layer type: " + type + " (known types: " + LayerTypeListString() + ")");
    return registry[type](param);
  }
  static std::vector<std::string> LayerTypeList() {
    std::vector<std::string> layer_types;
    for (typename CreatorRegistry::iterator iter = Registry().begin(); iter != Registry().end(); ++iter) layer_types.push_back(iter->first);
    return layer_types;
  }

 private:
  LayerRegistry() {}
  static std::string LayerTypeListString() {
    std::string s;
    for (const std::string& t : LayerTypeList()) s += (s.empty() ? "" : ", ") + t;
    return s;
  }
};
template <typename Dtype>
class LayerRegisterer {
 public:
  LayerRegisterer(const std::string& type, std::shared_ptr<Layer<Dtype> > (*creator)(const LayerParameter&)) {
    LayerRegistry<Dtype>::AddCreator(type, creator);
  }
};
// (the shim instantiates layers for float only; Caffe's macro registers float and double)
#define REGISTER_LAYER_CREATOR(type, creator) static ::caffe::LayerRegisterer<float> g_creator_f_##type(#type, creator<float>)
#define REGISTER_LAYER_CLASS(type)                                                                   \
  template <typename Dtype>                                                                          \
  std::shared_ptr<Layer<Dtype> > Creator_##type##Layer(const LayerParameter& param) {               \
    return std::shared_ptr<Layer<Dtype> >(new type##Layer<Dtype>(param));                            \
  }                                                                                                  \
  REGISTER_LAYER_CREATOR(type, Creator_##type##Layer)

}  // namespace caffe
