// DataGenerationLayer: the reference's Caffe-v1 layer surface (type "DataGeneration", 0 bottoms,
// 3 tops) over the B200 generator. Mirrors, member for member where it still makes sense,
//   /root/reference/include/caffe/layers/data_generation_layer.hpp:37-89
//   /root/reference/src/caffe/layers/data_generation_layer.cpp
// What changes underneath: the prefetch thread only draws scene parameters and flattens geometry
// (ofdg_prepare); Forward_gpu launches the sm_100a kernels straight into the top blobs' device
// memory (ofdg_render_prepared), so the reference's three host copies per sample and its wasted
// async H2D push (data_generation_layer.cpp:155-161, 242-250, 273-278) disappear.
#pragma once
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "caffe_shim.hpp"
#include "ofdg/ofdg.h"

namespace caffe {

template <typename Dtype>
class DataGenerationLayer : public Layer<Dtype> {
 public:
  explicit DataGenerationLayer(const LayerParameter& param);
  virtual ~DataGenerationLayer();
  virtual void LayerSetUp(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top);
  virtual void Reshape(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top) {}

  virtual inline bool ShareInParallel() const { return false; }
  virtual inline const char* type() const { return "DataGeneration"; }
  virtual inline int ExactNumBottomBlobs() const { return 0; }
  virtual inline int MinTopBlobs() const { return 1; }

  virtual void Forward_cpu(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top);
  virtual void Forward_gpu(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top);
  virtual void Backward_cpu(const std::vector<Blob<Dtype>*>&, const std::vector<bool>&, const std::vector<Blob<Dtype>*>&) {}
  virtual void Backward_gpu(const std::vector<Blob<Dtype>*>&, const std::vector<bool>&, const std::vector<Blob<Dtype>*>&) {}

  // Multi-GPU: every solver replica owns its own layer (ShareInParallel() == false); the reference
  // seeds all of them identically. Rank r offsets the 45 engine seeds by 45*r.
  // Call before constructing the layer (Caffe::solver_rank() in a real build).
  static void set_solver_rank(int rank) { solver_rank_ = rank; }
  // Checkpoint/resume: tasks commissioned so far; fast-forward a fresh layer to that point.
  uint64_t tasks_commissioned() const;

 protected:
  virtual void InternalThreadEntry();                 // prefetch producer
  virtual void load_batch(ofdg_prepared** out);       // draw + flatten + upload one batch
  void StartInternalThread();
  void StopInternalThread();
  ofdg_prepared* PopPrefetched();                      // blocks until the producer has a batch
  void BindExtraTops(const std::vector<Blob<Dtype>*>& top);  // top[3..6]: backward flow, occlusion, index images

  static int solver_rank_;
  static constexpr int kFieldPool = 40;   // mode 9: (flow, iflow) crops generated at set-up
  int device_ = 0;
  ofdg_generator* generator_ = nullptr;   // DataGenerator::DataGenerator data_generator_
  ofdg_params* params_ = nullptr;         // DataGenerator::ObjectParametersGenerator obj_params_generator_
  ofdg_tasks* tasks_ = nullptr;
  // prefetch_free_/prefetch_full_ of the reference collapse into one bounded queue of prepared batches
  std::deque<ofdg_prepared*> prefetch_full_;
  size_t prefetch_depth_ = 1;
  std::mutex mutex_, generator_mutex_;
  std::condition_variable cv_full_, cv_free_;
  std::thread thread_;
  bool must_stop_ = false;
  unsigned long long device_batches_ = 0;  // batches produced by the device-side stream (its sample counter)
  std::string producer_error_;
  ofdg_extra_tops extra_{};               // what the generator currently writes besides the three blobs
};

}  // namespace caffe
