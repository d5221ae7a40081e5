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

#include <cuda_runtime.h>

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
  // Prefetch-side timing since the last call: {ms spent drawing (incl. waiting for the stream), ms in ofdg_prepare, batches}.
  void producer_stats(double* out3);

 protected:
  struct Prefetched {
    ofdg_prepared* scene = nullptr;
    uint64_t gen_lo = 0, gen_hi = 0;  // mode 9: field generations (40 crops each) the batch's objects picked from
    bool uses_fields = false;
  };
  virtual void InternalThreadEntry(int producer);     // prefetch producer
  virtual void load_batch(Prefetched* out, int producer, uint64_t* ticket);  // draw + flatten + upload one batch
  void StartInternalThread();
  void StopInternalThread();
  Prefetched PopPrefetched();                          // blocks until the producer has a batch
  void BindExtraTops(const std::vector<Blob<Dtype>*>& top);  // top[3..6]: backward flow, occlusion, index images
  void BeginForward();                                 // layer stream <- default stream (the consumer's reads of the old blobs)
  void EndForward();                                   // default stream <- layer stream (the new blobs)

  static int solver_rank_;
  static constexpr int kFieldPool = 40;   // mode 9: (flow, iflow) crops per generation = per 3*max(W,H) canvas (WarpFields.cpp:619-637)
  int device_ = 0;
  ofdg_generator* generator_ = nullptr;   // DataGenerator::DataGenerator data_generator_
  ofdg_params* params_ = nullptr;         // DataGenerator::ObjectParametersGenerator obj_params_generator_
  static constexpr int kProducers = 2;    // prefetch threads: one draws batch k+1 (the RNG stream is sequential) while the other flattens batch k
  ofdg_tasks* tasks_[kProducers] = {nullptr, nullptr};
  // prefetch_free_/prefetch_full_ of the reference collapse into one bounded queue of prepared batches
  std::deque<Prefetched> prefetch_full_;
  // mode 9: the field pool is a ring of `field_ring_` generations of kFieldPool crops. The producer regenerates a
  // generation's slots right before the first batch that picks from it, after every batch that used the previous
  // occupant has finished rendering (in_flight_: rendered batches and the events that say when they are done).
  struct InFlight { cudaEvent_t done; uint64_t gen_lo; };
  std::deque<InFlight> in_flight_;
  bool popped_active_ = false;            // a batch Forward has popped and not yet queued (gpu) / finished (cpu) rendering
  uint64_t popped_gen_lo_ = 0;
  std::vector<std::pair<uint64_t, uint64_t> > drawn_;  // (ticket, first generation) of batches drawn but not queued yet
  std::vector<cudaEvent_t> event_pool_;
  int field_ring_ = 0;
  uint64_t next_generation_ = 0;          // first generation that has not been produced yet
  void EnsureFieldGenerations(uint64_t gen_lo, uint64_t gen_hi);
  void WaitGenerationRetired(uint64_t gen);
  void TrackInFlight(const Prefetched& b);
  uint32_t FieldSeed(uint64_t generation) const;
  size_t prefetch_depth_ = 1;
  std::mutex mutex_, generator_mutex_;
  std::condition_variable cv_full_, cv_free_, cv_push_;
  double stats_[3] = {0, 0, 0};
  std::thread thread_[kProducers];
  std::mutex draw_mutex_;                 // the parameter stream: batches are drawn one at a time, in ticket order
  uint64_t next_ticket_ = 0, next_push_ = 0;  // commission order of the batches = their order in the prefetch queue
  bool must_stop_ = false;
  unsigned long long device_batches_ = 0;  // batches produced by the device-side stream (its sample counter)
  std::string producer_error_;
  ofdg_extra_tops extra_{};               // what the generator currently writes besides the three blobs
  cudaStream_t stream_ = nullptr;         // Forward_gpu's stream (non-blocking)
  cudaEvent_t blobs_ready_ = nullptr, consumer_done_ = nullptr;
};

}  // namespace caffe
