// Host half of the uint8 transport used by the host-blob entry points (ofdg_render_host /
// ofdg_generate_host): frames cross PCIe as the bytes they are (the renderer's pixels are integers
// 0..255 -- Process_TaskBucket composes CImg<unsigned char> frames and widens them with
// static_cast<float> on the host as its last step, /root/reference/src/caffe/DataGenerator.cpp:1228-1244)
// and are widened to the
// caller's float blobs by a small pool of host threads while the next chunk is in flight.
// (float)uint8 is exact, so the blobs are bit-identical to a float transfer.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace ofdg {

// dst[i] = (float)src[i]; AVX-512 / AVX2 when the CPU has them. `streaming` uses non-temporal stores
// (the blobs are far larger than the caches and are not re-read by the producer).
void expand_u8_to_f32(const uint8_t* src, float* dst, size_t n, bool streaming);

// Small pool of host threads shared by the stages of the host-blob pipeline that run beside the GPU:
// drawing parameters and flattening task batches (host/flatten.cpp, one job per sample) -- `urgent`, they feed
// the GPU -- and widening byte planes, which only become runnable once their device-to-host copy has landed.
class HostPool {
 public:
  HostPool(int threads, int device);
  ~HostPool();
  HostPool(const HostPool&) = delete;
  HostPool& operator=(const HostPool&) = delete;
  // Urgent jobs are served before all others.
  void submit(std::function<void()> job, bool urgent = false);
  // The jobs become runnable once CUDA event `ready` (a cudaEvent_t) has completed. One extra thread waits for
  // the events, in submission order, so that no worker sits blocked on the GPU.
  void submit_after(void* ready, std::vector<std::function<void()>> jobs);
  // Job body: dst[i] = (float)src[i] for i < n
  std::function<void()> expand_job(const uint8_t* src, float* dst, size_t n) const;
  // Blocks until everything submitted so far is done; throws std::runtime_error with the first job failure.
  void wait();
  int threads() const { return (int)workers_.size(); }

 private:
  struct Deferred { void* ready; std::vector<std::function<void()>> jobs; };
  void run();
  void run_waiter();
  void finish(size_t n_jobs, const std::string& err);
  int device_;
  bool streaming_;
  std::vector<std::thread> workers_;
  std::thread waiter_;
  std::deque<std::function<void()>> urgent_, queue_;
  std::deque<Deferred> deferred_;
  std::mutex mu_;
  std::condition_variable cv_work_, cv_idle_, cv_deferred_;
  size_t pending_ = 0;
  bool stop_ = false;
  std::string error_;
};

}  // namespace ofdg
