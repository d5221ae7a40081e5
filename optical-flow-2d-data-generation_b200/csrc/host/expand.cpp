// See expand.hpp.
#include "expand.hpp"

#include <cuda_runtime_api.h>
#include <immintrin.h>

#include <cstdlib>
#include <stdexcept>

namespace ofdg {

namespace {

__attribute__((target("avx512f"))) void expand_avx512(const uint8_t* s, float* d, size_t n, bool streaming) {
  size_t i = 0;
  for (; i < n && ((uintptr_t)(d + i) & 63); ++i) d[i] = (float)s[i];
  if (streaming) {
    for (; i + 64 <= n; i += 64) {
      const __m512 f0 = _mm512_cvtepi32_ps(_mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i*)(s + i))));
      const __m512 f1 = _mm512_cvtepi32_ps(_mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i*)(s + i + 16))));
      const __m512 f2 = _mm512_cvtepi32_ps(_mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i*)(s + i + 32))));
      const __m512 f3 = _mm512_cvtepi32_ps(_mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i*)(s + i + 48))));
      _mm512_stream_ps(d + i, f0);
      _mm512_stream_ps(d + i + 16, f1);
      _mm512_stream_ps(d + i + 32, f2);
      _mm512_stream_ps(d + i + 48, f3);
    }
    _mm_sfence();
  } else {
    for (; i + 16 <= n; i += 16)
      _mm512_store_ps(d + i, _mm512_cvtepi32_ps(_mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i*)(s + i)))));
  }
  for (; i < n; ++i) d[i] = (float)s[i];
}

__attribute__((target("avx2"))) void expand_avx2(const uint8_t* s, float* d, size_t n, bool streaming) {
  size_t i = 0;
  for (; i < n && ((uintptr_t)(d + i) & 31); ++i) d[i] = (float)s[i];
  if (streaming) {
    for (; i + 16 <= n; i += 16) {
      const __m256 f0 = _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i*)(s + i))));
      const __m256 f1 = _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i*)(s + i + 8))));
      _mm256_stream_ps(d + i, f0);
      _mm256_stream_ps(d + i + 8, f1);
    }
    _mm_sfence();
  } else {
    for (; i + 8 <= n; i += 8)
      _mm256_store_ps(d + i, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i*)(s + i)))));
  }
  for (; i < n; ++i) d[i] = (float)s[i];
}

void expand_scalar(const uint8_t* s, float* d, size_t n) {
  for (size_t i = 0; i < n; ++i) d[i] = (float)s[i];
}

}  // namespace

void expand_u8_to_f32(const uint8_t* src, float* dst, size_t n, bool streaming) {
  static const int level = __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
  if (level == 2) expand_avx512(src, dst, n, streaming);
  else if (level == 1) expand_avx2(src, dst, n, streaming);
  else expand_scalar(src, dst, n);
}

HostPool::HostPool(int threads, int device) : device_(device) {
  const char* e = std::getenv("OFDG_EXPAND_STREAMING");
  streaming_ = e ? std::atoi(e) != 0 : true;
  if (threads < 1) threads = 1;
  for (int i = 0; i < threads; ++i) workers_.emplace_back([this] { run(); });
  waiter_ = std::thread([this] { run_waiter(); });
}

HostPool::~HostPool() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_work_.notify_all();
  cv_deferred_.notify_all();
  for (std::thread& t : workers_) t.join();
  waiter_.join();
}

void HostPool::submit(std::function<void()> job, bool urgent) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    (urgent ? urgent_ : queue_).push_back(std::move(job));
    ++pending_;
  }
  cv_work_.notify_one();
}

void HostPool::submit_after(void* ready, std::vector<std::function<void()>> jobs) {
  if (jobs.empty()) return;
  {
    std::lock_guard<std::mutex> lk(mu_);
    pending_ += jobs.size();
    deferred_.push_back(Deferred{ready, std::move(jobs)});
  }
  cv_deferred_.notify_one();
}

std::function<void()> HostPool::expand_job(const uint8_t* src, float* dst, size_t n) const {
  const bool streaming = streaming_;
  return [=] { expand_u8_to_f32(src, dst, n, streaming); };
}

void HostPool::wait() {
  std::unique_lock<std::mutex> lk(mu_);
  cv_idle_.wait(lk, [this] { return pending_ == 0; });
  if (!error_.empty()) {
    std::string e;
    e.swap(error_);
    throw std::runtime_error(e);
  }
}

void HostPool::finish(size_t n_jobs, const std::string& err) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!err.empty() && error_.empty()) error_ = err;
  pending_ -= n_jobs;
  if (pending_ == 0) cv_idle_.notify_all();
}

void HostPool::run_waiter() {
  cudaSetDevice(device_);
  for (;;) {
    Deferred d;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_deferred_.wait(lk, [this] { return stop_ || !deferred_.empty(); });
      if (deferred_.empty()) return;  // stop_
      d = std::move(deferred_.front());
      deferred_.pop_front();
    }
    const cudaError_t rc = d.ready ? cudaEventSynchronize((cudaEvent_t)d.ready) : cudaSuccess;
    if (rc != cudaSuccess) {
      finish(d.jobs.size(), std::string("cudaEventSynchronize (host pipeline): ") + cudaGetErrorString(rc));
      continue;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto& j : d.jobs) queue_.push_back(std::move(j));
    }
    cv_work_.notify_all();
  }
}

void HostPool::run() {
  for (;;) {
    std::function<void()> job;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_work_.wait(lk, [this] { return stop_ || !urgent_.empty() || !queue_.empty(); });
      std::deque<std::function<void()>>& q = !urgent_.empty() ? urgent_ : queue_;
      if (q.empty()) return;  // stop_
      job = std::move(q.front());
      q.pop_front();
    }
    std::string err;
    try {
      job();
    } catch (const std::exception& e) {
      err = e.what();
      if (err.empty()) err = "host job failed";
    } catch (...) {
      err = "host job failed";
    }
    finish(1, err);
  }
}

}  // namespace ofdg
