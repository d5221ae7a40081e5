// TEST INFRASTRUCTURE (oracle/shim): the slice of Caffe v1 (LMB fork) that
// /root/reference/src/caffe/layers/data_generation_layer.cpp and its header touch: Blob, Layer, InternalThread,
// BlockingQueue, caffe_copy, the glog macros and the layer-registration macros. CPU only (the reference's
// Forward_gpu is Forward_cpu, data_generation_layer.cpp:285-291). Caffe is not vendored by the reference.
#ifndef OFDG_ORACLE_CAFFE_SHIM_HPP_
#define OFDG_ORACLE_CAFFE_SHIM_HPP_
#ifndef CPU_ONLY
#define CPU_ONLY
#endif
#include <atomic>
#include <cassert>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "caffe/proto/caffe.pb.h"

namespace boost { struct thread_interrupted {}; }

namespace caffe {
using std::vector;
using std::string;

// ---- glog look-alikes: CHECK aborts, LOG streams to stderr (INFO is dropped unless OFDG_REF_VERBOSE is set)
struct LogLine {
  bool fatal, on;
  std::ostringstream s;
  LogLine(bool fatal_, bool on_) : fatal(fatal_), on(on_) {}
  ~LogLine() { if (on) std::cerr << s.str() << std::endl; if (fatal) std::abort(); }
  template <class T> LogLine& operator<<(const T& v) { if (on) s << v; return *this; }
};
inline bool log_verbose() { static const bool v = std::getenv("OFDG_REF_VERBOSE") != nullptr; return v; }
#define LOG(sev) ::caffe::LogLine(false, ::caffe::log_verbose())
#define DLOG(sev) ::caffe::LogLine(false, false)
#define CHECK(c) if (c) {} else ::caffe::LogLine(true, true) << "Check failed: " #c " "
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_EQ(a, b) CHECK((a) == (b))

template <typename Dtype> inline void caffe_copy(const int n, const Dtype* x, Dtype* y) { if (x != y) std::memcpy(y, x, sizeof(Dtype) * n); }

template <typename Dtype>
class Blob {
 public:
  Blob() : count_(0) {}
  void Reshape(const vector<int>& shape) {
    shape_ = shape;
    count_ = 1;
    for (int d : shape) count_ *= d;
    if ((int)data_.size() < count_) data_.resize(count_);
  }
  void Reshape(std::initializer_list<int> shape) { Reshape(vector<int>(shape)); }
  void ReshapeLike(const Blob& o) { Reshape(o.shape_); }
  const vector<int>& shape() const { return shape_; }
  int count() const { return count_; }
  int offset(int n, int c = 0, int h = 0, int w = 0) const { return ((n * shape_[1] + c) * shape_[2] + h) * shape_[3] + w; }
  const Dtype* cpu_data() const { return data_.data(); }
  Dtype* mutable_cpu_data() { return data_.data(); }
 private:
  vector<int> shape_;
  int count_;
  vector<Dtype> data_;
};

template <typename Dtype>
class Layer {
 public:
  explicit Layer(const LayerParameter& param) : layer_param_(param) {}
  virtual ~Layer() {}
  void SetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) { LayerSetUp(bottom, top); Reshape(bottom, top); }
  virtual void LayerSetUp(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) {}
  virtual void Reshape(const vector<Blob<Dtype>*>&, const vector<Blob<Dtype>*>&) = 0;
  virtual inline bool ShareInParallel() const { return false; }
  virtual inline const char* type() const { return ""; }
  virtual inline int ExactNumBottomBlobs() const { return -1; }
  virtual inline int MinTopBlobs() const { return -1; }
  void Forward(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) { Forward_cpu(bottom, top); }
 protected:
  virtual void Forward_cpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) = 0;
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) { Forward_cpu(bottom, top); }
  virtual void Backward_cpu(const vector<Blob<Dtype>*>&, const vector<bool>&, const vector<Blob<Dtype>*>&) = 0;
  virtual void Backward_gpu(const vector<Blob<Dtype>*>&, const vector<bool>&, const vector<Blob<Dtype>*>&) {}
  LayerParameter layer_param_;
};

class InternalThread {
 public:
  InternalThread() : stop_(false) {}
  virtual ~InternalThread() { StopInternalThread(); }
  void StartInternalThread() {
    stop_ = false;
    thread_.reset(new std::thread([this] { InternalThreadEntry(); }));
  }
  void StopInternalThread();
  bool must_stop() { return stop_; }
  // queues that may block the internal thread register a wake-up here
  std::vector<std::function<void()>> wakers_;
 protected:
  virtual void InternalThreadEntry() {}
 private:
  std::atomic<bool> stop_;
  std::unique_ptr<std::thread> thread_;
};
inline std::atomic<bool>& interrupt_flag() { static std::atomic<bool> f(false); return f; }
inline void InternalThread::StopInternalThread() {
  if (thread_ && thread_->joinable()) {
    stop_ = true;
    interrupt_flag() = true;  // boost::thread::interrupt(): blocked queue pops throw thread_interrupted
    thread_->join();
    interrupt_flag() = false;
  }
  thread_.reset();
}

template <typename T>
class BlockingQueue {
 public:
  void push(const T& t) { { std::lock_guard<std::mutex> l(m_); q_.push(t); } cv_.notify_one(); }
  T pop(const string& log_on_wait = "") {
    std::unique_lock<std::mutex> l(m_);
    while (q_.empty()) {
      if (interrupt_flag()) throw boost::thread_interrupted();
      cv_.wait_for(l, std::chrono::milliseconds(2));
    }
    T t = q_.front();
    q_.pop();
    return t;
  }
  size_t size() const { std::lock_guard<std::mutex> l(m_); return q_.size(); }
 private:
  mutable std::mutex m_;
  std::condition_variable cv_;
  std::queue<T> q_;
};

// layer registry (caffe/layer_factory.hpp): type string -> creator
template <typename Dtype>
class LayerRegistry {
 public:
  typedef std::shared_ptr<Layer<Dtype>> (*Creator)(const LayerParameter&);
  static std::map<string, Creator>& Registry() { static std::map<string, Creator> r; return r; }
  static void AddCreator(const string& type, Creator c) { Registry()[type] = c; }
  static std::shared_ptr<Layer<Dtype>> CreateLayer(const LayerParameter& p) {
    auto it = Registry().find(p.type());
    CHECK(it != Registry().end()) << "Unknown layer type: " << p.type();
    return it->second(p);
  }
};
template <typename Dtype>
struct LayerRegisterer {
  LayerRegisterer(const string& type, typename LayerRegistry<Dtype>::Creator c) { LayerRegistry<Dtype>::AddCreator(type, c); }
};
#define INSTANTIATE_CLASS(classname) template class classname<float>
#define STUB_GPU_FORWARD(classname, funcname)
#define REGISTER_LAYER_CLASS(type)                                                                                        \
  template <typename Dtype> std::shared_ptr<Layer<Dtype>> Creator_##type##Layer(const LayerParameter& p) {              \
    return std::shared_ptr<Layer<Dtype>>(new type##Layer<Dtype>(p));                                                     \
  }                                                                                                                       \
  static LayerRegisterer<float> g_creator_f_##type(#type, Creator_##type##Layer<float>)

}  // namespace caffe
#endif
