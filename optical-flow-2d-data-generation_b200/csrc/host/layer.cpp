// See layer.hpp / caffe_shim.hpp.
#include "layer.hpp"
#include "texture_io.hpp"

#include <cuda_runtime.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace caffe {

#define SHIM_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t e_ = (expr);                                                                         \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// ---- Blob -------------------------------------------------------------------------------------------
template <typename Dtype>
Blob<Dtype>::~Blob() {
  if (cpu_) cudaFreeHost(cpu_);
  if (gpu_) cudaFree(gpu_);
}
template <typename Dtype>
void Blob<Dtype>::Reshape(const std::vector<int>& shape) {
  shape_ = shape;
  size_t n = 1;
  for (int d : shape) n *= (size_t)d;
  count_ = (int)n;
  if (n > capacity_) {
    if (cpu_) cudaFreeHost(cpu_);
    if (gpu_) cudaFree(gpu_);
    cpu_ = nullptr; gpu_ = nullptr;
    capacity_ = n;
    head_ = UNINITIALIZED;
  }
}
template <typename Dtype>
void Blob<Dtype>::to_cpu() {
  if (!cpu_) {
    SHIM_CUDA(cudaMallocHost((void**)&cpu_, capacity_ * sizeof(Dtype)));
    if (head_ == UNINITIALIZED) std::memset(cpu_, 0, capacity_ * sizeof(Dtype));
  }
  if (head_ == HEAD_AT_GPU) {
    SHIM_CUDA(cudaMemcpy(cpu_, gpu_, (size_t)count_ * sizeof(Dtype), cudaMemcpyDeviceToHost));
    head_ = SYNCED;
  }
}
template <typename Dtype>
void Blob<Dtype>::to_gpu() {
  if (!gpu_) {
    SHIM_CUDA(cudaMalloc((void**)&gpu_, capacity_ * sizeof(Dtype)));
    if (head_ == UNINITIALIZED) SHIM_CUDA(cudaMemset(gpu_, 0, capacity_ * sizeof(Dtype)));
  }
  if (head_ == HEAD_AT_CPU) {
    SHIM_CUDA(cudaMemcpy(gpu_, cpu_, (size_t)count_ * sizeof(Dtype), cudaMemcpyHostToDevice));
    head_ = SYNCED;
  }
}
template <typename Dtype>
const Dtype* Blob<Dtype>::cpu_data() { to_cpu(); if (head_ == UNINITIALIZED) head_ = HEAD_AT_CPU; return cpu_; }
template <typename Dtype>
Dtype* Blob<Dtype>::mutable_cpu_data() { to_cpu(); head_ = HEAD_AT_CPU; return cpu_; }
template <typename Dtype>
const Dtype* Blob<Dtype>::gpu_data() { to_gpu(); if (head_ == UNINITIALIZED) head_ = HEAD_AT_GPU; return gpu_; }
template <typename Dtype>
Dtype* Blob<Dtype>::mutable_gpu_data() { to_gpu(); head_ = HEAD_AT_GPU; return gpu_; }
template class Blob<float>;

// ---- prototxt text parser -----------------------------------------------------------------------------
namespace {
struct Tok {
  enum K { IDENT, STRING, NUMBER, LBRACE, RBRACE, COLON, END } k;
  std::string s;
};
struct Lexer {
  const std::string& t;
  size_t i = 0;
  explicit Lexer(const std::string& text) : t(text) {}
  Tok next() {
    for (;;) {
      while (i < t.size() && std::isspace((unsigned char)t[i])) ++i;
      if (i < t.size() && t[i] == '#') { while (i < t.size() && t[i] != '\n') ++i; continue; }
      break;
    }
    if (i >= t.size()) return {Tok::END, ""};
    char c = t[i];
    if (c == '{') { ++i; return {Tok::LBRACE, "{"}; }
    if (c == '}') { ++i; return {Tok::RBRACE, "}"}; }
    if (c == ':') { ++i; return {Tok::COLON, ":"}; }
    if (c == '"' || c == '\'') {
      size_t j = ++i;
      std::string out;
      while (j < t.size() && t[j] != c) { if (t[j] == '\\' && j + 1 < t.size()) ++j; out += t[j++]; }
      if (j >= t.size()) throw std::runtime_error("prototxt: unterminated string");
      i = j + 1;
      return {Tok::STRING, out};
    }
    size_t j = i;
    while (j < t.size() && !std::isspace((unsigned char)t[j]) && t[j] != '{' && t[j] != '}' && t[j] != ':' && t[j] != '#') ++j;
    std::string w = t.substr(i, j - i);
    i = j;
    bool num = !w.empty() && (std::isdigit((unsigned char)w[0]) || w[0] == '-' || w[0] == '+' || w[0] == '.');
    return {num ? Tok::NUMBER : Tok::IDENT, w};
  }
};
int to_int(const Tok& v, const std::string& f) {
  if (v.k != Tok::NUMBER) throw std::runtime_error("prototxt: field '" + f + "' expects an integer");
  return std::atoi(v.s.c_str());
}
bool to_bool(const Tok& v, const std::string& f) {
  if (v.s == "true" || v.s == "1") return true;
  if (v.s == "false" || v.s == "0") return false;
  throw std::runtime_error("prototxt: field '" + f + "' expects true/false");
}
std::string to_str(const Tok& v, const std::string& f) {
  if (v.k != Tok::STRING) throw std::runtime_error("prototxt: field '" + f + "' expects a quoted string");
  return v.s;
}
// parses `name: value` or `name { ... }` pairs until the closing brace
template <class F>
void parse_message(Lexer& lx, bool top_level, F&& field) {
  for (;;) {
    Tok name = lx.next();
    if (name.k == Tok::END) { if (top_level) return; throw std::runtime_error("prototxt: missing '}'"); }
    if (name.k == Tok::RBRACE) { if (top_level) throw std::runtime_error("prototxt: stray '}'"); return; }
    if (name.k != Tok::IDENT) throw std::runtime_error("prototxt: expected a field name, got '" + name.s + "'");
    Tok t = lx.next();
    if (t.k == Tok::COLON) {
      Tok v = lx.next();
      if (v.k == Tok::LBRACE) field(name.s, nullptr, lx);
      else field(name.s, &v, lx);
    } else if (t.k == Tok::LBRACE) {
      field(name.s, nullptr, lx);
    } else {
      throw std::runtime_error("prototxt: expected ':' or '{' after '" + name.s + "'");
    }
  }
}
void skip_message(Lexer& lx) {
  parse_message(lx, false, [](const std::string&, const Tok* v, Lexer& l) { if (!v) skip_message(l); });
}
}  // namespace

LayerParameter ParseLayerPrototxt(const std::string& text) {
  LayerParameter p;
  Lexer lx(text);
  bool seen_layer = false, seen_mode = false;
  auto layer_fields = [&](const std::string& n, const Tok* v, Lexer& l) {
    if (!v && (n == "name" || n == "type" || n == "top" || n == "bottom"))  // e.g. `top { }`: a scalar field written as a message
      throw std::runtime_error("prototxt: field '" + n + "' takes a value, not a message");
    if (n == "name") p.name_ = to_str(*v, n);
    else if (n == "type") p.type_ = to_str(*v, n);
    else if (n == "top") p.top_.push_back(to_str(*v, n));
    else if (n == "bottom") p.bottom_.push_back(to_str(*v, n));
    else if (n == "data_param") {
      if (v) throw std::runtime_error("prototxt: data_param is a message");
      parse_message(l, false, [&](const std::string& f, const Tok* fv, Lexer&) {
        if (!fv) throw std::runtime_error("prototxt: unexpected message '" + f + "' in data_param");
        if (f == "batch_size") p.data_param_.batch_size_ = to_int(*fv, f);
        else if (f == "prefetch") p.data_param_.prefetch_ = to_int(*fv, f);
        else if (f == "block_size") p.data_param_.block_size_ = to_int(*fv, f);
        else if (f == "verbose") p.data_param_.verbose_ = to_bool(*fv, f);
        else if (f == "sample") p.data_param_.sample_.push_back(fv->s);
        else throw std::runtime_error("prototxt: unknown data_param field '" + f + "'");
      });
    } else if (n == "data_generation_param") {
      if (v) throw std::runtime_error("prototxt: data_generation_param is a message");
      parse_message(l, false, [&](const std::string& f, const Tok* fv, Lexer&) {
        if (!fv) throw std::runtime_error("prototxt: unexpected message '" + f + "' in data_generation_param");
        DataGenerationParameter& g = p.data_generation_param_;
        if (f == "mode") { g.mode_ = to_int(*fv, f); seen_mode = true; }
        else if (f == "texture_dbases") g.texture_dbases_.push_back(to_str(*fv, f));
        else if (f == "first_level_threads") g.first_level_threads_ = to_int(*fv, f);
        else if (f == "second_level_threads") g.second_level_threads_ = to_int(*fv, f);
        else if (f == "use_antialiasing") g.use_antialiasing_ = to_bool(*fv, f);
        else if (f == "device_params") g.device_params_ = to_bool(*fv, f);
        else if (f == "seed") g.seed_ = std::strtoull(fv->s.c_str(), nullptr, 10);
        else throw std::runtime_error("prototxt: unknown data_generation_param field '" + f + "'");
      });
    } else if (n == "include" || n == "exclude" || n == "phase") {
      if (!v) skip_message(l);  // net-level plumbing, irrelevant to the layer itself
    } else {
      throw std::runtime_error("prototxt: unknown layer field '" + n + "'");
    }
  };
  parse_message(lx, true, [&](const std::string& n, const Tok* v, Lexer& l) {
    if (n != "layer" && n != "layers") throw std::runtime_error("prototxt: expected a 'layer { ... }' block, got '" + n + "'");
    if (v) throw std::runtime_error("prototxt: 'layer' is a message");
    if (seen_layer) throw std::runtime_error("prototxt: more than one layer block");
    seen_layer = true;
    parse_message(l, false, layer_fields);
  });
  if (!seen_layer) throw std::runtime_error("prototxt: no layer block found");
  (void)seen_mode;  // `required` in the .proto but carries a default; protobuf text format would insist, Caffe users rely on the default
  return p;
}

// ---- DataGenerationLayer ----------------------------------------------------------------------------------
#define OFDG_CHECK(expr)                                                                 \
  do {                                                                                   \
    if ((expr) != OFDG_OK) throw std::runtime_error(std::string(ofdg_last_error()));     \
  } while (0)

template <typename Dtype>
int DataGenerationLayer<Dtype>::solver_rank_ = 0;

template <typename Dtype>
DataGenerationLayer<Dtype>::DataGenerationLayer(const LayerParameter& param) : Layer<Dtype>(param) {
  const DataGenerationParameter& gp = param.data_generation_param();
  if (param.data_param().batch_size() <= 0) throw std::runtime_error("data_param.batch_size must be positive");
  if (gp.texture_dbases_size() < 1) throw std::runtime_error("data_generation_param.texture_dbases(0) is required");  // DataGenerator.cpp:992
  prefetch_depth_ = (size_t)std::max(1, std::min(param.data_param().prefetch(), 64));
  SHIM_CUDA(cudaGetDevice(&device_));
  ofdg_config cfg{};
  cfg.device = device_;
  cfg.width = 512; cfg.height = 384;  // DGEN_WIDTH / DGEN_HEIGHT, DataGenerator.h:55-56
  cfg.mode = gp.mode();
  cfg.use_antialiasing = gp.use_antialiasing() ? 1 : 0;
  cfg.max_batch = param.data_param().batch_size();
  OFDG_CHECK(ofdg_create(&cfg, &generator_));
  const std::string& db = gp.texture_dbases(0);
  if (db.compare(0, 10, "synthetic:") == 0) {  // "synthetic:<count>[:<seed>]": procedural pool generated on the device
    int count = 0; unsigned long long seed = 0;
    if (std::sscanf(db.c_str() + 10, "%d:%llu", &count, &seed) < 1 || count <= 0) throw std::runtime_error("bad synthetic texture spec: " + db);
    OFDG_CHECK(ofdg_synth_textures(generator_, count, 2 * cfg.width, 2 * cfg.height, seed));
  } else {
    // TextureCollection ctor, DataGenerator.cpp:117-149: every image the list names, whatever its size
    OFDG_CHECK(ofdg_clear_textures(generator_));
    for (const std::string& path : ofdg::read_texture_list(db)) {
      const ofdg::TextureImage t = ofdg::load_texture_file(path);
      OFDG_CHECK(ofdg_add_textures(generator_, t.planar_bgr.data(), 1, t.w, t.h));
    }
  }
  // Mode 9: the reference owns a WarpFields::CropGenerator that keeps producing (flow, iflow) crops on 10 CPU threads, seeded
  // from std::random_device; every crop is handed out three times, then dropped (DataGenerator.cpp:1016-1019,
  // WarpFields.cpp:516-538, 540-641). Here the pool is a ring of generations of kFieldPool crops (one 3*max(W,H) canvas each):
  // pick k of the parameter stream reads slot (k / 3) % pool size, and the producer thread regenerates a generation's slots on
  // the GPU (reproducible seeds) right before the first batch that picks from it (EnsureFieldGenerations). The device-side
  // parameter stream (device_params) draws its picks from the pool as installed here and does not refresh it.
  int n_fields = 0;
  if (gp.mode() == 9) {
    const int bs = param.data_param().batch_size();
    const double gens_per_batch = bs * (23 * 0.2 + 0.2) / (3.0 * kFieldPool);  // about a fifth of the objects (and backgrounds) deform
    field_ring_ = (int)std::min(96.0, std::ceil(gens_per_batch * ((double)prefetch_depth_ + 3.0)) + 4.0);
    if (gp.device_params()) field_ring_ = 1;
    n_fields = kFieldPool * field_ring_;
    OFDG_CHECK(ofdg_generate_fields(generator_, FieldSeed(0), kFieldPool, nullptr));  // generation 0; sizes the tables
    OFDG_CHECK(ofdg_reserve_fields(generator_, n_fields));
    next_generation_ = 1;
  }
  OFDG_CHECK(ofdg_params_create(gp.mode(), cfg.width, cfg.height, 45 * solver_rank_, n_fields, 0, &params_));
  {
    // look-ahead threads of the parameter stream: the engines' values for the batch after next are produced while the
    // prefetch threads walk / flatten the current ones (the stream itself is unchanged)
    int helpers = std::min(6, (int)std::thread::hardware_concurrency() / 3);
    if (const char* lw = std::getenv("LOCAL_WORLD_SIZE")) helpers = std::min(helpers, (int)std::thread::hardware_concurrency() / (3 * std::max(1, std::atoi(lw))));
    if (const char* t = std::getenv("OFDG_PARAM_THREADS")) helpers = std::atoi(t);
    if (helpers > 0 && !gp.device_params()) OFDG_CHECK(ofdg_params_set_threads(params_, helpers));
  }
  for (int i = 0; i < kProducers; ++i) OFDG_CHECK(ofdg_tasks_create(&tasks_[i]));
  // Forward_gpu enqueues on a stream of the layer's own and never blocks the solver thread: the legacy default stream (where
  // Caffe's other layers run) is made to wait for the blobs by an event, and the layer's stream waits for the default
  // stream's earlier work before it overwrites them.
  SHIM_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  SHIM_CUDA(cudaEventCreateWithFlags(&blobs_ready_, cudaEventDisableTiming));
  SHIM_CUDA(cudaEventCreateWithFlags(&consumer_done_, cudaEventDisableTiming));
}

template <typename Dtype>
DataGenerationLayer<Dtype>::~DataGenerationLayer() {
  StopInternalThread();
  if (stream_) cudaStreamSynchronize(stream_);
  for (const Prefetched& b : prefetch_full_) ofdg_prepared_destroy(b.scene);
  for (const InFlight& f : in_flight_) cudaEventDestroy(f.done);
  for (cudaEvent_t e : event_pool_) cudaEventDestroy(e);
  if (blobs_ready_) cudaEventDestroy(blobs_ready_);
  if (consumer_done_) cudaEventDestroy(consumer_done_);
  if (stream_) cudaStreamDestroy(stream_);
  for (int i = 0; i < kProducers; ++i) if (tasks_[i]) ofdg_tasks_destroy(tasks_[i]);
  if (params_) ofdg_params_destroy(params_);
  if (generator_) ofdg_destroy(generator_);
}

template <typename Dtype>
void DataGenerationLayer<Dtype>::LayerSetUp(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top) {
  if (!bottom.empty()) throw std::runtime_error("DataGeneration takes no bottom blobs");
  if (top.size() < 3) throw std::runtime_error("DataGeneration needs 3 top blobs (first image, second image, flow)");  // the reference indexes top[0..2]
  const int batch_size = this->layer_param_.data_param().batch_size();
  if (!this->layer_param_.data_generation_param().device_params()) StartInternalThread();  // the device stream needs no producer
  if (top.size() > 7) throw std::runtime_error("DataGeneration fills at most 7 top blobs");
  top[0]->Reshape({batch_size, 3, 384, 512});  // data_generation_layer.cpp:128-130
  top[1]->Reshape({batch_size, 3, 384, 512});
  top[2]->Reshape({batch_size, 2, 384, 512});
  BindExtraTops(top);
}

// Tops beyond the reference's three (MinTopBlobs() = 1 leaves the count open):
//   top[3] backward flow {N,2,H,W}   RenderCore::flow1 = computeFlowImage(objects, true), DataGenerator.cpp:801-818
//   top[4] occlusion     {N,1,H,W}   include/ofdg/ofdg.h
//   top[5], top[6]       {N,1,H,W}   RenderCore::index_image0 / index_image1 as float
template <typename Dtype>
void DataGenerationLayer<Dtype>::BindExtraTops(const std::vector<Blob<Dtype>*>& top) {
  if (top.size() <= 3) return;
  const int batch_size = this->layer_param_.data_param().batch_size();
  ofdg_extra_tops t{};
  top[3]->Reshape({batch_size, 2, 384, 512});
  t.flow_bw = top[3]->mutable_gpu_data();
  for (size_t i = 4; i < top.size(); ++i) top[i]->Reshape({batch_size, 1, 384, 512});
  if (top.size() > 4) t.occlusion = top[4]->mutable_gpu_data();
  if (top.size() > 5) t.id0 = top[5]->mutable_gpu_data();
  if (top.size() > 6) t.id1 = top[6]->mutable_gpu_data();
  if (t.flow_bw == extra_.flow_bw && t.occlusion == extra_.occlusion && t.id0 == extra_.id0 && t.id1 == extra_.id1) return;
  std::lock_guard<std::mutex> g(generator_mutex_);
  OFDG_CHECK(ofdg_set_extra_tops(generator_, &t));
  extra_ = t;
}

template <typename Dtype>
void DataGenerationLayer<Dtype>::StartInternalThread() {
  if (thread_[0].joinable()) return;
  must_stop_ = false;
  for (int i = 0; i < kProducers; ++i) thread_[i] = std::thread([this, i] { this->InternalThreadEntry(i); });
}
template <typename Dtype>
void DataGenerationLayer<Dtype>::StopInternalThread() {
  {
    std::lock_guard<std::mutex> l(mutex_);
    must_stop_ = true;
  }
  cv_free_.notify_all();
  cv_full_.notify_all();
  cv_push_.notify_all();
  for (int i = 0; i < kProducers; ++i) if (thread_[i].joinable()) thread_[i].join();
}

// Draws batch_size tasks in commission order, flattens them and uploads the scene (load_batch,
// data_generation_layer.cpp:183-216; the retrieval half of the reference's load_batch is now the
// kernel launch in Forward).
template <typename Dtype>
void DataGenerationLayer<Dtype>::load_batch(Prefetched* out, int producer, uint64_t* ticket) {
  const int batch_size = this->layer_param_.data_param().batch_size();
  ofdg_tasks* tasks = tasks_[producer];
  const auto t0 = std::chrono::steady_clock::now();
  auto t1 = t0;
  {
    // The parameter stream is sequential (45 mt19937 engines consumed in commission order): one batch is drawn at a time.
    // The other producer thread meanwhile flattens and uploads the batch it drew before.
    std::lock_guard<std::mutex> draw(draw_mutex_);
    {
      std::lock_guard<std::mutex> l(mutex_);
      *ticket = next_ticket_++;
    }
    ofdg_tasks_clear(tasks);
    const uint64_t d0 = ofdg_params_field_draws(params_);
    OFDG_CHECK(ofdg_params_generate(params_, batch_size, tasks));
    const uint64_t d1 = ofdg_params_field_draws(params_);
    if (d1 > d0) {  // mode 9: the generations this batch picks from must hold their crops (and their reach) before it is flattened
      out->uses_fields = true;
      out->gen_lo = (d0 / 3) / kFieldPool;
      out->gen_hi = ((d1 - 1) / 3) / kFieldPool;
      EnsureFieldGenerations(out->gen_lo, out->gen_hi);
      std::lock_guard<std::mutex> l(mutex_);
      drawn_.push_back(std::make_pair(*ticket, out->gen_lo));  // drawn, not queued yet: its generations must stay as they are
    }
    t1 = std::chrono::steady_clock::now();
  }
  // (no generator lock: ofdg_prepare flattens on its own host pool and uploads on its own stream, beside a running Forward)
  ofdg_task_batch view;
  OFDG_CHECK(ofdg_tasks_view(tasks, &view));
  const auto t2 = std::chrono::steady_clock::now();
  OFDG_CHECK(ofdg_prepare(generator_, &view, &out->scene));
  const auto t3 = std::chrono::steady_clock::now();
  std::lock_guard<std::mutex> l(mutex_);
  stats_[0] += std::chrono::duration<double, std::milli>(t1 - t0).count();  // waiting for the stream + drawing
  stats_[1] += std::chrono::duration<double, std::milli>(t3 - t2).count();  // flatten + upload (incl. waiting for its turn)
  stats_[2] += 1.0;
}

template <typename Dtype>
void DataGenerationLayer<Dtype>::producer_stats(double* out3) {
  std::lock_guard<std::mutex> l(mutex_);
  for (int i = 0; i < 3; ++i) { out3[i] = stats_[i]; stats_[i] = 0; }
}

template <typename Dtype>
uint32_t DataGenerationLayer<Dtype>::FieldSeed(uint64_t generation) const {
  const DataGenerationParameter& gp = this->layer_param_.data_generation_param();
  return (uint32_t)(gp.seed() + 7919u * (unsigned)solver_rank_ + 1u + 104729u * (uint32_t)generation);
}

// Producer thread. Generation G lives in slots (G % ring) * kFieldPool ...; it replaces generation G - ring, so every batch
// that picked from that one must have finished rendering first.
template <typename Dtype>
void DataGenerationLayer<Dtype>::EnsureFieldGenerations(uint64_t gen_lo, uint64_t gen_hi) {
  if (gen_hi - gen_lo + 2 > (uint64_t)field_ring_)
    throw std::runtime_error("mode 9: one batch picks from more warp-field generations than the pool ring holds (lower batch_size or prefetch)");
  for (uint64_t G = std::max(gen_lo, next_generation_); G <= gen_hi; ++G) {
    if (G >= (uint64_t)field_ring_) WaitGenerationRetired(G - (uint64_t)field_ring_);
    OFDG_CHECK(ofdg_refresh_fields(generator_, FieldSeed(G), (int)(G % (uint64_t)field_ring_) * kFieldPool, kFieldPool));
    next_generation_ = G + 1;
  }
}

// Blocks the producer until no batch that picked from generations <= gen is queued or still rendering.
template <typename Dtype>
void DataGenerationLayer<Dtype>::WaitGenerationRetired(uint64_t gen) {
  for (;;) {
    cudaEvent_t wait_for = nullptr;
    {
      std::unique_lock<std::mutex> l(mutex_);
      if (!in_flight_.empty()) {
        if (in_flight_.front().gen_lo > gen) return;  // (ranges grow with the commission order: nothing older is left)
        wait_for = in_flight_.front().done;
      } else {
        bool queued = false;
        for (const Prefetched& b : prefetch_full_) queued = queued || (b.uses_fields && b.gen_lo <= gen);
        for (const std::pair<uint64_t, uint64_t>& d : drawn_) queued = queued || d.second <= gen;  // still with the other producer
        queued = queued || (popped_active_ && popped_gen_lo_ <= gen);                              // popped by Forward, render not queued / done yet
        if (!queued) return;
        // still waiting in the queue (or on its way there): Forward will pop it (cv_free_ is signalled on every pop)
        cv_free_.wait_for(l, std::chrono::milliseconds(2), [this] { return must_stop_ || !in_flight_.empty(); });
        if (must_stop_) throw std::runtime_error("stopped");
        continue;
      }
    }
    cudaEventSynchronize(wait_for);
    std::lock_guard<std::mutex> l(mutex_);
    if (!in_flight_.empty() && in_flight_.front().done == wait_for) {
      event_pool_.push_back(wait_for);
      in_flight_.pop_front();
    }
  }
}

// Solver thread, after the render of `b` has been queued on stream_.
template <typename Dtype>
void DataGenerationLayer<Dtype>::TrackInFlight(const Prefetched& b) {
  if (!b.uses_fields) return;
  std::lock_guard<std::mutex> l(mutex_);
  popped_active_ = false;
  if (field_ring_ <= 1) return;
  cudaEvent_t e = nullptr;
  if (!event_pool_.empty()) { e = event_pool_.back(); event_pool_.pop_back(); }
  else SHIM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  SHIM_CUDA(cudaEventRecord(e, stream_));
  in_flight_.push_back(InFlight{e, b.gen_lo});
  cv_free_.notify_all();
}

template <typename Dtype>
void DataGenerationLayer<Dtype>::InternalThreadEntry(int producer) {
  cudaSetDevice(device_);
  try {
    for (;;) {
      {
        // room for this batch once the ones ahead of it are in: queued + being produced <= prefetch depth
        std::unique_lock<std::mutex> l(mutex_);
        cv_free_.wait(l, [this] { return must_stop_ || prefetch_full_.size() + (size_t)(next_ticket_ - next_push_) < prefetch_depth_; });
        if (must_stop_) return;
      }
      Prefetched b;
      uint64_t ticket = 0;
      load_batch(&b, producer, &ticket);
      {
        std::unique_lock<std::mutex> l(mutex_);
        cv_push_.wait(l, [this, ticket] { return must_stop_ || next_push_ == ticket; });  // commission order
        if (must_stop_) { ofdg_prepared_destroy(b.scene); return; }
        prefetch_full_.push_back(b);
        ++next_push_;
        for (size_t i = 0; i < drawn_.size(); ++i)
          if (drawn_[i].first == ticket) { drawn_.erase(drawn_.begin() + (long)i); break; }
      }
      cv_push_.notify_all();
      cv_full_.notify_one();
    }
  } catch (const std::exception& e) {
    std::lock_guard<std::mutex> l(mutex_);
    if (producer_error_.empty()) producer_error_ = e.what();
    must_stop_ = true;
    cv_full_.notify_all();
    cv_push_.notify_all();
    cv_free_.notify_all();
  }
}

// prefetch_full_.pop("Data layer prefetch queue empty"), data_generation_layer.cpp:270
template <typename Dtype>
typename DataGenerationLayer<Dtype>::Prefetched DataGenerationLayer<Dtype>::PopPrefetched() {
  Prefetched b;
  {
    std::unique_lock<std::mutex> l(mutex_);
    cv_full_.wait(l, [this] { return !prefetch_full_.empty() || !producer_error_.empty() || must_stop_; });
    if (!producer_error_.empty()) throw std::runtime_error(producer_error_);
    if (prefetch_full_.empty()) throw std::runtime_error("Data layer prefetch queue empty");
    b = prefetch_full_.front();
    prefetch_full_.pop_front();
    if (b.uses_fields) { popped_active_ = true; popped_gen_lo_ = b.gen_lo; }  // neither queued nor in flight yet, but its generations are in use
  }
  cv_free_.notify_all();
  return b;
}

// Orders the layer's stream after everything the consumer has queued on the legacy default stream so far (it may still be
// reading the top blobs of the previous iteration); only the kernel that writes the blobs waits, the batch's front end
// (background preparation, mask rasterisation) runs beside the consumer's work.
template <typename Dtype>
void DataGenerationLayer<Dtype>::BeginForward() {
  SHIM_CUDA(cudaEventRecord(consumer_done_, cudaStreamLegacy));
  SHIM_CUDA(cudaStreamWaitEvent(stream_, consumer_done_, 0));
}
// ... and the default stream after the blobs: whatever the consumer launches next sees them complete. No host blocking.
template <typename Dtype>
void DataGenerationLayer<Dtype>::EndForward() {
  SHIM_CUDA(cudaEventRecord(blobs_ready_, stream_));
  SHIM_CUDA(cudaStreamWaitEvent(cudaStreamLegacy, blobs_ready_, 0));
}

template <typename Dtype>
void DataGenerationLayer<Dtype>::Forward_gpu(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top) {
  const DataGenerationParameter& gp = this->layer_param_.data_generation_param();
  const int bs = this->layer_param_.data_param().batch_size();
  top[0]->Reshape({bs, 3, 384, 512});
  top[1]->Reshape({bs, 3, 384, 512});
  top[2]->Reshape({bs, 2, 384, 512});
  BindExtraTops(top);
  if (gp.device_params()) {
    // production mode: parameters drawn and flattened on the device, straight into the top blobs
    const unsigned long long seed = gp.seed() ^ ((unsigned long long)solver_rank_ << 40);  // distinct streams per replica
    BeginForward();
    if (ofdg_generate_philox(generator_, seed, device_batches_ * (unsigned long long)bs, bs, 0, top[0]->mutable_gpu_data(),
                             top[1]->mutable_gpu_data(), top[2]->mutable_gpu_data(), stream_) != OFDG_OK)
      throw std::runtime_error(ofdg_last_error());
    EndForward();
    ++device_batches_;
    return;
  }
  const Prefetched b = PopPrefetched();
  ofdg_prepared* p = b.scene;
  BeginForward();
  int rc;
  {
    std::lock_guard<std::mutex> g(generator_mutex_);
    rc = ofdg_render_prepared(generator_, p, top[0]->mutable_gpu_data(), top[1]->mutable_gpu_data(), top[2]->mutable_gpu_data(), stream_);
  }
  EndForward();
  TrackInFlight(b);
  ofdg_prepared_destroy(p);  // back to the generator's free list; its buffers are reused once this render is done
  if (rc != OFDG_OK) throw std::runtime_error(ofdg_last_error());
}

template <typename Dtype>
void DataGenerationLayer<Dtype>::Forward_cpu(const std::vector<Blob<Dtype>*>& bottom, const std::vector<Blob<Dtype>*>& top) {
  // There is no CPU generator any more. With the reference's three tops and the host parameter stream the blobs are
  // produced straight into cpu_data() by the pipelined host-blob path (frames cross PCIe as bytes, host threads widen them:
  // the copy-out of Process_TaskBucket, DataGenerator.cpp:1228-1244). Otherwise (extra tops, device-side stream) they are
  // produced on the device and become visible through the usual synced-memory copy.
  const DataGenerationParameter& gp = this->layer_param_.data_generation_param();
  if (top.size() != 3 || gp.device_params()) {
    Forward_gpu(bottom, top);
    for (size_t i = 0; i < top.size(); ++i) top[i]->cpu_data();
    return;
  }
  const Prefetched b = PopPrefetched();  // (rendered synchronously below: nothing of it is in flight afterwards)
  ofdg_prepared* p = b.scene;
  const int batch_size = this->layer_param_.data_param().batch_size();
  top[0]->Reshape({batch_size, 3, 384, 512});
  top[1]->Reshape({batch_size, 3, 384, 512});
  top[2]->Reshape({batch_size, 2, 384, 512});
  int rc;
  {
    std::lock_guard<std::mutex> g(generator_mutex_);
    rc = ofdg_render_prepared_host(generator_, p, top[0]->mutable_cpu_data(), top[1]->mutable_cpu_data(), top[2]->mutable_cpu_data());
  }
  ofdg_prepared_destroy(p);
  if (b.uses_fields) {  // rendered synchronously: its field generations are free again
    std::lock_guard<std::mutex> l(mutex_);
    popped_active_ = false;
  }
  cv_free_.notify_all();
  if (rc != OFDG_OK) throw std::runtime_error(ofdg_last_error());
}

template <typename Dtype>
uint64_t DataGenerationLayer<Dtype>::tasks_commissioned() const { return ofdg_params_tasks_generated(params_); }

template class DataGenerationLayer<float>;
REGISTER_LAYER_CLASS(DataGeneration);  // src/caffe/layers/data_generation_layer.cpp:298-299

}  // namespace caffe

// ---- C wrappers so the layer can be driven through the C ABI (tests, foreign hosts) --------------------
namespace {
thread_local std::string g_layer_error;
struct LayerBox {
  std::shared_ptr<caffe::Layer<float>> layer;  // built by LayerRegistry<float>::CreateLayer from the prototxt's type string
  std::vector<std::unique_ptr<caffe::Blob<float>>> tops;
  std::vector<caffe::Blob<float>*> top_ptrs, bottom_ptrs;
  caffe::LayerParameter param;
};
template <class F>
int layer_guard(F&& f) {
  try { f(); return 0; } catch (const std::exception& e) { g_layer_error = e.what(); return 1; }
}
}  // namespace

extern "C" {
const char* ofdg_layer_last_error(void) { return g_layer_error.c_str(); }

int ofdg_decode_texture_file(const char* path, int32_t* w, int32_t* h, uint8_t* planar_bgr, uint64_t cap) {
  return layer_guard([&] {
    if (!path || !w || !h) throw std::runtime_error("null pointer");
    const ofdg::TextureImage t = ofdg::load_texture_file(path);
    *w = t.w; *h = t.h;
    if (planar_bgr) {
      if (cap < t.planar_bgr.size()) throw std::runtime_error("buffer too small");
      std::memcpy(planar_bgr, t.planar_bgr.data(), t.planar_bgr.size());
    }
  });
}

int ofdg_read_texture_list(const char* listfile, char* out, int32_t cap, int32_t* count) {
  return layer_guard([&] {
    if (!listfile || !count) throw std::runtime_error("null pointer");
    const std::vector<std::string> paths = ofdg::read_texture_list(listfile);
    *count = (int32_t)paths.size();
    std::string joined;
    for (const std::string& p : paths) joined += p + "\n";
    if (out && cap > 0) std::snprintf(out, (size_t)cap, "%s", joined.c_str());
  });
}

int ofdg_layer_parse_prototxt(const char* text, int32_t* ints /*[7]: batch,prefetch,mode,first,second,aa,n_top*/, char* texture_db, int32_t cap,
                              char* type, int32_t type_cap) {
  return layer_guard([&] {
    caffe::LayerParameter p = caffe::ParseLayerPrototxt(text);
    ints[0] = p.data_param().batch_size(); ints[1] = p.data_param().prefetch();
    ints[2] = p.data_generation_param().mode(); ints[3] = p.data_generation_param().first_level_threads();
    ints[4] = p.data_generation_param().second_level_threads(); ints[5] = p.data_generation_param().use_antialiasing();
    ints[6] = p.top_size();
    std::string db = p.data_generation_param().texture_dbases_size() ? p.data_generation_param().texture_dbases(0) : "";
    std::snprintf(texture_db, cap, "%s", db.c_str());
    std::snprintf(type, type_cap, "%s", p.type().c_str());
  });
}

int ofdg_layer_create(const char* prototxt, const char* texture_db_override, int32_t solver_rank, void** out) {
  return layer_guard([&] {
    std::unique_ptr<LayerBox> b(new LayerBox);
    b->param = caffe::ParseLayerPrototxt(prototxt);
    if (texture_db_override && *texture_db_override) b->param.data_generation_param_.texture_dbases_ = {texture_db_override};
    caffe::DataGenerationLayer<float>::set_solver_rank(solver_rank);
    b->layer = caffe::LayerRegistry<float>::CreateLayer(b->param);  // Net::Init's path: the type string picks the class
    for (int i = 0; i < std::max(3, b->param.top_size()); ++i) {
      b->tops.emplace_back(new caffe::Blob<float>());
      b->top_ptrs.push_back(b->tops.back().get());
    }
    *out = b.release();
  });
}
void ofdg_layer_destroy(void* l) { delete (LayerBox*)l; }
int ofdg_layer_setup(void* l) {
  return layer_guard([&] { LayerBox* b = (LayerBox*)l; b->layer->SetUp(b->bottom_ptrs, b->top_ptrs); });
}
int ofdg_layer_top_shape(void* l, int32_t i, int32_t* shape4) {
  return layer_guard([&] {
    LayerBox* b = (LayerBox*)l;
    const std::vector<int>& s = b->top_ptrs.at(i)->shape();
    for (int k = 0; k < 4; ++k) shape4[k] = k < (int)s.size() ? s[k] : 0;
  });
}
int ofdg_layer_forward(void* l, int32_t gpu) {
  return layer_guard([&] {
    LayerBox* b = (LayerBox*)l;
    if (gpu) b->layer->Forward_gpu(b->bottom_ptrs, b->top_ptrs); else b->layer->Forward_cpu(b->bottom_ptrs, b->top_ptrs);
  });
}
const float* ofdg_layer_top_data(void* l, int32_t i, int32_t gpu) {
  LayerBox* b = (LayerBox*)l;
  try { return gpu ? b->top_ptrs.at(i)->gpu_data() : b->top_ptrs.at(i)->cpu_data(); } catch (const std::exception& e) { g_layer_error = e.what(); return nullptr; }
}
const char* ofdg_layer_type(void* l) { return ((LayerBox*)l)->layer->type(); }
int ofdg_layer_producer_stats(void* l, double* out3) {
  return layer_guard([&] {
    caffe::DataGenerationLayer<float>* dl = dynamic_cast<caffe::DataGenerationLayer<float>*>(((LayerBox*)l)->layer.get());
    if (!dl || !out3) throw std::runtime_error("not a DataGeneration layer");
    dl->producer_stats(out3);
  });
}
int ofdg_layer_registered_types(char* out, int32_t cap) {
  std::string s;
  const std::vector<std::string> types = caffe::LayerRegistry<float>::LayerTypeList();
  for (const std::string& t : types) s += (s.empty() ? "" : ",") + t;
  if (out && cap > 0) std::snprintf(out, (size_t)cap, "%s", s.c_str());
  return (int)types.size();
}
}
