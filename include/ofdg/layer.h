/*
 * ofdg/layer.h -- C entry points that drive the Caffe-style DataGenerationLayer
 * (optical-flow-2d-data-generation_b200/csrc/host/layer.hpp) from a host that cannot include C++
 * headers. A C++ host uses caffe::DataGenerationLayer<float> directly, exactly like the reference's
 * /root/reference/include/caffe/layers/data_generation_layer.hpp:37-89.
 * All functions return 0 on success; message via ofdg_layer_last_error().
 */
#ifndef OFDG_LAYER_H_
#define OFDG_LAYER_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* ofdg_layer_last_error(void);
/* Parses a prototxt `layer { ... }` block (the reference's example-prototxt/train.prototxt is accepted verbatim).
 * ints[7] = {batch_size, prefetch, mode, first_level_threads, second_level_threads, use_antialiasing, top_size}. */
int ofdg_layer_parse_prototxt(const char* text, int32_t* ints, char* texture_db, int32_t cap, char* type, int32_t type_cap);
/* DataGenerationLayer(const LayerParameter&) on the current CUDA device. texture_db_override (may be NULL)
 * replaces texture_dbases(0): either a list file of image paths (binary PPM, uncompressed BMP, 8-bit PNG, JPEG; any
 * mix of sizes) or "synthetic:<count>[:<seed>]". */
int ofdg_layer_create(const char* prototxt, const char* texture_db_override, int32_t solver_rank, void** out);
void ofdg_layer_destroy(void* layer);
int ofdg_layer_setup(void* layer);                               /* Layer::SetUp -> LayerSetUp: starts prefetching, shapes the 3 tops */
int ofdg_layer_top_shape(void* layer, int32_t i, int32_t* shape4);
int ofdg_layer_forward(void* layer, int32_t gpu);                /* Forward_gpu (1) / Forward_cpu (0) */
const float* ofdg_layer_top_data(void* layer, int32_t i, int32_t gpu); /* top[i]->gpu_data() / cpu_data() */
const char* ofdg_layer_type(void* layer);                        /* "DataGeneration" */
/* Comma-separated type strings in the shim's caffe::LayerRegistry<float> (REGISTER_LAYER_CLASS(DataGeneration),
 * /root/reference/src/caffe/layers/data_generation_layer.cpp:298-299): ofdg_layer_create builds the layer through
 * LayerRegistry<float>::CreateLayer(param), the way Net::Init does. Returns the number of types. */
int ofdg_layer_registered_types(char* out, int32_t cap);
/* Prefetch-side timing since the last call: out3 = {ms drawing parameters (incl. waiting for the sequential stream), ms in
 * ofdg_prepare (flatten + upload), batches produced}. */
int ofdg_layer_producer_stats(void* layer, double* out3);
/* The texture-file decoder the layer uses (TextureCollection ctor, DataGenerator.cpp:128-133): size of the image,
 * and, when `planar_bgr` is non-NULL and `cap` >= 3*w*h, its pixels as 3 x h x w planes in B,G,R order. PPM, BMP and PNG
 * need no GPU; JPEG is decoded by nvJPEG on the current CUDA device. */
int ofdg_decode_texture_file(const char* path, int32_t* w, int32_t* h, uint8_t* planar_bgr, uint64_t cap);
/* The paths a texture list names, with the reference's getline / eof semantics (TextureCollection ctor,
 * DataGenerator.cpp:123-126: a last line without a trailing newline is not read). `out` receives them newline-separated. */
int ofdg_read_texture_list(const char* listfile, char* out, int32_t cap, int32_t* count);

#ifdef __cplusplus
}
#endif
#endif
