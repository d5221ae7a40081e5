// Texture files -> planar B,G,R uint8, the form TextureCollection's constructor leaves its images in
// (/root/reference/src/caffe/DataGenerator.cpp:117-149: CImg::load, then channels 0 and 2 swapped).
// The reference decodes through CImg (libpng / libjpeg / ImageMagick); none of those is available here, so
// this file carries small decoders of its own: binary PPM (P6), uncompressed BMP (24 / 32 bit) and
// non-interlaced 8-bit PNG (gray, RGB, palette, with or without alpha; inflate through zlib). JPEG (baseline and
// progressive, the format of the authors' Flickr texture database) is decoded by nvJPEG on the current CUDA device.
// Not decoded: 16-bit and interlaced PNG (convert once, or upload decoded pixels through ofdg_add_textures).
// Gray images are replicated to three channels; alpha is dropped.
#pragma once
#include <string>
#include <vector>

namespace ofdg {

struct TextureImage {
  int w = 0, h = 0;
  std::vector<unsigned char> planar_bgr;  // 3 x h x w
};

// Throws std::runtime_error naming the file and the reason.
TextureImage load_texture_file(const std::string& path);

// One path per line with the reference's getline / eof semantics (DataGenerator.cpp:123-126): a last line without a
// trailing newline is not read; an empty line is an error.
// Throws "Could not open texture collection" (DataGenerator.cpp:121) when the list cannot be read.
std::vector<std::string> read_texture_list(const std::string& listfile);

}  // namespace ofdg
