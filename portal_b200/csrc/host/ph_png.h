// PNG in and out for the callers either side of the ray loop: textures arrive as PNG files
// (/root/reference/src/main.rs:1066-1085, Texture2D::from_file_with_format) and rendered frames leave as PNG
// (main.rs:2939-2943, Image::export_png).  Self-contained (no zlib / libpng in the image): RFC 1950 / 1951 / 2083
// written out here.  Host-side image IO only -- nothing of this is on the GPU path.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace ph {

// Decode a PNG into tightly packed RGBA8, top row first (what the reference uploads as a texture: the `image` crate's
// `to_rgba8()` of the decoded file; ancillary chunks -- gAMA, iCCP, ... -- are ignored there and here).
// Supported: colour types 0 / 2 / 3 / 4 / 6 at bit depths <= 8 (tRNS honoured), non-interlaced -- a superset of every
// asset under the reference's scenes/img (all 8-bit RGBA).  16-bit samples and Adam7 are refused with a message.
bool png_decode(const uint8_t* data, size_t len, std::vector<uint8_t>& rgba, int& width, int& height, std::string& err);

// Encode RGBA8 as a PNG (colour type 6, 8 bit, adaptive row filters, one zlib stream: LZ77 + fixed Huffman codes).
void png_encode_rgba8(const uint8_t* rgba, int width, int height, std::vector<uint8_t>& out);

// zlib pieces, exposed for the tests
bool zlib_inflate(const uint8_t* data, size_t len, std::vector<uint8_t>& out, std::string& err);
void zlib_deflate(const uint8_t* data, size_t len, std::vector<uint8_t>& out);
uint32_t crc32(const uint8_t* data, size_t len, uint32_t crc = 0);
uint32_t adler32(const uint8_t* data, size_t len);

}  // namespace ph
