// image_writer.cpp -- float image -> 8-bit PNG / BMP with the reference's conventions, headless.
//
// Follows (paths relative to the reference repo root):
//   runCuda() save path        src/main.cpp:118-139   mirror x (pixel W-1-x <- image[x + y*W]); gamma{true, 1.0, 1}
//                                                     (identity: pow(f/1, 1)); ".<frame>" spliced before .bmp/.png
//   image::saveImageRGB        src/image.cpp:46-88    (unsigned char)clamp(f*255, 0, 255) (truncation), rows top-down,
//                                                     BMP iff the name ends in "bmp" (":68-80", the OSX branch tolerates
//                                                     a trailing '\r'), else PNG
//   utilityCore::clamp         src/utilities.cpp:15-23
// The reference writes files through the vendored stb_image_write; this writer produces standard PNG (zlib deflate,
// filter 0) and 24-bit BMP files with identical pixel content.
#include "../../include/pt_b200.h"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

extern "C" void pt_set_error_(const char* fmt, ...);

namespace {

inline unsigned char to_u8(float f) {
  // applyGamma with gamma 1.0 and divisor 1 is the identity (powf(x, 1.0f) == x); then clamp(f*255, 0, 255)
  float v = f * 255;
  if (v < 0) v = 0;
  else if (v > 255) v = 255;
  if (!(v == v)) v = 0;  // NaN: the reference's cast is undefined; x86 yields 0
  return (unsigned char)v;
}

void put32be(std::vector<unsigned char>& o, uint32_t v) {
  o.push_back((unsigned char)(v >> 24)); o.push_back((unsigned char)(v >> 16));
  o.push_back((unsigned char)(v >> 8)); o.push_back((unsigned char)v);
}

void png_chunk(std::vector<unsigned char>& o, const char* tag, const unsigned char* data, size_t n) {
  put32be(o, (uint32_t)n);
  size_t start = o.size();
  o.insert(o.end(), tag, tag + 4);
  if (n) o.insert(o.end(), data, data + n);
  uint32_t crc = (uint32_t)crc32(0L, o.data() + start, (uInt)(n + 4));
  put32be(o, crc);
}

int write_png(const char* path, int W, int H, const unsigned char* rgb8) {
  std::vector<unsigned char> raw((size_t)H * (1 + 3 * (size_t)W));
  for (int y = 0; y < H; y++) {
    unsigned char* row = raw.data() + (size_t)y * (1 + 3 * (size_t)W);
    row[0] = 0;  // filter: none
    memcpy(row + 1, rgb8 + (size_t)y * 3 * W, 3 * (size_t)W);
  }
  uLongf clen = compressBound((uLong)raw.size());
  std::vector<unsigned char> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) {
    pt_set_error_("zlib compress2 failed");
    return PT_ERR_IO;
  }
  std::vector<unsigned char> out;
  const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  out.insert(out.end(), sig, sig + 8);
  std::vector<unsigned char> ihdr;
  put32be(ihdr, (uint32_t)W);
  put32be(ihdr, (uint32_t)H);
  const unsigned char rest[5] = {8, 2, 0, 0, 0};  // 8-bit, truecolour, deflate, adaptive, no interlace
  ihdr.insert(ihdr.end(), rest, rest + 5);
  png_chunk(out, "IHDR", ihdr.data(), ihdr.size());
  png_chunk(out, "IDAT", comp.data(), clen);
  png_chunk(out, "IEND", nullptr, 0);
  FILE* f = fopen(path, "wb");
  if (!f) { pt_set_error_("cannot open %s for writing", path); return PT_ERR_IO; }
  size_t w = fwrite(out.data(), 1, out.size(), f);
  fclose(f);
  if (w != out.size()) { pt_set_error_("short write to %s", path); return PT_ERR_IO; }
  return PT_OK;
}

int write_bmp(const char* path, int W, int H, const unsigned char* rgb8) {
  const int pad = (4 - (3 * W) % 4) % 4;
  const uint32_t data = (uint32_t)((3 * W + pad) * H);
  std::vector<unsigned char> out;
  auto p16 = [&](uint16_t v) { out.push_back((unsigned char)v); out.push_back((unsigned char)(v >> 8)); };
  auto p32 = [&](uint32_t v) { p16((uint16_t)v); p16((uint16_t)(v >> 16)); };
  out.push_back('B'); out.push_back('M');
  p32(14 + 40 + data); p16(0); p16(0); p32(14 + 40);
  p32(40); p32((uint32_t)W); p32((uint32_t)H); p16(1); p16(24); p32(0); p32(0); p32(0); p32(0); p32(0); p32(0);
  for (int y = H - 1; y >= 0; y--) {  // bottom-up, BGR
    for (int x = 0; x < W; x++) {
      const unsigned char* px = rgb8 + 3 * ((size_t)y * W + x);
      out.push_back(px[2]); out.push_back(px[1]); out.push_back(px[0]);
    }
    for (int i = 0; i < pad; i++) out.push_back(0);
  }
  FILE* f = fopen(path, "wb");
  if (!f) { pt_set_error_("cannot open %s for writing", path); return PT_ERR_IO; }
  size_t w = fwrite(out.data(), 1, out.size(), f);
  fclose(f);
  if (w != out.size()) { pt_set_error_("short write to %s", path); return PT_ERR_IO; }
  return PT_OK;
}

// utilityCore::replaceString: first occurrence only
bool replace_first(std::string& s, const std::string& from, const std::string& to) {
  size_t pos = s.find(from);
  if (pos == std::string::npos) return false;
  s.replace(pos, from.length(), to);
  return true;
}

}  // namespace

extern "C" int pt_image_to_rgb8(const float* rgb, int W, int H, uint8_t* rgb8) {
  if (!rgb || !rgb8 || W <= 0 || H <= 0) { pt_set_error_("bad arguments"); return PT_ERR_INVALID; }
  for (int y = 0; y < H; y++) {
    for (int ox = 0; ox < W; ox++) {
      const float* src = rgb + 3 * ((size_t)(W - 1 - ox) + (size_t)y * W);  // main.cpp:120-125
      uint8_t* dst = rgb8 + 3 * ((size_t)y * W + ox);
      dst[0] = to_u8(src[0]); dst[1] = to_u8(src[1]); dst[2] = to_u8(src[2]);
    }
  }
  return PT_OK;
}

extern "C" int pt_save_image(const float* rgb, int W, int H, const char* image_name, int frame, int force_png,
                             char* out_name, int out_name_cap) {
  if (!rgb || !image_name || W <= 0 || H <= 0) { pt_set_error_("bad arguments"); return PT_ERR_INVALID; }
  std::string name = image_name;
  if (force_png && name.size() >= 4 && name.compare(name.size() - 4, 4, ".bmp") == 0)
    name.replace(name.size() - 4, 4, ".png");
  const std::string s = std::to_string(frame);
  replace_first(name, ".bmp", "." + s + ".bmp");
  replace_first(name, ".png", "." + s + ".png");
  bool bmp = false;  // image.cpp:68-80
  const size_t n = name.size();
  if (n >= 4 && name[n - 1] == '\r') bmp = name[n - 4] == 'b' && name[n - 3] == 'm' && name[n - 2] == 'p';
  else if (n >= 3) bmp = name[n - 3] == 'b' && name[n - 2] == 'm' && name[n - 1] == 'p';
  std::vector<unsigned char> px((size_t)W * H * 3);
  int rc = pt_image_to_rgb8(rgb, W, H, px.data());
  if (rc) return rc;
  rc = bmp ? write_bmp(name.c_str(), W, H, px.data()) : write_png(name.c_str(), W, H, px.data());
  if (rc) return rc;
  if (out_name && out_name_cap > 0) {
    strncpy(out_name, name.c_str(), out_name_cap - 1);
    out_name[out_name_cap - 1] = 0;
  }
  return PT_OK;
}
