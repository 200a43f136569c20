/* Magick++.h — stand-in for ImageMagick's C++ API (not installed).  TEST INFRASTRUCTURE ONLY.
 * cell/main.cpp:35-38 builds an Image from the renderer's RGBA8 frame, sets a JPEG quality and writes it. The stand-in
 * writes the same pixels raw: one text line "YVRGBA <width> <height>\n", then width*height*4 bytes, to the name
 * main.cpp chose. A NULL frame (RenderFrame failed) writes only the header with size 0 0. */
#ifndef YV_REF_SHIM_MAGICK_H
#define YV_REF_SHIM_MAGICK_H
#include <stdio.h>
#include <string>
#include <vector>
namespace Magick {
enum StorageType { CharPixel };
class Image {
  unsigned m_w, m_h;
  std::vector<unsigned char> m_px;
public:
  Image(unsigned w, unsigned h, const std::string &map, StorageType, const void *pixels) : m_w(w), m_h(h) {
    if (pixels && map == "RGBA") m_px.assign((const unsigned char *)pixels, (const unsigned char *)pixels + (size_t)w * h * 4);
    else m_w = m_h = 0;
  }
  void quality(unsigned) {}
  void write(const std::string &fn) {
    FILE *f = fopen(fn.c_str(), "wb");
    if (!f) return;
    fprintf(f, "YVRGBA %u %u\n", m_w, m_h);
    if (!m_px.empty()) fwrite(&m_px[0], 1, m_px.size(), f);
    fclose(f);
  }
};
}
#endif
