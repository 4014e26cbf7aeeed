// scale_math.h - the arithmetic of ExtRandomScale + ExtRandomCrop(pad_if_needed) (utils/ext_transforms.py:94-111,
// :366-393) per table entry and per output pixel, written once for the device kernels (scale_crop.cu) and for the
// host emulation the CPU tests compile with g++ (tests/host_emul/scale_emul.cpp): the index arithmetic of the
// kernels is checked on the build container, without a GPU, against the oracle and the reference-generated goldens.
//
// The resampling arithmetic is Pillow's (reached by the reference through torchvision F.resize -> Image.resize):
//   bilinear, images : Resample.c precompute_coeffs / normalize_coeffs_8bpc (double weights -> 2^-22 fixed point),
//                      ImagingResampleHorizontal_8bpc into a uint8 intermediate, then ImagingResampleVertical_8bpc
//   nearest, labels  : Geometry.c ImagingScaleAffine (source index = (int) of a RUNNING double sum)
// Every double operation is a separately rounded IEEE operation (no fused multiply-add): on the device the explicit
// _rn intrinsics keep nvcc from contracting them.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ISWM_HD __host__ __device__ __forceinline__
#else
#define ISWM_HD inline
#endif

namespace iswm {
namespace scale {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kHalf = 1 << (kPrecisionBits - 1);
constexpr int kMaxTaps = 16;           // ceil(1 / min scale) * 2 + 1 <= 16: scales down to 1/7

#if defined(__CUDA_ARCH__)
ISWM_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
ISWM_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
ISWM_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
ISWM_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
#else
ISWM_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
ISWM_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
ISWM_HD double dsub(double a, double b) { volatile double r = a - b; return r; }
ISWM_HD double ddiv(double a, double b) { volatile double r = a / b; return r; }
#endif

// geometry of one sample (int32[8], written by the host that drew the random numbers)
struct Geom {
  int32_t sh, sw;      // size after ExtRandomScale: (int(Hs * scale), int(Ws * scale))
  int32_t pad;         // zero padding ExtRandomCrop(pad_if_needed) adds on EVERY side (0 when the scaled tile is large enough)
  int32_t y0, x0;      // crop origin in the padded tile
  int32_t flip;        // ExtRandomHorizontalFlip decision
  int32_t r0, r1;
};

// table layout of one sample, int32 words: [tab_w][2 + kmax] | [tab_h][2 + kmax] | [tab_w] | [tab_h]
ISWM_HD int64_t table_words(int tab_w, int tab_h, int kmax) { return (int64_t)(tab_w + tab_h) * (3 + kmax); }
ISWM_HD int64_t off_hx(int, int, int) { return 0; }
ISWM_HD int64_t off_vy(int tab_w, int, int kmax) { return (int64_t)tab_w * (2 + kmax); }
ISWM_HD int64_t off_xn(int tab_w, int tab_h, int kmax) { return (int64_t)(tab_w + tab_h) * (2 + kmax); }
ISWM_HD int64_t off_yn(int tab_w, int tab_h, int kmax) { return (int64_t)(tab_w + tab_h) * (2 + kmax) + tab_w; }

ISWM_HD int ksize_for(int in_size, int out_size) {
  double scale = ddiv((double)((float)in_size - 0.0f), (double)out_size);
  double fs = scale < 1.0 ? 1.0 : scale;
  int c = (int)fs;
  if ((double)c < fs) c++;             // ceil(support), support = 1.0 * filterscale
  return c * 2 + 1;
}

// one row of precompute_coeffs + normalize_coeffs_8bpc: entry = (xmin, count, k[0..kmax))
ISWM_HD void bilinear_entry(int in_size, int out_size, int xx, int kmax, int32_t* entry) {
  const double scale = ddiv((double)((float)in_size - 0.0f), (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = dmul(1.0, filterscale);
  const double center = dadd(0.0, dmul(dadd((double)xx, 0.5), scale));
  const double ss = ddiv(1.0, filterscale);
  int xmin = (int)dadd(dsub(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)dadd(dadd(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > kmax) xmax = kmax;        // cannot happen when kmax >= ksize_for(); keeps the writes inside the entry
  double w[kMaxTaps];
  double ww = 0.0;
  for (int x = 0; x < xmax; x++) {
    double t = dmul(dadd(dsub((double)(x + xmin), center), 0.5), ss);
    if (t < 0.0) t = -t;
    w[x] = t < 1.0 ? dsub(1.0, t) : 0.0;
    ww = dadd(ww, w[x]);
  }
  entry[0] = xmin;
  entry[1] = xmax;
  for (int x = 0; x < kmax; x++) {
    int32_t k = 0;
    if (x < xmax) {
      double v = ww != 0.0 ? ddiv(w[x], ww) : w[x];
      k = v < 0 ? (int32_t)dadd(-0.5, dmul(v, (double)(1 << kPrecisionBits))) : (int32_t)dadd(0.5, dmul(v, (double)(1 << kPrecisionBits)));
    }
    entry[2 + x] = k;
  }
}

// ImagingScaleAffine's index table: sequential by construction
ISWM_HD void nearest_table(int in_size, int out_size, int32_t* tab) {
  const double a = ddiv((double)((float)in_size - 0.0f), (double)out_size);
  double xo = dadd(0.0, dmul(a, 0.5));
  for (int x = 0; x < out_size; x++) {
    int xin = xo < 0.0 ? -1 : (int)xo;
    tab[x] = (xin >= 0 && xin < in_size) ? xin : -1;
    xo = dadd(xo, a);
  }
}

ISWM_HD int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// position of output pixel (y, x) of the H x W crop in the SCALED tile; false = it lies in the zero padding
ISWM_HD bool crop_to_scaled(const Geom& g, int W, int y, int x, int& Y, int& X) {
  const int xf = g.flip ? (W - 1 - x) : x;
  Y = g.y0 + y - g.pad;
  X = g.x0 + xf - g.pad;
  return Y >= 0 && Y < g.sh && X >= 0 && X < g.sw;
}

// resampled uint8 value of channels 0..C-1 at (Y, X) of the scaled tile; src = this sample's [Hs][Ws][C] tile
template <int C>
ISWM_HD void bilinear_pixel(const uint8_t* src, int Ws, const int32_t* hx_entry, const int32_t* vy_entry, int out[C]) {
  const int xmin = hx_entry[0], xcnt = hx_entry[1], ymin = vy_entry[0], ycnt = vy_entry[1];
  int acc_v[C];
  for (int c = 0; c < C; c++) acc_v[c] = kHalf;
  for (int ky = 0; ky < ycnt; ky++) {
    const uint8_t* row = src + ((int64_t)(ymin + ky) * Ws + xmin) * C;
    int acc_h[C];
    for (int c = 0; c < C; c++) acc_h[c] = kHalf;
    for (int kx = 0; kx < xcnt; kx++) {
      const int k = hx_entry[2 + kx];
      for (int c = 0; c < C; c++) acc_h[c] += (int)row[kx * C + c] * k;
    }
    const int kv = vy_entry[2 + ky];
    for (int c = 0; c < C; c++) acc_v[c] += clip8(acc_h[c]) * kv;
  }
  for (int c = 0; c < C; c++) out[c] = clip8(acc_v[c]);
}

}  // namespace scale
}  // namespace iswm
