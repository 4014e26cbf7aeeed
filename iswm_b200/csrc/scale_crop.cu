// scale_crop.cu - the reference's whole TRAIN transform on the device (train.py:355-362):
//   ExtRandomScale((0.5, 2.0)) -> ExtRandomCrop(crop, pad_if_needed=True) -> ExtRandomHorizontalFlip -> ExtToTensor -> ExtNormalize
// over uint8 tiles, bit-identical to Pillow's resampling (scale_math.h restates it). Only the H x W crop window of the
// scaled tile is ever computed: a 0.5x..2x tile of 512^2 is up to 1024^2 pixels of which the crop keeps 512^2.
//   plan kernel   per sample and axis: the 2^-22 fixed-point bilinear coefficient rows (threads in parallel) and the
//                 nearest-neighbour index table (ONE thread: Pillow's running double sum is sequential by definition)
//   image kernel  one thread per output pixel: <= kmax x kmax taps of the source tile (horizontal pass rounded to uint8
//                 per source row, then the vertical pass, exactly the two-pass order), zero fill in the padding,
//                 mirror, /255, -mean, /std -> fp32 NCHW
//   label kernel  one thread per output pixel through the two index tables
// HBM-bound on paper (3 B/px read through L1/L2, 12 + 1 B/px written), in practice bound by the byte loads of the taps.
#include "ew_common.cuh"
#include "scale_math.h"

namespace iswm {
namespace {

using scale::Geom;

// a geometry record comes from device memory the library cannot check on the host: clamp it to the table space
__device__ __forceinline__ Geom load_geom(const Geom* geom, int b, int tab_w, int tab_h) {
  Geom g = geom[b];
  g.sw = max(0, min(g.sw, tab_w));
  g.sh = max(0, min(g.sh, tab_h));
  return g;
}

__global__ void __launch_bounds__(kT)
scale_plan_kernel(const Geom* __restrict__ geom, int Hs, int Ws, int kmax, int tab_w, int tab_h, int32_t* __restrict__ tables) {
  pdl_wait();
  pdl_launch();
  const int b = blockIdx.x, axis = blockIdx.y;            // axis 0: columns (Ws -> sw), 1: rows (Hs -> sh)
  const Geom g = load_geom(geom, b, tab_w, tab_h);
  int32_t* base = tables + (int64_t)b * scale::table_words(tab_w, tab_h, kmax);
  const int in_size = axis == 0 ? Ws : Hs, out_size = axis == 0 ? g.sw : g.sh;
  int32_t* coef = base + (axis == 0 ? scale::off_hx(tab_w, tab_h, kmax) : scale::off_vy(tab_w, tab_h, kmax));
  int32_t* nidx = base + (axis == 0 ? scale::off_xn(tab_w, tab_h, kmax) : scale::off_yn(tab_w, tab_h, kmax));
  if (threadIdx.x == kT - 1) scale::nearest_table(in_size, out_size, nidx);
  for (int xx = threadIdx.x; xx < out_size; xx += kT) scale::bilinear_entry(in_size, out_size, xx, kmax, coef + (int64_t)xx * (2 + kmax));
}

template <int C>
__global__ void __launch_bounds__(kT)
scale_crop_image_kernel(const uint8_t* __restrict__ src, int Hs, int Ws, const Geom* __restrict__ geom, int kmax, int tab_w, int tab_h,
                        const int32_t* __restrict__ tables, float4 mean, float4 stdv, int H, int W, float* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  const int y = blockIdx.x % H, b = blockIdx.x / H;
  const Geom g = load_geom(geom, b, tab_w, tab_h);
  const int32_t* base = tables + (int64_t)b * scale::table_words(tab_w, tab_h, kmax);
  const int32_t* hx = base + scale::off_hx(tab_w, tab_h, kmax);
  const int32_t* vy = base + scale::off_vy(tab_w, tab_h, kmax);
  const uint8_t* tile = src + (int64_t)b * Hs * Ws * C;
  const float mu[4] = {mean.x, mean.y, mean.z, mean.w}, sd[4] = {stdv.x, stdv.y, stdv.z, stdv.w};
  const int64_t plane = (int64_t)H * W;
  float* obase = out + (int64_t)b * C * plane + (int64_t)y * W;
  for (int x = blockIdx.y * kT + threadIdx.x; x < W; x += gridDim.y * kT) {
    int Y, X, v[C];
    if (scale::crop_to_scaled(g, W, y, x, Y, X)) {
      scale::bilinear_pixel<C>(tile, Ws, hx + (int64_t)X * (2 + kmax), vy + (int64_t)Y * (2 + kmax), v);
    } else {
#pragma unroll
      for (int c = 0; c < C; c++) v[c] = 0;                 // F.pad(..., fill=0) before ToTensor
    }
#pragma unroll
    for (int c = 0; c < C; c++) obase[(int64_t)c * plane + x] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v[c], 255.f), mu[c]), sd[c]);
  }
}

__global__ void __launch_bounds__(kT)
scale_crop_label_kernel(const uint8_t* __restrict__ src, int Hs, int Ws, const Geom* __restrict__ geom, int kmax, int tab_w, int tab_h,
                        const int32_t* __restrict__ tables, int H, int W, uint8_t* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  const int y = blockIdx.x % H, b = blockIdx.x / H;
  const Geom g = load_geom(geom, b, tab_w, tab_h);
  const int32_t* base = tables + (int64_t)b * scale::table_words(tab_w, tab_h, kmax);
  const int32_t* xn = base + scale::off_xn(tab_w, tab_h, kmax);
  const int32_t* yn = base + scale::off_yn(tab_w, tab_h, kmax);
  const uint8_t* tile = src + (int64_t)b * Hs * Ws;
  uint8_t* orow = out + ((int64_t)b * H + y) * W;
  for (int x = blockIdx.y * kT + threadIdx.x; x < W; x += gridDim.y * kT) {
    int Y, X;
    uint8_t v = 0;                                          // label padding is 0 as well (F.pad default fill)
    if (scale::crop_to_scaled(g, W, y, x, Y, X)) {
      const int ys = yn[Y], xs = xn[X];
      if (ys >= 0 && xs >= 0) v = tile[(int64_t)ys * Ws + xs];
    }
    orow[x] = v;
  }
}

}  // namespace
}  // namespace iswm

using namespace iswm;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int64_t iswm_random_scale_table_words(int tab_w, int tab_h, int kmax) { return scale::table_words(tab_w, tab_h, kmax); }

extern "C" int iswm_random_scale_crop(const uint8_t* d_img, const uint8_t* d_lbl, int B, int Hs, int Ws, int C, const int32_t* d_geom,
                                      int kmax, int tab_w, int tab_h, int32_t* d_tables, const float* mean, const float* stdv,
                                      int H, int W, float* d_out, uint8_t* d_lbl_out, void* stream) {
  static_assert(sizeof(Geom) == 32, "iswm_random_scale_crop geometry record");
  ISWM_REQUIRE(d_img && d_geom && d_tables && d_out && mean && stdv, "random_scale_crop: null");
  ISWM_REQUIRE((d_lbl == nullptr) == (d_lbl_out == nullptr), "random_scale_crop: label input and output go together");
  ISWM_REQUIRE(C >= 1 && C <= 4, "random_scale_crop: C=%d (1..4 channels)", C);
  ISWM_REQUIRE(kmax >= 3 && kmax <= scale::kMaxTaps, "random_scale_crop: kmax=%d (3..%d)", kmax, scale::kMaxTaps);
  ISWM_REQUIRE(B >= 0 && Hs >= 1 && Ws >= 1 && H >= 1 && W >= 1 && tab_w >= 1 && tab_h >= 1, "random_scale_crop: bad sizes");
  ISWM_REQUIRE(Hs < (1 << 24) && Ws < (1 << 24), "random_scale_crop: tile larger than Pillow's float box");
  ISWM_REQUIRE((int64_t)B * H < (1ll << 31), "random_scale_crop: too many rows");
  if (B == 0) return 0;
  const Geom* geom = reinterpret_cast<const Geom*>(d_geom);
  launch_k(scale_plan_kernel, dim3((unsigned)B, 2), dim3(kT), 0, ST(stream), geom, Hs, Ws, kmax, tab_w, tab_h, d_tables);
  if (int rc = check_launch("random_scale_crop(plan)")) return rc;
  float mu[4] = {0, 0, 0, 0}, sd[4] = {1, 1, 1, 1};
  for (int c = 0; c < C; c++) { mu[c] = mean[c]; sd[c] = stdv[c]; }
  const float4 m4 = make_float4(mu[0], mu[1], mu[2], mu[3]), s4 = make_float4(sd[0], sd[1], sd[2], sd[3]);
  dim3 grid((unsigned)(B * H), (unsigned)std::max(1, std::min(8, (W + kT - 1) / kT)));
#define ISWM_SCALE_LAUNCH(CC) \
  launch_k(scale_crop_image_kernel<CC>, grid, dim3(kT), 0, ST(stream), d_img, Hs, Ws, geom, kmax, tab_w, tab_h, (const int32_t*)d_tables, m4, s4, H, W, d_out)
  switch (C) {
    case 1: ISWM_SCALE_LAUNCH(1); break;
    case 2: ISWM_SCALE_LAUNCH(2); break;
    case 3: ISWM_SCALE_LAUNCH(3); break;
    default: ISWM_SCALE_LAUNCH(4); break;
  }
#undef ISWM_SCALE_LAUNCH
  if (int rc = check_launch("random_scale_crop(image)")) return rc;
  if (d_lbl) {
    launch_k(scale_crop_label_kernel, grid, dim3(kT), 0, ST(stream), d_lbl, Hs, Ws, geom, kmax, tab_w, tab_h, (const int32_t*)d_tables, H, W, d_lbl_out);
    if (int rc = check_launch("random_scale_crop(label)")) return rc;
  }
  return 0;
}
