// Host emulation of scale_crop.cu for the CPU tests: the SAME per-entry / per-pixel functions (iswm_b200/csrc/scale_math.h)
// driven by plain loops in place of the CUDA grid. Built by tests/test_scale_oracle.py with
//   g++ -O1 -ffp-contract=off -shared -fPIC
// TEST INFRASTRUCTURE: nothing in iswm_b200/ loads this; the product path is the CUDA library.
#include <stdint.h>
#include <vector>
#include "../../iswm_b200/csrc/scale_math.h"

using namespace iswm::scale;

template <int C>
static void image_rows(const uint8_t* tile, int Ws, const Geom& g, int kmax, const int32_t* hx, const int32_t* vy, const float* mean,
                       const float* stdv, int H, int W, float* out) {
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      int Y, X, v[C];
      if (crop_to_scaled(g, W, y, x, Y, X)) bilinear_pixel<C>(tile, Ws, hx + (int64_t)X * (2 + kmax), vy + (int64_t)Y * (2 + kmax), v);
      else for (int c = 0; c < C; c++) v[c] = 0;
      for (int c = 0; c < C; c++) {
        volatile float t = (float)v[c] / 255.f;
        volatile float u = t - mean[c];
        out[((int64_t)c * H + y) * W + x] = u / stdv[c];
      }
    }
}

extern "C" int emul_random_scale_crop(const uint8_t* img, const uint8_t* lbl, int B, int Hs, int Ws, int C, const int32_t* geom_words,
                                      int kmax, int tab_w, int tab_h, const float* mean, const float* stdv, int H, int W, float* out,
                                      uint8_t* lbl_out) {
  if (C != 3 && C != 1) return 2;
  const Geom* geom = reinterpret_cast<const Geom*>(geom_words);
  std::vector<int32_t> tab((size_t)table_words(tab_w, tab_h, kmax));
  for (int b = 0; b < B; b++) {
    const Geom g = geom[b];
    if (g.sw > tab_w || g.sh > tab_h || ksize_for(Ws, g.sw) > kmax || ksize_for(Hs, g.sh) > kmax) return 3;
    int32_t* hx = tab.data() + off_hx(tab_w, tab_h, kmax);
    int32_t* vy = tab.data() + off_vy(tab_w, tab_h, kmax);
    int32_t* xn = tab.data() + off_xn(tab_w, tab_h, kmax);
    int32_t* yn = tab.data() + off_yn(tab_w, tab_h, kmax);
    for (int xx = 0; xx < g.sw; xx++) bilinear_entry(Ws, g.sw, xx, kmax, hx + (int64_t)xx * (2 + kmax));
    for (int yy = 0; yy < g.sh; yy++) bilinear_entry(Hs, g.sh, yy, kmax, vy + (int64_t)yy * (2 + kmax));
    nearest_table(Ws, g.sw, xn);
    nearest_table(Hs, g.sh, yn);
    const uint8_t* tile = img + (int64_t)b * Hs * Ws * C;
    float* o = out + (int64_t)b * C * H * W;
    if (C == 3) image_rows<3>(tile, Ws, g, kmax, hx, vy, mean, stdv, H, W, o);
    else image_rows<1>(tile, Ws, g, kmax, hx, vy, mean, stdv, H, W, o);
    if (lbl) {
      const uint8_t* lt = lbl + (int64_t)b * Hs * Ws;
      for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
          int Y, X;
          uint8_t v = 0;
          if (crop_to_scaled(g, W, y, x, Y, X)) {
            const int ys = yn[Y], xs = xn[X];
            if (ys >= 0 && xs >= 0) v = lt[(int64_t)ys * Ws + xs];
          }
          lbl_out[((int64_t)b * H + y) * W + x] = v;
        }
    }
  }
  return 0;
}

extern "C" int emul_ksize_for(int in_size, int out_size) { return ksize_for(in_size, out_size); }
