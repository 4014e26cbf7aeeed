// tail_fused.cu - the train-mode tail of the network without its full-resolution tensors (SURVEY kernels K11 + K12 + K13):
//   network/utils.py:22   F.interpolate(low-res logits, size=input, bilinear, align_corners=False)
//   train.py:1046         criterion = CrossEntropyLoss(weight, ignore_index=255, reduction='mean') on the upsampled logits
//   train.py:1048         loss.backward() down to the classifier's output: CE gradient, adjoint of the upsample, bias gradient
// The unfused chain (iswm_logits_up_fwd -> iswm_class_hist -> iswm_wce_fwd_bwd -> iswm_logits_up_bwd) writes 33.5 MB of fp32
// logits, reads them back, writes 33.5 MB of gradient, reads it back and reads the labels twice (cfg2: 16 x 512^2, 2 classes);
// here a block owns an 8 x 8 tile of the LOW-RES grid: it interpolates the 36 x 36 full-resolution pixels that touch the
// tile from the 10 x 10 low-res logits in shared memory (the same fma chain as iswm_logits_up_fwd: identical logits), forms
// the per-pixel CE (the same arithmetic as wce2_kernel) and folds the gradient back onto the tile with a separable,
// fixed-order adjoint - no atomics on the gradient, bit-reproducible. Loss numerator and class histogram come from the 32 x 32
// pixels the tile owns. HBM traffic: labels once (B*H*W*sizeof(label)) + 2 * B*h*w*8 bytes.
// The 1 / sum_c w_c n_c normaliser needs the histogram of the WHOLE (data-parallel: global) batch, so the tile gradient is
// stored unscaled (fp32) and iswm_tail_bwd applies it while converting to the bf16 operand of the classifier's backward and
// summing the classifier bias gradient (fixed-order: per-block partials, the last block adds them up).
// Two classes (the reference's binary task) and an exact x4 upsample (output stride 4 decoder); anything else runs the unfused chain.
#include "common.cuh"
#include <algorithm>

namespace iswm {
namespace {

constexpr int kTT = 256;
constexpr int TL = 8;              // low-res tile edge
constexpr int TF = 4 * TL + 4;     // full-resolution rows / columns whose bilinear footprint touches the tile

__device__ __forceinline__ double tail_denominator(const float* w, const int64_t* hist, int ignore_index) {
  double D = 0.0;
  for (int c = 0; c < 2; c++) {
    if (c == ignore_index) continue;
    D += (double)(w ? w[c] : 1.0f) * (double)hist[c];
  }
  return D;
}

template <typename YT>
__global__ void __launch_bounds__(kTT)
tail_fused_kernel(const float* __restrict__ lo, int h, int w, const YT* __restrict__ labels, int H, int W,
                  const float* __restrict__ weight, int ignore_index, float* __restrict__ dlo_acc,
                  unsigned long long* __restrict__ hist, double* __restrict__ loss_num) {
  pdl_wait();
  pdl_launch();
  __shared__ float2 s_lo[(TL + 2) * (TL + 2)];     // (class 0, class 1) of the low-res logits around the tile
  __shared__ float s_g[TF][TF + 1];
  __shared__ float s_t[TL][TF + 1];
  __shared__ int s_i0[2][TF], s_i1[2][TF];         // low-res tap indices of every full-resolution row / column (-1000 = outside the image)
  __shared__ int s_o0[2][TF], s_o1[2][TF];         // the same as element offsets into s_lo (rows pre-multiplied by the tile pitch)
  __shared__ float s_l1[2][TF];
  __shared__ float s_red[kTT / 32];
  __shared__ unsigned s_cnt[kTT / 32][2];
  const int b = blockIdx.z, I0 = blockIdx.y * TL, J0 = blockIdx.x * TL;
  const int ybase = 4 * I0 - 2, xbase = 4 * J0 - 2;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  // every label this thread will need, requested up front: NIT loads in flight instead of a chain of NIT cold misses
  constexpr int NIT = (TF * TF + kTT - 1) / kTT;
  YT yv[NIT];
#pragma unroll
  for (int k = 0; k < NIT; k++) {
    const int t = threadIdx.x + k * kTT;
    const int py = t / TF, px = t - py * TF;
    const int y = ybase + py, x = xbase + px;
    yv[k] = (t < TF * TF && y >= 0 && y < H && x >= 0 && x < W) ? labels[((int64_t)b * H + y) * W + x] : (YT)0;
  }
  for (int t = threadIdx.x; t < (TL + 2) * (TL + 2); t += kTT) {
    const int r = t / (TL + 2), c = t % (TL + 2);
    const int ii = min(max(I0 - 1 + r, 0), h - 1), jj = min(max(J0 - 1 + c, 0), w - 1);
    s_lo[t] = *reinterpret_cast<const float2*>(lo + (((int64_t)b * h + ii) * w + jj) * 2);
  }
  if (threadIdx.x < 2 * TF) {
    const int axis = threadIdx.x / TF, k = threadIdx.x % TF;
    const int o = (axis == 0 ? ybase : xbase) + k;
    const int n_out = axis == 0 ? H : W, n_in = axis == 0 ? h : w;
    int i0 = -1000, i1 = -1000;
    float l1 = 0.f;
    if (o >= 0 && o < n_out) bil_src(o, axis == 0 ? sh : sw, n_in, i0, i1, l1);
    s_i0[axis][k] = i0;
    s_i1[axis][k] = i1;
    s_l1[axis][k] = l1;
    const int org = (axis == 0 ? I0 : J0) - 1, pitch = axis == 0 ? (TL + 2) : 1;
    s_o0[axis][k] = i0 < 0 ? -1 : (i0 - org) * pitch;
    s_o1[axis][k] = i0 < 0 ? -1 : (i1 - org) * pitch;
  }
  __syncthreads();
  const float w0 = weight ? weight[0] : 1.0f, w1 = weight ? weight[1] : 1.0f;
  float acc = 0.f;
  unsigned c0 = 0, c1 = 0;
#pragma unroll
  for (int k = 0; k < NIT; k++) {
    const int t = threadIdx.x + k * kTT;
    if (t >= TF * TF) break;
    const int py = t / TF, px = t - py * TF;
    const int ry0 = s_o0[0][py], cx0 = s_o0[1][px];
    float g = 0.f;
    if ((ry0 | cx0) >= 0) {                                      // inside the image on both axes
      const int ry1 = s_o1[0][py], cx1 = s_o1[1][px];
      const float ly = s_l1[0][py], lx = s_l1[1][px];
      const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
      const float2 v00 = s_lo[ry0 + cx0], v01 = s_lo[ry0 + cx1], v10 = s_lo[ry1 + cx0], v11 = s_lo[ry1 + cx1];
      const float a0 = bil_mix(w00, v00.x, w01, v01.x, w10, v10.x, w11, v11.x);
      const float a1 = bil_mix(w00, v00.y, w01, v01.y, w10, v10.y, w11, v11.y);
      const long long yy = (long long)yv[k];
      const bool valid = (yy == 0 || yy == 1) && yy != ignore_index;
      // the arithmetic of wce2_kernel: d = x_other - x_target, nll = softplus(d), p_other = sigmoid(d)
      const float d = (yy == 1) ? (a0 - a1) : (a1 - a0);
      const float e = expf(-fabsf(d));
      const float nll = fmaxf(d, 0.0f) + log1pf(e);
      const float inv = 1.0f / (1.0f + e);
      const float p_other = (d >= 0.0f) ? inv : e * inv;
      const float wv = (yy == 1) ? w1 : w0;
      const float gt = valid ? wv * p_other : 0.0f;
      g = (yy == 1) ? gt : -gt;                                 // d loss / d logit of class 0 (class 1: the negative), unnormalised
      if (py >= 2 && py < TF - 2 && px >= 2 && px < TF - 2) {   // the 32 x 32 pixels this tile owns
        if (valid) acc += wv * nll;
        c0 += (yy == 0);
        c1 += (yy == 1);
      }
    }
    s_g[py][px] = g;
  }
  __syncthreads();
  // adjoint of the upsample, rows: low row I0 + r collects the <= 8 full-resolution rows whose taps name it, in row order
  for (int t = threadIdx.x; t < TL * TF; t += kTT) {
    const int r = t / TF, px = t % TF, ii = I0 + r;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int py = 4 * r + k;
      float wy = 0.f;
      if (s_i0[0][py] == ii) wy += 1.f - s_l1[0][py];
      if (s_i1[0][py] == ii) wy += s_l1[0][py];
      v = fmaf(wy, s_g[py][px], v);
    }
    s_t[r][px] = v;
  }
  __syncthreads();
  if (threadIdx.x < TL * TL) {
    const int r = threadIdx.x / TL, c = threadIdx.x % TL, jj = J0 + c;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int px = 4 * c + k;
      float wx = 0.f;
      if (s_i0[1][px] == jj) wx += 1.f - s_l1[1][px];
      if (s_i1[1][px] == jj) wx += s_l1[1][px];
      v = fmaf(wx, s_t[r][px], v);
    }
    if (I0 + r < h && jj < w) *reinterpret_cast<float2*>(dlo_acc + (((int64_t)b * h + I0 + r) * w + jj) * 2) = make_float2(v, -v);
  }
  // loss numerator (fp64 accumulator) and class counts of the owned pixels
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_red[warp] = acc; s_cnt[warp][0] = c0; s_cnt[warp][1] = c1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    unsigned long long n0 = 0, n1 = 0;
    for (int i = 0; i < kTT / 32; i++) { t += (double)s_red[i]; n0 += s_cnt[i][0]; n1 += s_cnt[i][1]; }
    if (t != 0.0) atomicAdd(loss_num, t);
    if (n0) atomicAdd(hist + 0, n0);
    if (n1) atomicAdd(hist + 1, n1);
  }
}

__global__ void tail_loss_kernel(const double* loss_num, const float* weight, const int64_t* hist, int ignore_index, float* loss) {
  pdl_wait();
  pdl_launch();
  const double D = tail_denominator(weight, hist, ignore_index);
  *loss = (float)(*loss_num / D);                               // 0 / 0 -> nan, as torch for an all-ignored batch
}

// dlo = bf16(dlo_acc * g / D) padded to dx_ld channels (zeros), bias_grad[0] += sum, bias_grad[1] -= sum (fixed order)
__global__ void __launch_bounds__(kTT)
tail_bwd_kernel(const float* __restrict__ dlo_acc, int64_t n_px, const float* __restrict__ weight, const int64_t* __restrict__ hist,
                int ignore_index, const float* __restrict__ gscale, __nv_bfloat16* __restrict__ dlo, int dx_ld,
                float* __restrict__ bias_grad, double* __restrict__ partials, unsigned* __restrict__ counter) {
  pdl_wait();
  pdl_launch();
  __shared__ float s_scale;
  __shared__ double s_red[kTT / 32];
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    const double D = tail_denominator(weight, hist, ignore_index);
    float s = D != 0.0 ? (float)(1.0 / D) : 0.0f;               // an all-ignored batch has a zero gradient (and a nan loss)
    if (gscale) s *= *gscale;
    s_scale = s;
  }
  __syncthreads();
  const float s = s_scale;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kTT + threadIdx.x; i < n_px; i += (int64_t)gridDim.x * kTT) {
    const float2 v = *reinterpret_cast<const float2*>(dlo_acc + i * 2);
    const float g0 = v.x * s, g1 = v.y * s;
    acc += (double)g0;
    __nv_bfloat16* o = dlo + i * dx_ld;
    if (dx_ld == 8) {
      uint4 r;
      r.x = pack_bf16x2(g0, g1);
      r.y = r.z = r.w = 0u;
      *reinterpret_cast<uint4*>(o) = r;
    } else {
      o[0] = __float2bfloat16_rn(g0);
      o[1] = __float2bfloat16_rn(g1);
      for (int c = 2; c < dx_ld; c++) o[c] = __float2bfloat16_rn(0.f);
    }
  }
  if (bias_grad == nullptr) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kTT / 32; i++) t += s_red[i];
    partials[blockIdx.x] = t;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {                                                 // the last block to finish adds the partials up, in a fixed order
    __threadfence();
    double t = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += kTT) t += *reinterpret_cast<volatile double*>(partials + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();                                            // s_red is reused
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double sum = 0.0;
      for (int i = 0; i < kTT / 32; i++) sum += s_red[i];
      bias_grad[0] += (float)sum;
      bias_grad[1] -= (float)sum;
      *counter = 0;                                             // ready for the next launch (graph replays included)
    }
  }
}

}  // namespace
}  // namespace iswm

using namespace iswm;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int iswm_tail_fwd(const float* d_lo, int B, int h, int w, const void* d_labels, int label_dtype, int H, int W, const float* d_weight,
                             int ignore_index, float* d_dlo_acc, int64_t* d_hist, double* d_loss_num, void* stream) {
  ISWM_REQUIRE(d_lo && d_labels && d_dlo_acc && d_hist && d_loss_num, "tail_fwd: null");
  ISWM_REQUIRE(B >= 1 && h >= 1 && w >= 1 && H == 4 * h && W == 4 * w, "tail_fwd: the fused tail is the exact x4 upsample (%dx%d -> %dx%d)", h, w, H, W);
  ISWM_REQUIRE(B <= 65535 && (int64_t)B * H * W < (1ll << 40), "tail_fwd: batch too large");
  dim3 grid((unsigned)((w + TL - 1) / TL), (unsigned)((h + TL - 1) / TL), (unsigned)B);
  auto* hist = reinterpret_cast<unsigned long long*>(d_hist);
  switch (label_dtype) {
    case ISWM_U8: launch_k(tail_fused_kernel<uint8_t>, grid, dim3(kTT), 0, ST(stream), d_lo, h, w, (const uint8_t*)d_labels, H, W, d_weight, ignore_index, d_dlo_acc, hist, d_loss_num); break;
    case ISWM_I32: launch_k(tail_fused_kernel<int32_t>, grid, dim3(kTT), 0, ST(stream), d_lo, h, w, (const int32_t*)d_labels, H, W, d_weight, ignore_index, d_dlo_acc, hist, d_loss_num); break;
    case ISWM_I64: launch_k(tail_fused_kernel<int64_t>, grid, dim3(kTT), 0, ST(stream), d_lo, h, w, (const int64_t*)d_labels, H, W, d_weight, ignore_index, d_dlo_acc, hist, d_loss_num); break;
    default: set_error("tail_fwd: label dtype %d", label_dtype); return 2;
  }
  return check_launch("tail_fwd");
}

extern "C" int iswm_tail_loss(const double* d_loss_num, const float* d_weight, const int64_t* d_hist, int ignore_index, float* d_loss, void* stream) {
  ISWM_REQUIRE(d_loss_num && d_hist && d_loss, "tail_loss: null");
  launch_k(tail_loss_kernel, dim3(1), dim3(1), 0, ST(stream), d_loss_num, d_weight, d_hist, ignore_index, d_loss);
  return check_launch("tail_loss");
}

extern "C" int iswm_tail_bwd(const float* d_dlo_acc, int B, int h, int w, const float* d_weight, const int64_t* d_hist, int ignore_index,
                             const float* d_gscale, void* d_dlo, int dx_ld, float* d_bias_grad, void* d_scratch, void* stream) {
  ISWM_REQUIRE(d_dlo_acc && d_hist && d_dlo && d_scratch, "tail_bwd: null");
  ISWM_REQUIRE(B >= 1 && h >= 1 && w >= 1 && dx_ld >= 2, "tail_bwd: bad sizes");
  const int64_t n_px = (int64_t)B * h * w;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_px + kTT - 1) / kTT, 1024));
  double* partials = static_cast<double*>(d_scratch);
  unsigned* counter = reinterpret_cast<unsigned*>(partials + 1024);
  launch_k(tail_bwd_kernel, dim3((unsigned)grid), dim3(kTT), 0, ST(stream), d_dlo_acc, n_px, d_weight, d_hist, ignore_index, d_gscale,
           static_cast<__nv_bfloat16*>(d_dlo), dx_ld, d_bias_grad, partials, counter);
  return check_launch("tail_bwd");
}
