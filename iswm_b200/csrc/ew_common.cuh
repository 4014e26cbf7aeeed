// ew_common.cuh - thread layout and 16-byte row helpers shared by the HBM-bound BatchNorm kernels (elementwise.cu, bn_dual.cu).
#pragma once
#include "common.cuh"
#include <algorithm>

namespace iswm {

constexpr int kT = 256;

struct F8 { float v[8]; };

__device__ __forceinline__ F8 load8(const __nv_bfloat16* p) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  F8 o;
  unpack_bf16x2(r.x, o.v[0], o.v[1]);
  unpack_bf16x2(r.y, o.v[2], o.v[3]);
  unpack_bf16x2(r.z, o.v[4], o.v[5]);
  unpack_bf16x2(r.w, o.v[6], o.v[7]);
  return o;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const F8& f) {
  uint4 r;
  r.x = pack_bf16x2(f.v[0], f.v[1]);
  r.y = pack_bf16x2(f.v[2], f.v[3]);
  r.z = pack_bf16x2(f.v[4], f.v[5]);
  r.w = pack_bf16x2(f.v[6], f.v[7]);
  *reinterpret_cast<uint4*>(p) = r;
}

__device__ __forceinline__ uint4 load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ F8 unpack8(const uint4& r) {
  F8 o;
  unpack_bf16x2(r.x, o.v[0], o.v[1]);
  unpack_bf16x2(r.y, o.v[2], o.v[3]);
  unpack_bf16x2(r.z, o.v[4], o.v[5]);
  unpack_bf16x2(r.w, o.v[6], o.v[7]);
  return o;
}

struct RowWalk {
  int64_t first;      // first row of this thread
  int n;              // rows this thread owns
};
__device__ __forceinline__ RowWalk row_walk(int64_t M, int rows_per_block, int ty, int ny) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(r0 + (int64_t)rows_per_block, M);
  RowWalk w;
  w.first = r0 + ty;
  w.n = (w.first < r1) ? (int)((r1 - w.first + ny - 1) / ny) : 0;
  return w;
}

static inline void bn_red_shape(int C, int& nx, int& ny) {
  const int nvec = C / 8;
  nx = std::min(nvec, kT);
  ny = std::max(1, kT / nx);
}
// rows per block: every thread walks >= min_rows rows (amortises the per-channel constant setup), but small
// tensors still get several blocks per SM (the kernels are latency-bound below ~4 resident blocks per SM);
// capped at 6 blocks per SM
static inline void bn_row_grid(int C, int64_t M, int min_rows, int& nx, int& ny, int& rows_per_block, int& blocks, int blocks_per_sm = 6) {
  bn_red_shape(C, nx, ny);
  int64_t want = (M + (int64_t)ny * min_rows - 1) / ((int64_t)ny * min_rows);
  want = std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms() * blocks_per_sm));
  rows_per_block = (int)((M + want - 1) / want);
  blocks = (int)((M + rows_per_block - 1) / rows_per_block);
}


}  // namespace iswm
