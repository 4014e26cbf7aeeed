// sgd_pack.cu — the optimiser step and the bf16 operand repack as ONE pass over the weights.
//
// train.py:421-431 builds torch.optim.SGD(momentum, nesterov, weight_decay) and train.py:1049 steps it; the engine then needs
// every convolution's weights again as packed bf16 MMA operands (forward [Cout][tap][Cin_pad], data gradient
// [Cin][tap][Cout_pad]). Run separately that is a pass over the fp32 master weights for the update (read w, g, m; write w, m)
// and a second pass re-reading them twice for the two packings. Here a job table covers the whole flat parameter buffer:
//   conv jobs   tiles of 16 output channels x TC input channels x all taps: w / g / m read once (coalesced, the source is
//               [Cout][Cin][tap] with tap innermost), the updated weights go back to the master buffer AND, from the shared-memory
//               tile, into both packed layouts with sector-sized stores
//   plain jobs  everything between the convolution weights in the flat buffer (BatchNorm affine parameters, the classifier
//               bias): the update alone
// The arithmetic is sgd_step_kernel's (same fma chain): weights and momentum are bit-identical to the two-kernel path, and
// so are the packed operands (bf16 round-to-nearest of the same fp32 values).
#include "common.cuh"
#include "ew_common.cuh"
#include <stdlib.h>

namespace iswm {

struct SgdPackJob {
  float* w;                 // master weights (in/out)
  const float* g;           // gradient
  float* m;                 // momentum buffer (in/out) or nullptr
  __nv_bfloat16* dst_f;     // forward operand or nullptr (plain job)
  __nv_bfloat16* dst_d;     // data-gradient operand or nullptr
  long long n;              // elements (plain jobs and the stem)
  int Cout, Cin, RS;
  int pad_f, row_ld_f;      // forward: channels per tap (Cin padded to 64), elements per output channel
  int pad_d, row_ld_d;      // dgrad: Cout padded to 64; taps per row of a K-concatenated operand (0 = RS)
  int mode;                 // 0 plain, 1 convolution tiles, 2 stem (row-tap forward operand, no dgrad operand)
  int TC;                   // input channels per tile
  int blk_begin, blk_count;
};

struct SgdHyper {
  float lr, momentum, wd;
  int nesterov, first;
  const float* d_lr;
};

__device__ __forceinline__ float sgd_update(float w, float grad, float* mom_slot, const SgdHyper& h, float lr) {
  if (h.wd != 0.f) grad = fmaf(h.wd, w, grad);
  if (h.momentum != 0.f) {
    const float buf = h.first ? grad : fmaf(h.momentum, *mom_slot, grad);
    *mom_slot = buf;
    grad = h.nesterov ? fmaf(h.momentum, buf, grad) : buf;
  }
  return fmaf(-lr, grad, w);
}

constexpr int kTO = 16;                 // output channels per tile
constexpr int kTileMax = 16 * 144;      // floats of shared memory per tile

__global__ void __launch_bounds__(kT, 4)
sgd_pack_kernel(const SgdPackJob* __restrict__ jobs, int n_jobs, SgdHyper h) {
  pdl_wait();
  pdl_launch();
  __shared__ float tile[kTileMax];
  const float lr = h.d_lr ? *h.d_lr : h.lr;
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].blk_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const SgdPackJob j = jobs[lo];
  const int lb = (int)blockIdx.x - j.blk_begin, nb = j.blk_count;
  if (lb >= nb) return;
  const int tid = threadIdx.x;
  float dummy = 0.f;
  if (j.mode == 0 || j.mode == 2) {
    int ks = 1;
    while (ks * ks < j.RS) ks++;
    for (long long i = (long long)lb * kT + tid; i < j.n; i += (long long)nb * kT) {
      const float wn = sgd_update(j.w[i], j.g[i], j.m ? j.m + i : &dummy, h, lr);
      j.w[i] = wn;
      if (j.mode == 2) {
        // stem: every weight has exactly one place in the row-tap forward operand [o][r][s * Cin + c] (the padding never changes)
        const int t = (int)(i % j.RS), c = (int)((i / j.RS) % j.Cin), o = (int)(i / ((long long)j.RS * j.Cin));
        const int r = t / ks, sx = t - r * ks;
        j.dst_f[(long long)o * j.row_ld_f + r * j.pad_f + sx * j.Cin + c] = __float2bfloat16_rn(wn);
      }
    }
    return;
  }
  // convolution: tiles of kTO output channels x TC input channels x RS taps
  const int RS = j.RS, TC = j.TC;
  const int ctiles = (j.Cin + TC - 1) / TC, otiles = (j.Cout + kTO - 1) / kTO;
  const int F = TC * RS;                                  // floats per output channel of a full tile
  for (int u = lb; u < ctiles * otiles; u += nb) {
    const int ot = u / ctiles, ct = u - ot * ctiles;
    const int o0 = ot * kTO, c0 = ct * TC;
    const int no = min(kTO, j.Cout - o0), nc = min(TC, j.Cin - c0);
    const int nf = nc * RS;                               // valid floats per output channel in this tile
    __syncthreads();                                      // the previous tile's readers are done
    {
      // all of a thread's loads first (w, g, m of up to 9 elements), then the updates and stores: written as one loop the
      // compiler orders every iteration's loads after the previous iteration's stores (the pointers may alias) and the tile
      // costs nine dependent memory round trips
      constexpr int kIter = kTileMax / kT;
      float wv[kIter], gv[kIter], mv[kIter];
      int gidx[kIter];                                    // offsets from the tile's first element (16 output channels: fits 31 bits)
      const long long base = ((long long)o0 * j.Cin + c0) * RS;
      float* const wb = j.w + base;
      const float* const gb = j.g + base;
      float* const mb = j.m ? j.m + base : nullptr;
      const int ostride = j.Cin * RS;
#pragma unroll
      for (int k = 0; k < kIter; k++) {
        const int i = tid + k * kT;
        const int o = i / F, f = i - o * F;
        const bool ok = (i < no * F) && (f < nf);
        gidx[k] = ok ? o * ostride + f : -1;
        wv[k] = ok ? wb[gidx[k]] : 0.f;
        gv[k] = ok ? gb[gidx[k]] : 0.f;
        mv[k] = (ok && mb && !h.first) ? mb[gidx[k]] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < kIter; k++) {
        if (gidx[k] < 0) continue;
        const int i = tid + k * kT;
        float mslot = mv[k];
        const float wn = sgd_update(wv[k], gv[k], &mslot, h, lr);
        wb[gidx[k]] = wn;
        if (mb && h.momentum != 0.f) mb[gidx[k]] = mslot;
        tile[i] = wn;                                     // tile[o][c * RS + t] (i = o * F + f)
      }
    }
    __syncthreads();
    // forward operand: dst_f[o][t][c0 + c]: 8 channels (16 bytes) per store where the run allows it
    {
      const int cg = (nc + 7) >> 3;
      for (int i = tid; i < no * RS * cg; i += kT) {
        const int c = (i % cg) << 3;
        const int t = (i / cg) % RS, o = i / (cg * RS);
        __nv_bfloat16* dst = j.dst_f + (long long)(o0 + o) * j.row_ld_f + (long long)t * j.pad_f + c0 + c;
        if (c + 8 <= nc) {
          F8 v;
#pragma unroll
          for (int k = 0; k < 8; k++) v.v[k] = tile[o * F + (c + k) * RS + t];
          store8(dst, v);
        } else {
          for (int k = 0; c + k < nc; k++) dst[k] = __float2bfloat16_rn(tile[o * F + (c + k) * RS + t]);
        }
      }
    }
    // data-gradient operand: dst_d[row(c0 + c, t)][o0 + o]: 8 output channels per store
    if (j.dst_d) {
      const int og = (no + 7) >> 3;
      for (int i = tid; i < nf * og; i += kT) {
        const int o = (i % og) << 3;
        const int f = i / og;                             // c * RS + t inside the tile
        const int c = f / RS, t = f - c * RS;
        const long long rowi = j.row_ld_d > 0 ? (long long)(c0 + c) * j.row_ld_d + t : (long long)(c0 + c) * RS + t;
        __nv_bfloat16* dst = j.dst_d + rowi * j.pad_d + o0 + o;
        if (o + 8 <= no) {
          F8 v;
#pragma unroll
          for (int k = 0; k < 8; k++) v.v[k] = tile[(o + k) * F + f];
          store8(dst, v);
        } else {
          for (int k = 0; o + k < no; k++) dst[k] = __float2bfloat16_rn(tile[(o + k) * F + f]);
        }
      }
    }
  }
}

}  // namespace iswm

using namespace iswm;

extern "C" int iswm_sgd_pack_batched(const void* d_jobs, int n_jobs, int total_blocks, float lr, float momentum, float weight_decay,
                                     int nesterov, int first_step, const float* d_lr, void* stream) {
  ISWM_REQUIRE(d_jobs && n_jobs >= 1 && total_blocks >= 1, "sgd_pack_batched: empty job list");
  static_assert(sizeof(SgdPackJob) == sizeof(iswm_sgd_pack_job), "job struct mirrors the header");
  const SgdHyper h{lr, momentum, weight_decay, nesterov, first_step, d_lr};
  launch_k(sgd_pack_kernel, dim3((unsigned)total_blocks), dim3(kT), 0, static_cast<cudaStream_t>(stream),
           static_cast<const SgdPackJob*>(d_jobs), n_jobs, h);
  return check_launch("sgd_pack_batched");
}
