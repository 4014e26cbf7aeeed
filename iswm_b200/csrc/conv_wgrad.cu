// conv_wgrad.cu — weight gradient of a convolution on tcgen05 tensor cores.
//
//   dW[cout, tap, cin] = sum_pixels dy[pixel, cout] * x[pixel + offset(tap), cin]
//
// GEMM view: M = cout (128 per tile), N = cin (<= 256 per tile), K = pixels. Both
// operands are NHWC, i.e. the contraction index (pixel) is the slow one, so both are fed
// to the MMA as MN-major 128B-swizzled tiles: a TMA box of {64 channels, 64 pixels} lands
// as 64 rows (K) of 128 bytes (64 channels of M or N). The tap shift and the zero padding
// come from the TMA coordinates exactly as in conv_igemm.cu. The pixel range is split
// across CTAs (split-K) and partial tiles are reduced with fp32 atomics into dW.
//
// Replaces the autograd weight gradients of every nn.Conv2d on the hot path
// (network/backbone/resnet.py:27-35, network/_deeplab.py:37-51,124,134,149,162)
// produced by loss.backward() at train.py:1048.
#include "tc_common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace iswm {

constexpr int kWStages = 8;
constexpr int kWTileM = 128;
constexpr int kPixBlock = 64;               // pixels per k-block
constexpr int kChunkBytes = 64 * 128;       // one {64 ch x 64 px} box = 8 KiB
constexpr int kWTmemCols = 512;
constexpr int kWSmemBudget = 196608;

struct WgradKParams {
  int Cout, Cin, ntaps;
  int lgBW, lgBH;                 // pixel box: BW*BH*BB == 64
  int tiles_w, tiles_h, tiles_b;  // pixel blocks
  int pix_blocks;
  int tiles_m, tiles_n, BN, nchunks_b, stages;
  int n_tiles;                    // output tiles = tiles_m * ntaps * tiles_n
  long long total_kblocks;        // n_tiles * pix_blocks
  long long kblocks_per_cta;      // stream-K mode: equal contiguous ranges of the linearised (tile, pixel block) space
  int splits, split_len;          // split mode (splits > 0): CTA = (tile, split); same-split CTAs walk the same pixels
  int vec_red;
  int nprod;                      // producer warps (1 or 2) taking alternate pixel blocks
  int n_img_per_phase;
  int8_t dh[ISWM_MAX_TAPS], dw[ISWM_MAX_TAPS], phase[ISWM_MAX_TAPS];
  float* dwgt;
  int* abort_flag;
};

__global__ void __launch_bounds__(256, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy,
                  const __grid_constant__ CUtensorMap tmap_x, const WgradKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = tc::smem_u32(smem_raw);
  const uint32_t ring = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (ring - raw_addr);
  const uint32_t a_bytes = 2 * kChunkBytes;
  const uint32_t stage_bytes = a_bytes + (uint32_t)p.nchunks_b * kChunkBytes;
  uint8_t* tail = smem + (size_t)p.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWStages + 4);

  const uint32_t bar_full = tc::smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kWStages;
  const uint32_t bar_tfull = bar_empty + 8 * kWStages;
  const uint32_t bar_tempty = bar_tfull + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_dy);
    tc::tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; s++) {
      tc::mbar_init(bar_full + 8 * s, 1);
      tc::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; s++) {
      tc::mbar_init(bar_tfull + 8 * s, 1);
      tc::mbar_init(bar_tempty + 8 * s, 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tc::smem_u32(tmem_slot), kWTmemCols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // prologue done (barriers, TMEM, descriptor prefetch touch no global data): wait for the producer grid, then let
  // the next kernel of the stream start its own prologue under our main loop
  pdl_wait();
  pdl_launch();

  const int BW = 1 << p.lgBW, BH = 1 << p.lgBH;
  const int BB = kPixBlock >> (p.lgBW + p.lgBH);

  // Stream-K: the (output tile, pixel block) space is linearised and cut into equal contiguous
  // ranges, one per CTA; a CTA flushes its partial tile with fp32 reductions whenever its range
  // crosses a tile boundary. tile -> (m tile, tap, n tile), n fastest (neighbours share dy in L2).
  auto decode = [&](int tile, int& mt, int& tap, int& nt) {
    nt = tile % p.tiles_n;
    const int r = tile / p.tiles_n;
    tap = r % p.ntaps;
    mt = r / p.ntaps;
  };
  long long range_lo, range_hi;
  if (p.splits > 0) {
    // every CTA of one split index reads the SAME pixel blocks at the same time (different output tiles), so dy / x
    // stream from HBM once and are shared through L2; partial tiles are reduced with fp32 atomics at the end
    const int t = (int)(blockIdx.x % (unsigned)p.n_tiles), sp = (int)(blockIdx.x / (unsigned)p.n_tiles);
    range_lo = (long long)t * p.pix_blocks + (long long)sp * p.split_len;
    range_hi = min(range_lo + p.split_len, (long long)(t + 1) * p.pix_blocks);
  } else {
    range_lo = (long long)blockIdx.x * p.kblocks_per_cta;
    range_hi = min(range_lo + p.kblocks_per_cta, p.total_kblocks);
  }

  // producer / MMA roles: ONE elected thread runs the whole schedule (see conv_igemm.cu); producer warps 0 and 3 take
  // alternate k-blocks (pixel blocks) of the CTA's range; pixel-block coordinates advance incrementally (no divisions
  // inside the k loop)
  if (warp == 0 || (warp == 3 && p.nprod == 2)) {
    if (tc::elect_one()) {
      const int me = (warp == 0) ? 0 : 1;
      const int np = p.nprod;
      int stage = me % p.stages;
      uint32_t phase = 0;
      long long q0 = 0;                         // k-blocks of this CTA issued before the current segment
      bool ok = true;
      for (long long cur = range_lo; cur < range_hi && ok;) {
        const int tile = (int)(cur / p.pix_blocks);
        const int pb0 = (int)(cur - (long long)tile * p.pix_blocks);
        const int pb1 = (int)min((long long)p.pix_blocks, pb0 + (range_hi - cur));
        int mt, tap, nt;
        decode(tile, mt, tap, nt);
        const int m0 = mt * kWTileM, n0 = nt * p.BN;
        int off = me - (int)(q0 % np);
        if (off < 0) off += np;
        int pb = pb0 + off;
        int tw = pb % p.tiles_w, th = (pb / p.tiles_w) % p.tiles_h, tb = pb / (p.tiles_w * p.tiles_h);
        const int ddw = p.dw[tap], ddh = p.dh[tap], dph = p.phase[tap] * p.n_img_per_phase;
        for (; pb < pb1; pb += np) {
          const int w0 = tw * BW, h0 = th * BH, b0 = tb * BB;
          ok = tc::mbar_wait(bar_empty + 8 * stage, phase ^ 1, p.abort_flag, 11);
          if (!ok) break;
          const uint32_t dst = ring + stage * stage_bytes;
          const uint32_t bar = bar_full + 8 * stage;
          const int xw = w0 + ddw, xh = h0 + ddh, xb = dph + b0;
          tc::mbar_expect_tx(bar, stage_bytes);
          tc::tma_load_4d(dst, &tmap_dy, bar, m0, w0, h0, b0);
          tc::tma_load_4d(dst + kChunkBytes, &tmap_dy, bar, m0 + 64, w0, h0, b0);
          for (int j = 0; j < p.nchunks_b; j++)
            tc::tma_load_4d(dst + a_bytes + j * kChunkBytes, &tmap_x, bar, n0 + 64 * j, xw, xh, xb);
          stage += np;
          if (stage >= p.stages) { stage -= p.stages; phase ^= 1; }
          tw += np;
          while (tw >= p.tiles_w) { tw -= p.tiles_w; th++; }
          while (th >= p.tiles_h) { th -= p.tiles_h; tb++; }
        }
        q0 += pb1 - pb0;
        cur += pb1 - pb0;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_bf16(kWTileM, p.BN, 1, 1);   // both operands MN-major
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      bool ok = true;
      for (long long cur = range_lo; cur < range_hi && ok;) {
        const int tile = (int)(cur / p.pix_blocks);
        const int pb0 = (int)(cur - (long long)tile * p.pix_blocks);
        const int pb1 = (int)min((long long)p.pix_blocks, pb0 + (range_hi - cur));
        ok = tc::mbar_wait(bar_tempty + 8 * as, aphase ^ 1, p.abort_flag, 12);
        if (!ok) break;
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.BN);
        for (int pb = pb0; pb < pb1; pb++) {
          ok = tc::mbar_wait(bar_full + 8 * stage, phase, p.abort_flag, 13);
          if (!ok) break;
          tc::tc_fence_after();
          const uint32_t a_addr = ring + stage * stage_bytes;
          // MN-major SW128: LBO = byte stride between 64-channel chunks, SBO = 8 pixel rows
          const uint64_t da = tc::make_smem_desc_sw128(a_addr, kChunkBytes, 1024);
          const uint64_t db = tc::make_smem_desc_sw128(a_addr + a_bytes, kChunkBytes, 1024);
          const uint32_t first = (pb > pb0) ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < kPixBlock / 16; k++)     // 16 pixel rows = 2048 bytes per MMA
            tc::umma_bf16(d_tmem, da + 128 * k, db + 128 * k, idesc, first | (uint32_t)(k > 0));
          tc::umma_commit(bar_empty + 8 * stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
        tc::umma_commit(bar_tfull + 8 * as);
        as ^= 1;
        if (as == 0) aphase ^= 1;
        cur += pb1 - pb0;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (long long cur = range_lo; cur < range_hi;) {
      const int tile = (int)(cur / p.pix_blocks);
      const int pb0 = (int)(cur - (long long)tile * p.pix_blocks);
      const int pb1 = (int)min((long long)p.pix_blocks, pb0 + (range_hi - cur));
      int mt, tap, nt;
      decode(tile, mt, tap, nt);
      const int m = mt * kWTileM + row, n0 = nt * p.BN;
      if (!tc::mbar_wait(bar_tfull + 8 * as, aphase, p.abort_flag, 14)) break;
      tc::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * p.BN);
      float* orow = p.dwgt + ((size_t)m * p.ntaps + tap) * p.Cin;
      for (int c = 0; c < p.BN; c += 16) {
        uint32_t v[16];
        tc::tmem_ld16(t_row + c, v);
        tc::tmem_ld_wait();
        const int n = n0 + c;
        if (m < p.Cout) {
          if (p.vec_red && n + 16 <= p.Cin) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + n + j),
                           "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                           "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 16; j++)
              if (n + j < p.Cin) atomicAdd(orow + n + j, __uint_as_float(v[j]));
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_tempty + 8 * as);
      as ^= 1;
      if (as == 0) aphase ^= 1;
      cur += pb1 - pb0;
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, kWTmemCols);
}

static int ilog2_ceil_w(int v) {
  int l = 0;
  while ((1 << l) < v) l++;
  return l;
}

}  // namespace iswm

using namespace iswm;

extern "C" int iswm_conv_wgrad(const iswm_conv_desc* d, const void* d_in, const void* d_dy,
                               float* d_dw, void* stream) {
  if (debug_skip(ISWM_SKIP_CONV_WGRAD)) return 0;
  ISWM_REQUIRE(d && d_in && d_dy && d_dw, "conv_wgrad: null argument");
  ISWM_REQUIRE(d->ntaps >= 1 && d->ntaps <= ISWM_MAX_TAPS, "conv_wgrad: ntaps=%d", d->ntaps);
  ISWM_REQUIRE((d->in_ld % 8) == 0 && (d->out_ld % 8) == 0, "conv_wgrad: row pitches must be multiples of 8 (in_ld=%d out_ld=%d)", d->in_ld, d->out_ld);
  ISWM_REQUIRE((reinterpret_cast<uintptr_t>(d_in) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_dy) & 15) == 0, "conv_wgrad: operands must be 16-byte aligned");
  int* abort_flag = abort_flag_ptr();
  ISWM_REQUIRE(abort_flag, "conv_wgrad: cannot allocate abort flag");

  WgradKParams p;
  memset(&p, 0, sizeof(p));
  int B = d->B, Hi = d->Hi, Wi = d->Wi, Ho = d->Ho, Wo = d->Wo, n_img = d->n_img;
  bool pointwise = (d->ntaps == 1 && d->dh[0] == 0 && d->dw[0] == 0 && d->phase[0] == 0 &&
                    Hi == Ho && Wi == Wo && n_img == B);
  if (pointwise) {
    const int64_t M = (int64_t)B * Ho * Wo;
    if (M < (1ll << 31)) {
      Wo = Wi = (int)M;
      Ho = Hi = 1;
      B = n_img = 1;
    }
  }
  p.Cout = d->Cout; p.Cin = d->Cin; p.ntaps = d->ntaps;
  p.lgBW = std::min(6, ilog2_ceil_w(Wo));
  p.lgBH = std::min(6 - p.lgBW, ilog2_ceil_w(Ho));
  const int BW = 1 << p.lgBW, BH = 1 << p.lgBH, BB = kPixBlock / (BW * BH);
  p.tiles_w = (Wo + BW - 1) / BW;
  p.tiles_h = (Ho + BH - 1) / BH;
  p.tiles_b = (B + BB - 1) / BB;
  const int64_t pbs = (int64_t)p.tiles_w * p.tiles_h * p.tiles_b;
  ISWM_REQUIRE(pbs < (1ll << 31), "conv_wgrad: too many pixel blocks");
  p.pix_blocks = (int)pbs;
  p.tiles_m = (d->Cout + kWTileM - 1) / kWTileM;
  const int n_split_n = (d->Cin + 255) / 256;
  p.BN = std::min(256, (((d->Cin + n_split_n - 1) / n_split_n + 15) / 16) * 16);
  p.tiles_n = (d->Cin + p.BN - 1) / p.BN;
  p.nchunks_b = (p.BN + 63) / 64;
  const int64_t n_tiles = (int64_t)p.tiles_m * p.ntaps * p.tiles_n;
  ISWM_REQUIRE(n_tiles < (1ll << 31), "conv_wgrad: too many output tiles");
  p.n_tiles = (int)n_tiles;
  p.total_kblocks = n_tiles * pbs;
  int grid;
  const int sms = num_sms();
  if (n_tiles <= 2 * sms) {
    // split mode: pick the split count (each split >= ~4 pixel blocks) that fills whole waves of SMs best
    const int max_splits = (int)std::max<int64_t>(1, std::min<int64_t>(pbs / 4, (4 * sms) / n_tiles));
    int best = 1;
    double best_eff = 0.0;
    for (int sp = 1; sp <= max_splits; sp++) {
      const int64_t ctas = n_tiles * sp;
      const double eff = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms) - 1e-4 * sp;
      if (eff > best_eff + 0.02) { best_eff = eff; best = sp; }
    }
    p.split_len = (int)((pbs + best - 1) / best);
    p.splits = (int)((pbs + p.split_len - 1) / p.split_len);
    grid = (int)(n_tiles * p.splits);
  } else {
    // stream-K: equal k-block ranges per CTA; at least ~4 k-blocks each so a flush is amortised
    grid = (int)std::max<int64_t>(1, std::min<int64_t>(sms, p.total_kblocks / 4));
    p.kblocks_per_cta = (p.total_kblocks + grid - 1) / grid;
    grid = (int)((p.total_kblocks + p.kblocks_per_cta - 1) / p.kblocks_per_cta);
  }
  p.vec_red = ((reinterpret_cast<uintptr_t>(d_dw) & 15) == 0 && (d->Cin % 4) == 0) ? 1 : 0;
  const int stage_bytes = (2 + p.nchunks_b) * kChunkBytes;
  p.stages = std::max(2, std::min(kWStages, kWSmemBudget / stage_bytes));
  p.n_img_per_phase = B;
  static const int env_np = [] { const char* e = getenv("ISWM_WGRAD_NPROD"); return e ? atoi(e) : 0; }();
  p.nprod = (env_np == 1 || env_np == 2) ? env_np : (p.stages >= 4 ? 2 : 1);
  for (int t = 0; t < d->ntaps; t++) {
    p.dh[t] = d->dh[t]; p.dw[t] = d->dw[t]; p.phase[t] = d->phase[t];
    ISWM_REQUIRE(d->coff[t] == 0, "conv_wgrad: per-tap channel offsets are a forward / data-gradient feature");
  }
  p.dwgt = d_dw;
  p.abort_flag = abort_flag;

  CUtensorMap tmap_dy, tmap_x;
  {
    const uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    const uint64_t str[4] = {1, (uint64_t)d->out_ld, (uint64_t)Wo * d->out_ld, (uint64_t)Ho * Wo * d->out_ld};
    const uint32_t box[4] = {64u, (uint32_t)BW, (uint32_t)BH, (uint32_t)BB};
    if (int rc = encode_tmap_bf16(&tmap_dy, d_dy, 4, dims, str, box)) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)n_img};
    const uint64_t str[4] = {1, (uint64_t)d->in_ld, (uint64_t)Wi * d->in_ld, (uint64_t)Hi * Wi * d->in_ld};
    const uint32_t box[4] = {64u, (uint32_t)BW, (uint32_t)BH, (uint32_t)BB};
    if (int rc = encode_tmap_bf16(&tmap_x, d_in, 4, dims, str, box)) return rc;
  }
  const int smem_bytes = p.stages * stage_bytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    ISWM_REQUIRE(e == cudaSuccess, "conv_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  launch_k(conv_wgrad_kernel, dim3(grid), dim3(256), smem_bytes, static_cast<cudaStream_t>(stream), tmap_dy, tmap_x, p);
  return check_launch("conv_wgrad");
}
