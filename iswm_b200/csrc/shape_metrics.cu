// shape_metrics.cu - the image-processing half of the reference's per-frame evaluators on the device:
//   metrics/utils/mask_utils.py:7-52      preprocess_mask: binarise, 3x3 close, 3x3 open, 8-connected components with areas,
//                                         largest region among those >= 0.1 % of the frame (cv2.morphologyEx,
//                                         cv2.connectedComponentsWithStats)
//   metrics/utils/mask_utils.py:54-76     find_front_positions: leftmost pixel of every row of that region
//   metrics/utils/mask_utils.py:105-135   calculate_stability: first pixel of the previous frame inside a +-window of the front
//   metrics/front_tracking_metrics.py:44-90  nearest front point in the other mask (two python loops, O(H^2) per frame pair)
//   metrics/region_metrics.py:6-12, :44-61, :74-92  repair_small_gaps (dilate x3, erode x2), intersection / union, region areas
//                                         (scipy.ndimage.label, 8-connected)
// Everything here is integer / byte work and bit-exact; the float64 scalar arithmetic on top of these results (a few hundred
// numbers per frame) stays on the host in iswm_b200/metrics/shape_metrics.py, in the reference's own operation order.
//
// Connected components: label-equivalence union-find over the pixel grid (atomicMin hooking of roots, Playne / Komura style):
// init -> union with the W, NW, N, NE neighbours -> flatten + per-root area / first-block key -> per-frame selection.
// A component's root is its smallest pixel index; OpenCV's label ORDER (needed for np.argmax ties between equal areas) is the
// block-raster order of each component's first 2x2 block, carried as key = (y/2) * ceil(W/2) + x/2.
#include "ew_common.cuh"
#include <limits.h>

namespace iswm {
namespace {

template <typename T>
__global__ void __launch_bounds__(kT)
morph_kernel(const T* __restrict__ src, uint8_t* __restrict__ dst, int N, int H, int W, int r, int erode) {
  pdl_wait();
  pdl_launch();
  const int64_t total = (int64_t)N * H * W;
  for (int64_t g = (int64_t)blockIdx.x * kT + threadIdx.x; g < total; g += (int64_t)gridDim.x * kT) {
    const int x = (int)(g % W), y = (int)((g / W) % H);
    const int64_t base = g - ((int64_t)y * W + x);
    const int y0 = max(0, y - r), y1 = min(H - 1, y + r), x0 = max(0, x - r), x1 = min(W - 1, x + r);
    int v = erode ? 1 : 0;                                   // pixels outside the frame are ignored (cv2 default border)
    for (int yy = y0; yy <= y1; yy++)
      for (int xx = x0; xx <= x1; xx++) {
        const int s = src[base + (int64_t)yy * W + xx] > 0 ? 1 : 0;
        v = erode ? (v & s) : (v | s);
      }
    dst[g] = (uint8_t)v;
  }
}

__device__ __forceinline__ int find_root(int* L, int i) {
  int p;
  while ((p = *reinterpret_cast<volatile int*>(L + i)) != i) i = p;
  return i;
}
__device__ __forceinline__ void unite(int* L, int a, int b) {
  bool done = false;
  while (!done) {
    a = find_root(L, a);
    b = find_root(L, b);
    if (a < b) {
      const int old = atomicMin(L + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(L + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  }
}

__global__ void __launch_bounds__(kT)
ccl_init_kernel(const uint8_t* __restrict__ m, int64_t total, int* __restrict__ L, int* __restrict__ area, int* __restrict__ key) {
  pdl_wait();
  pdl_launch();
  for (int64_t g = (int64_t)blockIdx.x * kT + threadIdx.x; g < total; g += (int64_t)gridDim.x * kT) {
    L[g] = m[g] ? (int)g : -1;
    area[g] = 0;
    key[g] = INT_MAX;
  }
}

__global__ void __launch_bounds__(kT)
ccl_union_kernel(const uint8_t* __restrict__ m, int N, int H, int W, int* __restrict__ L) {
  pdl_wait();
  pdl_launch();
  const int64_t total = (int64_t)N * H * W;
  for (int64_t g = (int64_t)blockIdx.x * kT + threadIdx.x; g < total; g += (int64_t)gridDim.x * kT) {
    if (!m[g]) continue;
    const int x = (int)(g % W), y = (int)((g / W) % H);
    // decision tree over the scan mask (W, NW, N, NE): a set N already holds W, NW and NE together (they are its row
    // neighbours, joined when their own pixels were visited); without N, W stands for NW (NW is W's N) and NE is on its own
    if (y > 0 && m[g - W]) {
      unite(L, (int)g, (int)(g - W));
    } else {
      if (x > 0 && m[g - 1]) unite(L, (int)g, (int)(g - 1));
      else if (y > 0 && x > 0 && m[g - W - 1]) unite(L, (int)g, (int)(g - W - 1));
      if (y > 0 && x < W - 1 && m[g - W + 1]) unite(L, (int)g, (int)(g - W + 1));
    }
  }
}

// every pixel points at its root; per-root area and first-block key
__global__ void __launch_bounds__(kT)
ccl_stats_kernel(const uint8_t* __restrict__ m, int N, int H, int W, int* __restrict__ L, int* __restrict__ area, int* __restrict__ key) {
  pdl_wait();
  pdl_launch();
  const int64_t total = (int64_t)N * H * W;
  const int wb = (W + 1) >> 1;
  for (int64_t g = (int64_t)blockIdx.x * kT + threadIdx.x; g < total; g += (int64_t)gridDim.x * kT) {
    if (!m[g]) continue;
    const int root = find_root(L, (int)g);
    if (root != (int)g) L[g] = root;                         // concurrent walkers see the old parent or the root: both are ancestors
    const int x = (int)(g % W), y = (int)((g / W) % H);
    atomicAdd(area + root, 1);
    atomicMin(key + root, (y >> 1) * wb + (x >> 1));
  }
}

constexpr int kSel = 1024;
// info[n] = {components, valid components (area >= min_valid_area), best area, support pixel count (filled by the next kernel),
//            best root (global pixel index, -1 = none), 0, 0, 0}: best = largest area, ties to the FIRST label in OpenCV's order
__global__ void __launch_bounds__(kSel)
ccl_select_kernel(const uint8_t* __restrict__ m, int H, int W, const int* __restrict__ L, const int* __restrict__ area,
                  const int* __restrict__ key, double min_valid_area, int* __restrict__ info) {
  pdl_wait();
  pdl_launch();
  __shared__ int s_comp[kSel / 32], s_valid[kSel / 32], s_area[kSel / 32], s_key[kSel / 32], s_root[kSel / 32];
  const int n = blockIdx.x;
  const int64_t base = (int64_t)n * H * W;
  int comp = 0, valid = 0, barea = -1, bkey = INT_MAX, broot = -1;
  for (int p = threadIdx.x; p < H * W; p += kSel) {
    const int64_t g = base + p;
    if (m[g] && L[g] == (int)g) {
      comp++;
      const int a = area[g], k = key[g];
      if ((double)a >= min_valid_area) {
        valid++;
        if (a > barea || (a == barea && k < bkey)) { barea = a; bkey = k; broot = (int)g; }
      }
    }
  }
  auto better = [](int a1, int k1, int a2, int k2) { return a1 > a2 || (a1 == a2 && k1 < k2); };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    comp += __shfl_xor_sync(0xffffffffu, comp, o);
    valid += __shfl_xor_sync(0xffffffffu, valid, o);
    const int a2 = __shfl_xor_sync(0xffffffffu, barea, o), k2 = __shfl_xor_sync(0xffffffffu, bkey, o), r2 = __shfl_xor_sync(0xffffffffu, broot, o);
    if (better(a2, k2, barea, bkey)) { barea = a2; bkey = k2; broot = r2; }
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_comp[w] = comp; s_valid[w] = valid; s_area[w] = barea; s_key[w] = bkey; s_root[w] = broot; }
  __syncthreads();
  if (threadIdx.x == 0) {
    comp = valid = 0; barea = -1; bkey = INT_MAX; broot = -1;
    for (int i = 0; i < kSel / 32; i++) {
      comp += s_comp[i];
      valid += s_valid[i];
      if (better(s_area[i], s_key[i], barea, bkey)) { barea = s_area[i]; bkey = s_key[i]; broot = s_root[i]; }
    }
    int* o = info + n * 8;
    o[0] = comp; o[1] = valid; o[2] = valid > 0 ? barea : 0; o[3] = 0; o[4] = valid > 0 ? broot : -1; o[5] = 0; o[6] = 0; o[7] = 0;
  }
}

// one block per row: support = pixels of the chosen component; front = its leftmost column in this row (or -1)
__global__ void __launch_bounds__(128)
support_front_kernel(const uint8_t* __restrict__ m, int H, int W, const int* __restrict__ L, int* __restrict__ info,
                     uint8_t* __restrict__ support, int* __restrict__ front) {
  pdl_wait();
  pdl_launch();
  __shared__ int s_min, s_cnt;
  const int row = blockIdx.x, n = row / H;
  const int best = info[n * 8 + 4];
  if (threadIdx.x == 0) { s_min = INT_MAX; s_cnt = 0; }
  __syncthreads();
  const int64_t base = (int64_t)row * W;
  int mn = INT_MAX, cnt = 0;
  for (int x = threadIdx.x; x < W; x += 128) {
    const int64_t g = base + x;
    const int s = (best >= 0 && m[g] && L[g] == best) ? 1 : 0;
    support[g] = (uint8_t)s;
    if (s) { cnt++; mn = min(mn, x); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt) { atomicMin(&s_min, mn); atomicAdd(&s_cnt, cnt); }
  __syncthreads();
  if (threadIdx.x == 0) {
    front[row] = s_min == INT_MAX ? -1 : s_min;
    if (s_cnt) atomicAdd(info + n * 8 + 3, s_cnt);
  }
}

// region_metrics.py: counts[n] = {sum(pred > 0), sum(gt > 0), |repaired & gt|, |repaired | gt|, regions >= min_area, components, 0, 0}
template <typename TP, typename TG>
__global__ void __launch_bounds__(kT)
region_counts_kernel(const TP* __restrict__ pred, const TG* __restrict__ gt, const uint8_t* __restrict__ repaired, int HW, int* __restrict__ counts) {
  pdl_wait();
  pdl_launch();
  __shared__ int s[4];
  const int n = blockIdx.y;
  if (threadIdx.x < 4) s[threadIdx.x] = 0;
  __syncthreads();
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (int p = blockIdx.x * kT + threadIdx.x; p < HW; p += gridDim.x * kT) {
    const int64_t g = (int64_t)n * HW + p;
    const int a = pred[g] > 0, b = gt[g] > 0, r = repaired[g];
    c0 += a; c1 += b; c2 += (r & b); c3 += (r | b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s[0], c0); atomicAdd(&s[1], c1); atomicAdd(&s[2], c2); atomicAdd(&s[3], c3); }
  __syncthreads();
  if (threadIdx.x < 4 && s[threadIdx.x]) atomicAdd(counts + n * 8 + threadIdx.x, s[threadIdx.x]);
}

// areas of the components with area >= min_area, in no particular order (the host sorts them), at most `cap` per frame
__global__ void __launch_bounds__(kSel)
region_list_kernel(const uint8_t* __restrict__ m, int HW, const int* __restrict__ L, const int* __restrict__ area, int min_area, int cap,
                   int* __restrict__ counts, int* __restrict__ areas) {
  pdl_wait();
  pdl_launch();
  const int n = blockIdx.x;
  const int64_t base = (int64_t)n * HW;
  for (int p = threadIdx.x; p < HW; p += kSel) {
    const int64_t g = base + p;
    if (m[g] && L[g] == (int)g) {
      atomicAdd(counts + n * 8 + 5, 1);
      const int a = area[g];
      if (a >= min_area) {
        const int slot = atomicAdd(counts + n * 8 + 4, 1);
        if (slot < cap) areas[(int64_t)n * cap + slot] = a;
      }
    }
  }
}

// front_tracking_metrics.py:48-63 / :72-85: for every front point of A the FIRST closest front point of B (row order)
__global__ void __launch_bounds__(kT)
front_nearest_kernel(const int* __restrict__ fa, const int* __restrict__ fb, int H, int* __restrict__ d2, int* __restrict__ dx) {
  pdl_wait();
  pdl_launch();
  extern __shared__ int s_fb[];
  const int n = blockIdx.x;
  for (int j = threadIdx.x; j < H; j += kT) s_fb[j] = fb[(int64_t)n * H + j];
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += kT) {
    const int ax = fa[(int64_t)n * H + i];
    long long best = LLONG_MAX;
    int bdx = -1;
    if (ax >= 0) {
      for (int j = 0; j < H; j++) {
        const int bx = s_fb[j];
        if (bx < 0) continue;
        const long long dy = i - j, ddx = ax - bx;
        const long long d = dy * dy + ddx * ddx;
        if (d < best) { best = d; bdx = (int)(ddx < 0 ? -ddx : ddx); }
      }
    }
    d2[(int64_t)n * H + i] = best == LLONG_MAX ? -1 : (int)best;
    dx[(int64_t)n * H + i] = bdx;
  }
}

// mask_utils.py:117-133: distance from this row's front to the first set pixel of the other mask inside [front - w, front + w)
__global__ void __launch_bounds__(kT)
front_window_kernel(const int* __restrict__ front, const uint8_t* __restrict__ other, int N, int H, int W, int window, int* __restrict__ diff) {
  pdl_wait();
  pdl_launch();
  const int total = N * H;
  for (int r = blockIdx.x * kT + threadIdx.x; r < total; r += gridDim.x * kT) {
    const int cf = front[r];
    int out = -1;
    if (cf >= 0) {
      const int s = max(0, cf - window), e = min(W, cf + window);
      const uint8_t* row = other + (int64_t)r * W;
      for (int x = s; x < e; x++)
        if (row[x]) { out = cf > x ? cf - x : x - cf; break; }
    }
    diff[r] = out;
  }
}

struct Work {
  uint8_t *a, *b;
  int *L, *area, *key;
};
inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }
inline Work carve(void* d_work, int64_t total) {
  uint8_t* p = static_cast<uint8_t*>(d_work);
  Work w;
  w.a = p; p += align256(total);
  w.b = p; p += align256(total);
  w.L = reinterpret_cast<int*>(p); p += align256(total * 4);
  w.area = reinterpret_cast<int*>(p); p += align256(total * 4);
  w.key = reinterpret_cast<int*>(p);
  return w;
}
inline int ew_grid(int64_t total) { return (int)std::max<int64_t>(1, std::min<int64_t>((total + kT - 1) / kT, (int64_t)num_sms() * 8)); }

template <typename T>
void morph_first(const void* src, uint8_t* dst, int N, int H, int W, int r, int erode, cudaStream_t st) {
  launch_k(morph_kernel<T>, dim3(ew_grid((int64_t)N * H * W)), dim3(kT), 0, st, static_cast<const T*>(src), dst, N, H, W, r, erode);
}
int morph_any(const void* src, int dtype, uint8_t* dst, int N, int H, int W, int r, int erode, cudaStream_t st) {
  switch (dtype) {
    case ISWM_U8: morph_first<uint8_t>(src, dst, N, H, W, r, erode, st); break;
    case ISWM_I32: morph_first<int32_t>(src, dst, N, H, W, r, erode, st); break;
    case ISWM_I64: morph_first<int64_t>(src, dst, N, H, W, r, erode, st); break;
    default: set_error("mask dtype %d (0 = uint8, 1 = int32, 2 = int64)", dtype); return 2;
  }
  return check_launch("morph");
}
int components(const uint8_t* m, int N, int H, int W, const Work& w, cudaStream_t st) {
  const int64_t total = (int64_t)N * H * W;
  launch_k(ccl_init_kernel, dim3(ew_grid(total)), dim3(kT), 0, st, m, total, w.L, w.area, w.key);
  if (int rc = check_launch("ccl_init")) return rc;
  launch_k(ccl_union_kernel, dim3(ew_grid(total)), dim3(kT), 0, st, m, N, H, W, w.L);
  if (int rc = check_launch("ccl_union")) return rc;
  launch_k(ccl_stats_kernel, dim3(ew_grid(total)), dim3(kT), 0, st, m, N, H, W, w.L, w.area, w.key);
  return check_launch("ccl_stats");
}

}  // namespace
}  // namespace iswm

using namespace iswm;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int64_t iswm_mask_work_bytes(int N, int H, int W) {
  const int64_t total = (int64_t)N * H * W;
  return 2 * align256(total) + 3 * align256(total * 4);
}

extern "C" int iswm_mask_preprocess(const void* d_mask, int dtype, int N, int H, int W, double min_valid_area, uint8_t* d_support,
                                    int32_t* d_front, int32_t* d_info, void* d_work, void* stream) {
  ISWM_REQUIRE(d_mask && d_support && d_front && d_info && d_work, "mask_preprocess: null");
  ISWM_REQUIRE(N >= 0 && H >= 1 && W >= 1 && (int64_t)N * H * W < (1ll << 31), "mask_preprocess: N*H*W must fit 31 bits");
  if (N == 0) return 0;
  const Work w = carve(d_work, (int64_t)N * H * W);
  // close = dilate, erode; open = erode, dilate: dilate 1, erode 2 (two box erosions compose), dilate 1
  if (int rc = morph_any(d_mask, dtype, w.a, N, H, W, 1, 0, ST(stream))) return rc;
  if (int rc = morph_any(w.a, ISWM_U8, w.b, N, H, W, 2, 1, ST(stream))) return rc;
  if (int rc = morph_any(w.b, ISWM_U8, w.a, N, H, W, 1, 0, ST(stream))) return rc;
  if (int rc = components(w.a, N, H, W, w, ST(stream))) return rc;
  launch_k(ccl_select_kernel, dim3((unsigned)N), dim3(kSel), 0, ST(stream), (const uint8_t*)w.a, H, W, (const int*)w.L, (const int*)w.area,
           (const int*)w.key, min_valid_area, d_info);
  if (int rc = check_launch("ccl_select")) return rc;
  launch_k(support_front_kernel, dim3((unsigned)(N * H)), dim3(128), 0, ST(stream), (const uint8_t*)w.a, H, W, (const int*)w.L, d_info, d_support, d_front);
  return check_launch("support_front");
}

extern "C" int iswm_region_components(const void* d_pred, int pred_dtype, const void* d_gt, int gt_dtype, int N, int H, int W, int min_area,
                                      int cap, int32_t* d_counts, int32_t* d_areas, void* d_work, void* stream) {
  ISWM_REQUIRE(d_pred && d_gt && d_counts && d_areas && d_work, "region_components: null");
  ISWM_REQUIRE(N >= 0 && H >= 1 && W >= 1 && cap >= 1 && (int64_t)N * H * W < (1ll << 31), "region_components: bad sizes");
  ISWM_REQUIRE(pred_dtype == gt_dtype, "region_components: prediction and ground truth must share a dtype (%d vs %d)", pred_dtype, gt_dtype);
  if (N == 0) return 0;
  const Work w = carve(d_work, (int64_t)N * H * W);
  if (cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * 8 * N, ST(stream)) != cudaSuccess) { set_error("region_components: memset"); return 1; }
  if (int rc = morph_any(d_pred, pred_dtype, w.b, N, H, W, 3, 0, ST(stream))) return rc;      // dilate, 3 iterations of a 3x3 box
  if (int rc = morph_any(w.b, ISWM_U8, w.a, N, H, W, 2, 1, ST(stream))) return rc;            // erode, 2 iterations
  dim3 grid((unsigned)std::max(1, std::min(64, (H * W + kT - 1) / kT)), (unsigned)N);
  switch (pred_dtype) {
    case ISWM_U8: launch_k(region_counts_kernel<uint8_t, uint8_t>, grid, dim3(kT), 0, ST(stream), (const uint8_t*)d_pred, (const uint8_t*)d_gt, (const uint8_t*)w.a, H * W, d_counts); break;
    case ISWM_I32: launch_k(region_counts_kernel<int32_t, int32_t>, grid, dim3(kT), 0, ST(stream), (const int32_t*)d_pred, (const int32_t*)d_gt, (const uint8_t*)w.a, H * W, d_counts); break;
    case ISWM_I64: launch_k(region_counts_kernel<int64_t, int64_t>, grid, dim3(kT), 0, ST(stream), (const int64_t*)d_pred, (const int64_t*)d_gt, (const uint8_t*)w.a, H * W, d_counts); break;
    default: set_error("region_components: dtype %d", pred_dtype); return 2;
  }
  if (int rc = check_launch("region_counts")) return rc;
  if (int rc = components(w.a, N, H, W, w, ST(stream))) return rc;
  launch_k(region_list_kernel, dim3((unsigned)N), dim3(kSel), 0, ST(stream), (const uint8_t*)w.a, H * W, (const int*)w.L, (const int*)w.area, min_area, cap,
           d_counts, d_areas);
  return check_launch("region_list");
}

extern "C" int iswm_front_nearest(const int32_t* d_front_a, const int32_t* d_front_b, int N, int H, int32_t* d_d2, int32_t* d_dx, void* stream) {
  ISWM_REQUIRE(d_front_a && d_front_b && d_d2 && d_dx, "front_nearest: null");
  ISWM_REQUIRE(N >= 0 && H >= 1 && H <= 8192, "front_nearest: H=%d (1..8192 rows)", H);
  if (N == 0) return 0;
  launch_k(front_nearest_kernel, dim3((unsigned)N), dim3(kT), (size_t)H * sizeof(int), ST(stream), d_front_a, d_front_b, H, d_d2, d_dx);
  return check_launch("front_nearest");
}

extern "C" int iswm_front_window_diff(const int32_t* d_front, const uint8_t* d_other, int N, int H, int W, int window, int32_t* d_diff, void* stream) {
  ISWM_REQUIRE(d_front && d_other && d_diff, "front_window_diff: null");
  ISWM_REQUIRE(N >= 0 && H >= 1 && W >= 1 && window >= 0 && (int64_t)N * H < (1ll << 31), "front_window_diff: bad sizes");
  if (N == 0) return 0;
  launch_k(front_window_kernel, dim3((unsigned)std::max(1, std::min((N * H + kT - 1) / kT, num_sms() * 8))), dim3(kT), 0, ST(stream), d_front, d_other, N, H,
           W, window, d_diff);
  return check_launch("front_window_diff");
}
