// Plan-time kernels: the stable grouping of a scatter entry list by destination -- the CSR
// every deterministic reduction of the hot path runs on -- built on the device without a library sort, a topology hash
// for the plan cache, and a generic deterministic CSR gather-sum for the stand-alone finite-volume API.
//
// Reference semantics: torch_scatter / index_add_ visit the entries of the loader's index tensors in their storage order
// (src/FVMmodel/Models/FVGN/blocks.py:24-51,84-99; FVdiscretization/FVgrad.py:264-325; utils/utilities.py:16-61).  The
// stable CSR lists every destination's entries in exactly that order, so fp32 row sums equal a sequential index_add_.
//
//   fvgn_csr_build : counts (integer atomics: order-free), exclusive scan, unordered fill (atomic cursors), then every row
//                    sorts its own few entries by entry number -- ascending entry number IS the stable order, whatever order
//                    the atomics ran in, so the result is deterministic.
#include "common.cuh"

#ifdef FVGN_EMU
// tests/emu build (host tensors, test infrastructure): the same entry points as plain sequential loops
#include <vector>
extern "C" int64_t fvgn_csr_build_workspace_bytes(int64_t n_rows) { return (n_rows + 2) * 4; }
extern "C" int fvgn_csr_build(const void* dest, int32_t dest_is_int64, int64_t m, int64_t n, int32_t* ptr, int32_t* perm, void*, void*) {
  auto at = [&](int64_t i) { return dest_is_int64 ? ((const int64_t*)dest)[i] : (int64_t)((const int32_t*)dest)[i]; };
  for (int64_t r = 0; r <= n; ++r) ptr[r] = 0;
  for (int64_t i = 0; i < m; ++i) if (at(i) >= 0 && at(i) < n) ptr[at(i) + 1]++;
  for (int64_t r = 0; r < n; ++r) ptr[r + 1] += ptr[r];
  std::vector<int32_t> cur(ptr, ptr + n);
  for (int64_t i = 0; i < m; ++i) if (at(i) >= 0 && at(i) < n) perm[cur[at(i)]++] = (int32_t)i;
  return FVGN_OK;
}
extern "C" int fvgn_hash_words(const void* data, int64_t n_words, int64_t seed, void* out_u64, void*) {
  uint64_t acc = 0;
  for (int64_t i = 0; i < n_words; ++i) acc += ((uint64_t)((const uint32_t*)data)[i] + 0x9e3779b97f4a7c15ULL * (uint64_t)(i + seed)) * 0xff51afd7ed558ccdULL;
  *(uint64_t*)out_u64 += acc;
  return FVGN_OK;
}
template <class T>
static int emu_wsum(const T* src, int32_t width, int32_t ld, const int32_t* ptr, const int32_t* idx, const T* w, int32_t mode, T* dst,
                    int64_t n_rows) {
  for (int64_t row = 0; row < n_rows; ++row)
    for (int col = 0; col < width; ++col) {
      T acc = 0, wsum = 0;
      for (int k = ptr[row]; k < ptr[row + 1]; ++k) {
        const int64_t j = idx ? idx[k] : k;
        T v = src[j * ld + col];
        if (w) { v = v * w[k]; wsum = wsum + w[k]; }
        acc = acc + v;
      }
      const int cnt = ptr[row + 1] - ptr[row];
      if (mode == 1) acc = acc / (T)(cnt > 1 ? cnt : 1);
      if (mode == 2) acc = acc / wsum;
      dst[row * (int64_t)width + col] = acc;
    }
  return FVGN_OK;
}
extern "C" int fvgn_csr_weighted_sum(const float* src, int32_t width, int32_t ld, const int32_t* ptr, const int32_t* idx, const float* w,
                                     int32_t mode, float* dst, int64_t n_rows, void*) {
  return emu_wsum<float>(src, width, ld, ptr, idx, w, mode, dst, n_rows);
}
extern "C" int fvgn_csr_weighted_sum_f64(const void* src, int32_t width, int32_t ld, const int32_t* ptr, const int32_t* idx, const void* w,
                                         int32_t mode, void* dst, int64_t n_rows, void*) {
  return emu_wsum<double>((const double*)src, width, ld, ptr, idx, (const double*)w, mode, (double*)dst, n_rows);
}
#else

namespace {

constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_TILE = SCAN_ITEMS * SCAN_BLOCK;

template <class IT>
__global__ void __launch_bounds__(256) csr_count_kernel(const IT* __restrict__ dest, int64_t m, int64_t n, int32_t* __restrict__ ptr,
                                                        int32_t* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = (int64_t)dest[i];
    if (d < 0 || d >= n) { *bad = 1; continue; }
    atomicAdd(ptr + d + 1, 1);
  }
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// inclusive scan of one SCAN_TILE-element tile held SCAN_ITEMS per thread; returns the tile total in every thread
__device__ __forceinline__ int block_scan_tile(int (&x)[SCAN_ITEMS], int* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int run = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) { run += x[j]; x[j] = run; }
  const int incl = warp_incl_scan(run, lane);
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = lane < SCAN_BLOCK / 32 ? s_warp[lane] : 0;
    const int ws = warp_incl_scan(w, lane);
    if (lane < SCAN_BLOCK / 32) s_warp[lane] = ws;
  }
  __syncthreads();
  const int base = (incl - run) + (warp > 0 ? s_warp[warp - 1] : 0);
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) x[j] += base;
  const int total = s_warp[SCAN_BLOCK / 32 - 1];
  __syncthreads();
  return total;
}

// phase 1: per-tile totals; phase 3 (apply = true): in-place inclusive scan of each tile plus its tile offset
template <bool APPLY>
__global__ void __launch_bounds__(SCAN_BLOCK) scan_tiles_kernel(int32_t* __restrict__ a, int64_t n, int32_t* __restrict__ tile_sums) {
  __shared__ int s_warp[SCAN_BLOCK / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int x[SCAN_ITEMS];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) x[j] = (base + j < n) ? a[base + j] : 0;
  const int total = block_scan_tile(x, s_warp);
  if (!APPLY) {
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
  } else {
    const int off = blockIdx.x > 0 ? tile_sums[blockIdx.x - 1] : 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j)
      if (base + j < n) a[base + j] = x[j] + off;
  }
}

// phase 2: inclusive scan of the tile totals by ONE block (serial over tiles of SCAN_TILE)
__global__ void __launch_bounds__(SCAN_BLOCK) scan_sums_kernel(int32_t* __restrict__ sums, int64_t nt) {
  __shared__ int s_warp[SCAN_BLOCK / 32];
  int carry = 0;
  for (int64_t t0 = 0; t0 < nt; t0 += SCAN_TILE) {
    const int64_t base = t0 + (int64_t)threadIdx.x * SCAN_ITEMS;
    int x[SCAN_ITEMS];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) x[j] = (base + j < nt) ? sums[base + j] : 0;
    const int total = block_scan_tile(x, s_warp);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j)
      if (base + j < nt) sums[base + j] = x[j] + carry;
    carry += total;
  }
}

template <class IT>
__global__ void __launch_bounds__(256) csr_fill_kernel(const IT* __restrict__ dest, int64_t m, int64_t n, const int32_t* __restrict__ ptr,
                                                       int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = (int64_t)dest[i];
    if (d < 0 || d >= n) continue;
    const int pos = atomicAdd(cursor + d, 1);
    perm[ptr[d] + pos] = (int32_t)i;
  }
}

// each row sorts its segment of entry numbers ascending (insertion sort for the usual short rows, heap sort beyond)
__global__ void __launch_bounds__(256) csr_sort_rows_kernel(const int32_t* __restrict__ ptr, int64_t n, int32_t* __restrict__ perm) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const int b = ptr[r], len = ptr[r + 1] - b;
    int32_t* a = perm + b;
    if (len <= 24) {
      for (int i = 1; i < len; ++i) {
        const int32_t v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
        a[j + 1] = v;
      }
    } else {
      auto sift = [&](int root, int end) {
        for (;;) {
          int c = 2 * root + 1;
          if (c >= end) break;
          if (c + 1 < end && a[c + 1] > a[c]) ++c;
          if (a[root] >= a[c]) break;
          const int32_t t = a[root]; a[root] = a[c]; a[c] = t;
          root = c;
        }
      };
      for (int s = len / 2 - 1; s >= 0; --s) sift(s, len);
      for (int e = len - 1; e > 0; --e) {
        const int32_t t = a[0]; a[0] = a[e]; a[e] = t;
        sift(0, e);
      }
    }
  }
}

// 64-bit order-sensitive content hash: sum over words of mix(word, position); integer adds commute, so the atomics'
// order does not matter
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__global__ void __launch_bounds__(256) hash_words_kernel(const uint32_t* __restrict__ w, int64_t n, uint64_t seed,
                                                         unsigned long long* __restrict__ out) {
  uint64_t acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += mix64(((uint64_t)w[i] << 32 | (uint64_t)(uint32_t)i) ^ mix64(seed + (uint64_t)(i >> 32)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc != 0) atomicAdd(out, (unsigned long long)acc);
}

// dst[i, c] = sum_{t in row i} w[t] * src[idx[t], c]   (products rounded before they are added, entries in CSR order:
// the fp32 result of a sequential index_add_ of (src * w)); mode 1: / max(count, 1); mode 2: / sum of w
template <class T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <class T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

template <class T>
__global__ void __launch_bounds__(256) csr_wsum_kernel(const T* __restrict__ src, int width, int ld, const int32_t* __restrict__ ptr,
                                                       const int32_t* __restrict__ idx, const T* __restrict__ w, int mode,
                                                       T* __restrict__ dst, int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t / width;
  const int col = (int)(t - row * width);
  if (row >= n) return;
  const int b = ptr[row], e = ptr[row + 1];
  T acc = 0, wsum = 0;
  for (int k = b; k < e; ++k) {
    const int64_t j = idx ? (int64_t)idx[k] : (int64_t)k;
    T v = src[j * ld + col];
    if (w) {
      const T ww = w[k];
      v = mul_rn<T>(v, ww);
      wsum = add_rn<T>(wsum, ww);
    }
    acc = add_rn<T>(acc, v);
  }
  if (mode == 1) acc = acc / (T)max(e - b, 1);
  if (mode == 2) acc = acc / wsum;
  dst[row * (int64_t)width + col] = acc;
}

unsigned grid_for(int64_t items, int per_block) {
  const int64_t want = (items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)fvgn_num_sms() * 16;
  return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <class IT>
int csr_build(const IT* dest, int64_t m, int64_t n, int32_t* ptr, int32_t* perm, int32_t* ws, cudaStream_t st) {
  // ws: [n] cursors | [ceil((n+1)/SCAN_TILE)] tile sums | [1] error flag
  const int64_t nt = (n + 1 + SCAN_TILE - 1) / SCAN_TILE;
  int32_t* cursor = ws;
  int32_t* sums = ws + n;
  int32_t* bad = sums + nt;
  if (cudaMemsetAsync(ptr, 0, (size_t)(n + 1) * 4, st) != cudaSuccess) return FVGN_ERR_LAUNCH;
  if (cudaMemsetAsync(ws, 0, (size_t)(n + nt + 1) * 4, st) != cudaSuccess) return FVGN_ERR_LAUNCH;
  if (m > 0) {
    csr_count_kernel<IT><<<grid_for(m, 256), 256, 0, st>>>(dest, m, n, ptr, bad);
    FVGN_CHECK_LAUNCH();
  }
  scan_tiles_kernel<false><<<(unsigned)nt, SCAN_BLOCK, 0, st>>>(ptr, n + 1, sums);
  FVGN_CHECK_LAUNCH();
  scan_sums_kernel<<<1, SCAN_BLOCK, 0, st>>>(sums, nt);
  FVGN_CHECK_LAUNCH();
  scan_tiles_kernel<true><<<(unsigned)nt, SCAN_BLOCK, 0, st>>>(ptr, n + 1, sums);
  FVGN_CHECK_LAUNCH();
  if (m > 0) {
    csr_fill_kernel<IT><<<grid_for(m, 256), 256, 0, st>>>(dest, m, n, ptr, cursor, perm);
    FVGN_CHECK_LAUNCH();
    csr_sort_rows_kernel<<<grid_for(n, 256), 256, 0, st>>>(ptr, n, perm);
    FVGN_CHECK_LAUNCH();
  }
  return FVGN_OK;
}

}  // namespace

extern "C" int64_t fvgn_csr_build_workspace_bytes(int64_t n_rows) {
  const int64_t nt = (n_rows + 1 + SCAN_TILE - 1) / SCAN_TILE;
  return (n_rows + nt + 1) * 4;
}

extern "C" int fvgn_csr_build(const void* dest, int32_t dest_is_int64, int64_t n_entries, int64_t n_rows, int32_t* ptr, int32_t* perm,
                              void* workspace, void* stream) {
  if (n_entries < 0 || n_rows < 0 || n_entries >= ((int64_t)1 << 31)) return FVGN_ERR_SHAPE;
  if (!ptr || !workspace || (n_entries > 0 && (!dest || !perm))) return FVGN_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* ws = reinterpret_cast<int32_t*>(workspace);
  return dest_is_int64 ? csr_build<int64_t>(reinterpret_cast<const int64_t*>(dest), n_entries, n_rows, ptr, perm, ws, st)
                       : csr_build<int32_t>(reinterpret_cast<const int32_t*>(dest), n_entries, n_rows, ptr, perm, ws, st);
}

extern "C" int fvgn_hash_words(const void* data, int64_t n_words, int64_t seed, void* out_u64, void* stream) {
  if (n_words < 0) return FVGN_ERR_SHAPE;
  if (n_words == 0) return FVGN_OK;
  if (!data || !out_u64) return FVGN_ERR_NULL;
  hash_words_kernel<<<grid_for(n_words, 1024), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(data), n_words,
                                                                               (uint64_t)seed,
                                                                               reinterpret_cast<unsigned long long*>(out_u64));
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

template <class T>
static int csr_wsum(const T* src, int32_t width, int32_t ld, const int32_t* ptr, const int32_t* idx, const T* w, int32_t mode, T* dst,
                    int64_t n_rows, void* stream) {
  if (n_rows < 0 || width < 1 || ld < width || mode < 0 || mode > 2) return FVGN_ERR_SHAPE;
  if (n_rows == 0) return FVGN_OK;
  if (!src || !ptr || !dst || (mode == 2 && !w)) return FVGN_ERR_NULL;
  const int64_t threads = n_rows * width;
  csr_wsum_kernel<T><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, width, ld, ptr, idx, w, mode, dst, n_rows);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}
extern "C" int fvgn_csr_weighted_sum(const float* src, int32_t width, int32_t ld, const int32_t* ptr, const int32_t* idx, const float* w,
                                     int32_t mode, float* dst, int64_t n_rows, void* stream) {
  return csr_wsum<float>(src, width, ld, ptr, idx, w, mode, dst, n_rows, stream);
}
extern "C" int fvgn_csr_weighted_sum_f64(const void* src, int32_t width, int32_t ld, const int32_t* ptr, const int32_t* idx, const void* w,
                                         int32_t mode, void* dst, int64_t n_rows, void* stream) {
  return csr_wsum<double>((const double*)src, width, ld, ptr, idx, (const double*)w, mode, (double*)dst, n_rows, stream);
}
#endif  // FVGN_EMU
