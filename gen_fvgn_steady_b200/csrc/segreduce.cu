// Deterministic CSR segmented reductions over the mesh graph (replace torch_scatter atomics).
//
// Reference semantics:
//   agg_i = sum_{j in Adj(i)} x_j                     blocks.py:92-99   (scatter_add of x[cat(r,s)] by cat(s,r))
//   a1_i  = sum_{f: s_f=i} e'_f[0:H/2] + sum_{f: r_f=i} e'_f[H/2:H]   blocks.py:24-42
//   a2_i  = (1/deg_i) sum_{j in Adj(i)} a1_j          blocks.py:44-51   (scatter_mean, deg clamped >= 1)
// The CSR lists are the *stable* grouping of the reference's scatter entry order by destination,
// so every row is summed in exactly the order a sequential index_add_ would use: results are
// bit-reproducible and equal to the CPU reference's fp32 sums.
// Backward of each is the same kernel (Adj is symmetric; the incidence transpose is a gather that
// the MLP backward prologue performs).
//
// Mapping: W/4 lanes per row, one float4 per lane (a 512-B row is one fully coalesced warp access).
#include "common.cuh"

// element-type helpers: 4 consecutive elements per lane, fp32 (16 B) or bf16 (8 B)
#ifndef FVGN_EMU
#include <cuda_bf16.h>
#endif
struct T_F32 { typedef float elem; };
struct T_BF16 { typedef uint16_t elem; };
template <class T> __device__ __forceinline__ float4 ldv(const typename T::elem* p);
template <> __device__ __forceinline__ float4 ldv<T_F32>(const float* p) { return ld4(p); }
template <class T> __device__ __forceinline__ void stv(typename T::elem* p, float4 v);
template <> __device__ __forceinline__ void stv<T_F32>(float* p, float4 v) { st4(p, v); }
#ifndef FVGN_EMU
template <> __device__ __forceinline__ float4 ldv<T_BF16>(const uint16_t* p) {
  const uint2 w = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
                     __uint_as_float(w.y & 0xFFFF0000u));
}
template <> __device__ __forceinline__ void stv<T_BF16>(uint16_t* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}
#endif

template <int W, class TS, class TD>
__global__ void __launch_bounds__(256) adj_reduce_kernel(const typename TS::elem* __restrict__ src, const int32_t* __restrict__ ptr,
                                                         const int32_t* __restrict__ nbr, typename TD::elem* __restrict__ dst,
                                                         int64_t n, int flags) {
  constexpr int LPR = W / 4;
  constexpr int RPB = 256 / LPR;
  const int64_t row = (int64_t)blockIdx.x * RPB + threadIdx.x / LPR;
  const int lane = threadIdx.x % LPR;
  if (row >= n) return;
  const int beg = ptr[row], end = ptr[row + 1];
  const bool div_src = flags & FVGN_ADJ_DIV_SRC_BY_DEG;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int t = beg;
  for (; t + 4 <= end; t += 4) {
    int j[4];
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) j[u] = nbr[t + u];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ldv<TS>(src + (size_t)j[u] * W + lane * 4);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (div_src) {
        const float d = (float)max(ptr[j[u] + 1] - ptr[j[u]], 1);
        v[u].x /= d; v[u].y /= d; v[u].z /= d; v[u].w /= d;
      }
      acc = add4(acc, v[u]);
    }
  }
  for (; t < end; ++t) {
    const int j = nbr[t];
    float4 v = ldv<TS>(src + (size_t)j * W + lane * 4);
    if (div_src) {
      const float d = (float)max(ptr[j + 1] - ptr[j], 1);
      v.x /= d; v.y /= d; v.z /= d; v.w /= d;
    }
    acc = add4(acc, v);
  }
  if (flags & FVGN_ADJ_DIV_DST_BY_DEG) {
    const float d = (float)max(end - beg, 1);
    acc.x /= d; acc.y /= d; acc.z /= d; acc.w /= d;
  }
  typename TD::elem* o = dst + (size_t)row * W + lane * 4;
  if (flags & FVGN_ADJ_ACCUMULATE) acc = add4(ldv<TD>(o), acc);
  stv<TD>(o, acc);
}

// dst[i, 0:W] = sum over incidence entries (edge f, role) of src[f, role*W : role*W + W]
template <int W, class TS, class TD>
__global__ void __launch_bounds__(256) inc_reduce_kernel(const typename TS::elem* __restrict__ src, const int32_t* __restrict__ ptr,
                                                         const int32_t* __restrict__ code, typename TD::elem* __restrict__ dst,
                                                         int64_t n) {
  constexpr int LPR = W / 4;
  constexpr int RPB = 256 / LPR;
  const int64_t row = (int64_t)blockIdx.x * RPB + threadIdx.x / LPR;
  const int lane = threadIdx.x % LPR;
  if (row >= n) return;
  const int beg = ptr[row], end = ptr[row + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int t = beg;
  for (; t + 4 <= end; t += 4) {
    int c[4];
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) c[u] = code[t + u];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ldv<TS>(src + (size_t)(c[u] >> 1) * (2 * W) + (c[u] & 1) * W + lane * 4);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc = add4(acc, v[u]);
  }
  for (; t < end; ++t) {
    const int c = code[t];
    acc = add4(acc, ldv<TS>(src + (size_t)(c >> 1) * (2 * W) + (c & 1) * W + lane * 4));
  }
  stv<TD>(dst + (size_t)row * W + lane * 4, acc);
}

#ifndef FVGN_EMU
// ------------------------------------------------------------------------------------------ pipelined kernels
// One warp per row, grid-stride over rows, software-pipelined three deep: while the gathers of row i are in flight the
// index entries of row i+1 and the row pointers of row i+2 are being fetched, so a warp pays ONE exposed memory latency
// per row instead of three (ptr -> entries -> rows).  Entries are accumulated strictly in CSR order: results are
// bit-identical to the one-shot kernels above.
template <int V> struct VecF;  // V fp32 accumulators per lane
template <> struct VecF<4> { float4 v; };
template <> struct VecF<2> { float2 v; };

template <class T, int V> struct LaneIO;
template <> struct LaneIO<T_F32, 4> {
  typedef float4 raw;
  static __device__ __forceinline__ raw ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ float4 up(raw r) { return r; }
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct LaneIO<T_BF16, 4> {
  typedef uint2 raw;
  static __device__ __forceinline__ raw ld(const uint16_t* p) { return *reinterpret_cast<const uint2*>(p); }
  static __device__ __forceinline__ float4 up(raw w) {
    return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
                       __uint_as_float(w.y & 0xFFFF0000u));
  }
  static __device__ __forceinline__ void st(uint16_t* p, float4 v) { stv<T_BF16>(p, v); }
};
template <> struct LaneIO<T_F32, 2> {
  typedef float2 raw;
  static __device__ __forceinline__ raw ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ float4 up(raw r) { return make_float4(r.x, r.y, 0.f, 0.f); }
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y); }
};
template <> struct LaneIO<T_BF16, 2> {
  typedef uint32_t raw;
  static __device__ __forceinline__ raw ld(const uint16_t* p) { return *reinterpret_cast<const uint32_t*>(p); }
  static __device__ __forceinline__ float4 up(raw w) { return make_float4(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u), 0.f, 0.f); }
  static __device__ __forceinline__ void st(uint16_t* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&a);
  }
};

// INC = false: entries are neighbour rows (src row = entry, W columns);  INC = true: entries are edge*2+role codes
// (src row = code >> 1 of a [E, 2W] array, column block = code & 1).
template <int W, class TS, class TD, bool INC>
__global__ void __launch_bounds__(256) pipe_reduce_kernel(const typename TS::elem* __restrict__ src, const int32_t* __restrict__ ptr,
                                                          const int32_t* __restrict__ ent, typename TD::elem* __restrict__ dst,
                                                          int64_t n, int flags) {
  // 512-B rows: one warp per row (32 lanes x 4 elements); 256-B rows: one HALF-warp per row (16 lanes x 4 elements), the
  // two halves of a warp run independent rows (shuffles are confined to the half by mask + width)
  constexpr int LPR = W / 4;   // lanes per row: 32 or 16
  constexpr int V = 4;         // elements per lane
  constexpr int NB = 4;        // gathers in flight per lane
  constexpr int ROW_LD = INC ? 2 * W : W;
  typedef LaneIO<TS, V> SI;
  typedef LaneIO<TD, V> DI;
  const int lane = threadIdx.x & (LPR - 1);
  const unsigned hmask = (LPR == 32) ? 0xffffffffu : (0xffffu << (threadIdx.x & 16));
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t GW = (int64_t)gridDim.x * blockDim.x / LPR;
  const bool div_src = !INC && (flags & FVGN_ADJ_DIV_SRC_BY_DEG);
  auto load_ptr = [&](int64_t row, int& b, int& e) {
    b = 0; e = 0;
    if (row < n) { b = __ldg(ptr + row); e = __ldg(ptr + row + 1); }
  };
  auto load_ent = [&](int b, int e) { return (b + lane < e) ? __ldg(ent + b + lane) : 0; };  // first LPR entries, one per lane
  int cb, ce, cent;  // row i: pointers and entries (arrived)
  int nb_, ne_;      // row i+1: pointers (arrived), entries being fetched into nent
  load_ptr(gw, cb, ce);
  load_ptr(gw + GW, nb_, ne_);
  cent = load_ent(cb, ce);
  for (int64_t row = gw; row < n; row += GW) {
    const int deg = ce - cb;
    // ---- issue: gathers of this row (first NB), entries of the next row, pointers of the one after
    typename SI::raw v[NB];
    int code[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      code[k] = __shfl_sync(hmask, cent, k, LPR);
      if (k < deg) {
        const int r = INC ? (code[k] >> 1) : code[k];
        const int coff = INC ? (code[k] & 1) * W : 0;
        v[k] = SI::ld(src + (size_t)r * ROW_LD + coff + lane * V);
      }
    }
    float mydiv = 1.f;  // lane k: degree of the source row of entry k (transposed mean)
    if (div_src && lane < deg) mydiv = (float)max(__ldg(ptr + cent + 1) - __ldg(ptr + cent), 1);
    const int nent = load_ent(nb_, ne_);
    int ab, ae;
    load_ptr(row + 2 * GW, ab, ae);
    float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
    typename TD::elem* o = dst + (size_t)row * W + lane * V;
    if (!INC && (flags & FVGN_ADJ_ACCUMULATE)) old = DI::up(DI::ld(o));
    // ---- consume in CSR order
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      if (k < deg) {
        float4 x = SI::up(v[k]);
        if (div_src) {
          const float dv = __shfl_sync(hmask, mydiv, k, LPR);
          x.x /= dv; x.y /= dv; x.z /= dv; x.w /= dv;
        }
        acc = add4(acc, x);
      }
    }
    for (int t = NB; t < deg; ++t) {  // long rows: the rest, one at a time (entries beyond LPR straight from memory)
      const int c = (t < LPR) ? __shfl_sync(hmask, cent, t, LPR) : __ldg(ent + cb + t);
      const int r = INC ? (c >> 1) : c;
      const int coff = INC ? (c & 1) * W : 0;
      float4 x = SI::up(SI::ld(src + (size_t)r * ROW_LD + coff + lane * V));
      if (div_src) {
        const float dv = (float)max(__ldg(ptr + c + 1) - __ldg(ptr + c), 1);
        x.x /= dv; x.y /= dv; x.z /= dv; x.w /= dv;
      }
      acc = add4(acc, x);
    }
    if (!INC && (flags & FVGN_ADJ_DIV_DST_BY_DEG)) {
      const float dd = (float)max(deg, 1);
      acc.x /= dd; acc.y /= dd; acc.z /= dd; acc.w /= dd;
    }
    if (!INC && (flags & FVGN_ADJ_ACCUMULATE)) acc = add4(old, acc);
    DI::st(o, acc);
    // ---- rotate the pipeline
    cb = nb_; ce = ne_; cent = nent;
    nb_ = ab; ne_ = ae;
  }
}
#endif  // FVGN_EMU

#ifndef FVGN_EMU
// persistent grid: as many 256-thread blocks as are resident at once (fewer for small inputs)
template <class K>
static unsigned pipe_grid(K kern, int64_t n_rows) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
  if (sms <= 0) sms = 148;
  const int64_t want = (n_rows + 7) / 8;
  const int64_t cap = (int64_t)sms * per_sm;
  return (unsigned)(want < cap ? want : cap);
}
template <int W, class TS, class TD, bool INC>
static void launch_pipe(const typename TS::elem* s, const int32_t* ptr, const int32_t* ent, typename TD::elem* o, int64_t n_rows,
                        int flags, void* stream) {
  auto kern = pipe_reduce_kernel<W, TS, TD, INC>;
  static unsigned cap_grid = 0;  // per instantiation
  if (cap_grid == 0) cap_grid = pipe_grid(kern, (int64_t)1 << 40);
  const int rows_per_block = 256 / (W / 4);
  const int64_t want = (n_rows + rows_per_block - 1) / rows_per_block;
  const unsigned grid = (unsigned)(want < cap_grid ? want : cap_grid);
  kern<<<grid, 256, 0, (cudaStream_t)stream>>>(s, ptr, ent, o, n_rows, flags);
}
#endif

template <class TS, class TD>
static int launch_adj(const void* src, const int32_t* ptr, const int32_t* nbr, void* dst, int64_t n_rows, int32_t width,
                      int32_t flags, void* stream) {
  const typename TS::elem* s = reinterpret_cast<const typename TS::elem*>(src);
  typename TD::elem* o = reinterpret_cast<typename TD::elem*>(dst);
#ifndef FVGN_EMU
  if (!(flags & FVGN_ADJ_SIMPLE_KERNEL) && (width == 128 || width == 64)) {
    if (width == 128) launch_pipe<128, TS, TD, false>(s, ptr, nbr, o, n_rows, flags, stream);
    else launch_pipe<64, TS, TD, false>(s, ptr, nbr, o, n_rows, flags, stream);
    FVGN_CHECK_LAUNCH();
    return FVGN_OK;
  }
#endif
  if (width == 128) {
    const unsigned grid = (unsigned)((n_rows + 7) / 8);
    auto kern = adj_reduce_kernel<128, TS, TD>;
    FVGN_LAUNCH_SEQ(kern, grid, 256, 0, stream, s, ptr, nbr, o, n_rows, flags);
  } else if (width == 64) {
    const unsigned grid = (unsigned)((n_rows + 15) / 16);
    auto kern = adj_reduce_kernel<64, TS, TD>;
    FVGN_LAUNCH_SEQ(kern, grid, 256, 0, stream, s, ptr, nbr, o, n_rows, flags);
  } else {
    return FVGN_ERR_UNSUPPORTED;
  }
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

template <class TS, class TD>
static int launch_inc(const void* src, const int32_t* ptr, const int32_t* code, void* dst, int64_t n_rows, int32_t width,
                      void* stream) {
  const typename TS::elem* s = reinterpret_cast<const typename TS::elem*>(src);
  typename TD::elem* o = reinterpret_cast<typename TD::elem*>(dst);
#ifndef FVGN_EMU
  if (width == 128 || width == 64) {
    if (width == 128) launch_pipe<128, TS, TD, true>(s, ptr, code, o, n_rows, 0, stream);
    else launch_pipe<64, TS, TD, true>(s, ptr, code, o, n_rows, 0, stream);
    FVGN_CHECK_LAUNCH();
    return FVGN_OK;
  }
#endif
  if (width == 128) {
    const unsigned grid = (unsigned)((n_rows + 7) / 8);
    auto kern = inc_reduce_kernel<128, TS, TD>;
    FVGN_LAUNCH_SEQ(kern, grid, 256, 0, stream, s, ptr, code, o, n_rows);
  } else if (width == 64) {
    const unsigned grid = (unsigned)((n_rows + 15) / 16);
    auto kern = inc_reduce_kernel<64, TS, TD>;
    FVGN_LAUNCH_SEQ(kern, grid, 256, 0, stream, s, ptr, code, o, n_rows);
  } else {
    return FVGN_ERR_UNSUPPORTED;
  }
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_adj_reduce(const float* src, const int32_t* ptr, const int32_t* nbr, float* dst, int64_t n_rows,
                               int32_t width, int32_t flags, void* stream) {
  if (n_rows < 0) return FVGN_ERR_SHAPE;
  if (n_rows == 0) return FVGN_OK;
  if (!fvgn_aligned16(src) || !fvgn_aligned16(dst)) return FVGN_ERR_ALIGN;
  return launch_adj<T_F32, T_F32>(src, ptr, nbr, dst, n_rows, width, flags, stream);
}

extern "C" int fvgn_inc_reduce(const float* src, const int32_t* ptr, const int32_t* code, float* dst, int64_t n_rows,
                               int32_t width, void* stream) {
  if (n_rows < 0) return FVGN_ERR_SHAPE;
  if (n_rows == 0) return FVGN_OK;
  if (!fvgn_aligned16(src) || !fvgn_aligned16(dst)) return FVGN_ERR_ALIGN;
  return launch_inc<T_F32, T_F32>(src, ptr, code, dst, n_rows, width, stream);
}

extern "C" int fvgn_adj_reduce_t(const void* src, int32_t src_type, const int32_t* ptr, const int32_t* nbr, void* dst,
                                 int32_t dst_type, int64_t n_rows, int32_t width, int32_t flags, void* stream) {
  if (n_rows < 0) return FVGN_ERR_SHAPE;
  if (n_rows == 0) return FVGN_OK;
  if (!fvgn_aligned16(src) || !fvgn_aligned16(dst)) return FVGN_ERR_ALIGN;
#ifndef FVGN_EMU
  if (src_type == FVGN_T_F32 && dst_type == FVGN_T_BF16) return launch_adj<T_F32, T_BF16>(src, ptr, nbr, dst, n_rows, width, flags, stream);
  if (src_type == FVGN_T_BF16 && dst_type == FVGN_T_F32) return launch_adj<T_BF16, T_F32>(src, ptr, nbr, dst, n_rows, width, flags, stream);
  if (src_type == FVGN_T_BF16 && dst_type == FVGN_T_BF16) return launch_adj<T_BF16, T_BF16>(src, ptr, nbr, dst, n_rows, width, flags, stream);
#endif
  if (src_type == FVGN_T_F32 && dst_type == FVGN_T_F32) return launch_adj<T_F32, T_F32>(src, ptr, nbr, dst, n_rows, width, flags, stream);
  return FVGN_ERR_UNSUPPORTED;
}

extern "C" int fvgn_inc_reduce_t(const void* src, int32_t src_type, const int32_t* ptr, const int32_t* code, void* dst,
                                 int32_t dst_type, int64_t n_rows, int32_t width, void* stream) {
  if (n_rows < 0) return FVGN_ERR_SHAPE;
  if (n_rows == 0) return FVGN_OK;
  if (!fvgn_aligned16(src) || !fvgn_aligned16(dst)) return FVGN_ERR_ALIGN;
#ifndef FVGN_EMU
  if (src_type == FVGN_T_F32 && dst_type == FVGN_T_BF16) return launch_inc<T_F32, T_BF16>(src, ptr, code, dst, n_rows, width, stream);
  if (src_type == FVGN_T_BF16 && dst_type == FVGN_T_F32) return launch_inc<T_BF16, T_F32>(src, ptr, code, dst, n_rows, width, stream);
  if (src_type == FVGN_T_BF16 && dst_type == FVGN_T_BF16) return launch_inc<T_BF16, T_BF16>(src, ptr, code, dst, n_rows, width, stream);
#endif
  if (src_type == FVGN_T_F32 && dst_type == FVGN_T_F32) return launch_inc<T_F32, T_F32>(src, ptr, code, dst, n_rows, width, stream);
  return FVGN_ERR_UNSUPPORTED;
}
