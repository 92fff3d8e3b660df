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
#include <cuda_fp16.h>
#endif
struct T_F32 { typedef float elem; };
struct T_BF16 { typedef uint16_t elem; };
struct T_F16 { typedef uint16_t elem; };   // IEEE half (FVGN_PREC_F16 streams); same storage as bf16
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
__device__ __forceinline__ float hf_lo(uint32_t w) { return __low2float(*reinterpret_cast<const __half2*>(&w)); }
__device__ __forceinline__ float hf_hi(uint32_t w) { return __high2float(*reinterpret_cast<const __half2*>(&w)); }
__device__ __forceinline__ uint32_t hf_pack(float lo, float hi) {  // round to nearest, saturate to +-65504
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <> __device__ __forceinline__ float4 ldv<T_F16>(const uint16_t* p) {
  const uint2 w = *reinterpret_cast<const uint2*>(p);
  return make_float4(hf_lo(w.x), hf_hi(w.x), hf_lo(w.y), hf_hi(w.y));
}
template <> __device__ __forceinline__ void stv<T_F16>(uint16_t* p, float4 v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(hf_pack(v.x, v.y), hf_pack(v.z, v.w));
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
// per-lane vector I/O: V consecutive elements (fp32 or bf16), unpacked to / packed from fp32
template <class T, int V> struct LaneIO;
template <> struct LaneIO<T_F32, 4> {
  struct raw { float4 a; };
  static __device__ __forceinline__ raw ld(const float* p) { return raw{*reinterpret_cast<const float4*>(p)}; }
  static __device__ __forceinline__ void up(const raw& r, float (&x)[4]) { x[0] = r.a.x; x[1] = r.a.y; x[2] = r.a.z; x[3] = r.a.w; }
  static __device__ __forceinline__ void st(float* p, const float (&x)[4]) { *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]); }
};
template <> struct LaneIO<T_F32, 8> {
  struct raw { float4 a, b; };
  static __device__ __forceinline__ raw ld(const float* p) {
    return raw{*reinterpret_cast<const float4*>(p), *reinterpret_cast<const float4*>(p + 4)};
  }
  static __device__ __forceinline__ void up(const raw& r, float (&x)[8]) {
    x[0] = r.a.x; x[1] = r.a.y; x[2] = r.a.z; x[3] = r.a.w; x[4] = r.b.x; x[5] = r.b.y; x[6] = r.b.z; x[7] = r.b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&x)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
  }
};
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t bf_pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> struct LaneIO<T_BF16, 4> {
  struct raw { uint2 a; };
  static __device__ __forceinline__ raw ld(const uint16_t* p) { return raw{*reinterpret_cast<const uint2*>(p)}; }
  static __device__ __forceinline__ void up(const raw& r, float (&x)[4]) {
    x[0] = bf_lo(r.a.x); x[1] = bf_hi(r.a.x); x[2] = bf_lo(r.a.y); x[3] = bf_hi(r.a.y);
  }
  static __device__ __forceinline__ void st(uint16_t* p, const float (&x)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(bf_pack(x[0], x[1]), bf_pack(x[2], x[3]));
  }
};
template <> struct LaneIO<T_BF16, 8> {
  struct raw { uint4 a; };
  static __device__ __forceinline__ raw ld(const uint16_t* p) { return raw{*reinterpret_cast<const uint4*>(p)}; }
  static __device__ __forceinline__ void up(const raw& r, float (&x)[8]) {
    x[0] = bf_lo(r.a.x); x[1] = bf_hi(r.a.x); x[2] = bf_lo(r.a.y); x[3] = bf_hi(r.a.y);
    x[4] = bf_lo(r.a.z); x[5] = bf_hi(r.a.z); x[6] = bf_lo(r.a.w); x[7] = bf_hi(r.a.w);
  }
  static __device__ __forceinline__ void st(uint16_t* p, const float (&x)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(bf_pack(x[0], x[1]), bf_pack(x[2], x[3]), bf_pack(x[4], x[5]), bf_pack(x[6], x[7]));
  }
};

template <> struct LaneIO<T_F16, 4> {
  struct raw { uint2 a; };
  static __device__ __forceinline__ raw ld(const uint16_t* p) { return raw{*reinterpret_cast<const uint2*>(p)}; }
  static __device__ __forceinline__ void up(const raw& r, float (&x)[4]) {
    x[0] = hf_lo(r.a.x); x[1] = hf_hi(r.a.x); x[2] = hf_lo(r.a.y); x[3] = hf_hi(r.a.y);
  }
  static __device__ __forceinline__ void st(uint16_t* p, const float (&x)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(hf_pack(x[0], x[1]), hf_pack(x[2], x[3]));
  }
};
template <> struct LaneIO<T_F16, 8> {
  struct raw { uint4 a; };
  static __device__ __forceinline__ raw ld(const uint16_t* p) { return raw{*reinterpret_cast<const uint4*>(p)}; }
  static __device__ __forceinline__ void up(const raw& r, float (&x)[8]) {
    x[0] = hf_lo(r.a.x); x[1] = hf_hi(r.a.x); x[2] = hf_lo(r.a.y); x[3] = hf_hi(r.a.y);
    x[4] = hf_lo(r.a.z); x[5] = hf_hi(r.a.z); x[6] = hf_lo(r.a.w); x[7] = hf_hi(r.a.w);
  }
  static __device__ __forceinline__ void st(uint16_t* p, const float (&x)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(hf_pack(x[0], x[1]), hf_pack(x[2], x[3]), hf_pack(x[4], x[5]), hf_pack(x[6], x[7]));
  }
};

// acc += x on register pairs (FADD2: IEEE per element, half the fp32 instructions of the accumulate)
template <int V>
__device__ __forceinline__ void add_vec(float (&acc)[V], const float (&x)[V]) {
#pragma unroll
  for (int j = 0; j < V; j += 2) {
    const float2 r = __fadd2_rn(make_float2(acc[j], acc[j + 1]), make_float2(x[j], x[j + 1]));
    acc[j] = r.x;
    acc[j + 1] = r.y;
  }
}

// lanes per row: 16 whenever a lane can move 16 bytes of the SOURCE row (bf16 512-col... i.e. 128 bf16 = 16 x 16 B, or
// 64 fp32 = 16 x 16 B), else 32; the two half-warps of a 16-lane configuration run independent rows
template <int W, class TS> struct PipeCfg { static constexpr int LPR = (W * (int)sizeof(typename TS::elem) >= 512) ? 32 : 16; };

// INC = false: entries are neighbour rows (src row = entry, W columns);  INC = true: entries are edge*2+role codes
// (src row = code >> 1 of a [E, 2W] array, column block = code & 1).
template <int W, class TS, class TD, bool INC>
__global__ void __launch_bounds__(256) pipe_reduce_kernel(const typename TS::elem* __restrict__ src, const int32_t* __restrict__ ptr,
                                                          const int32_t* __restrict__ ent, typename TD::elem* __restrict__ dst,
                                                          int64_t n, int flags) {
  constexpr int LPR = PipeCfg<W, TS>::LPR;  // lanes per row: 32 (one warp per row) or 16 (one HALF-warp per row)
  constexpr int V = W / LPR;                // elements per lane: 4 or 8
  constexpr int NB = 4;                     // gathers in flight per lane
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
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const int code = __shfl_sync(hmask, cent, k, LPR);
      if (k < deg) {
        const int r = INC ? (code >> 1) : code;
        const int coff = INC ? (code & 1) * W : 0;
        v[k] = SI::ld(src + (size_t)r * ROW_LD + coff + lane * V);
      }
    }
    float mydiv = 1.f;  // lane k: degree of the source row of entry k (transposed mean)
    if (div_src && lane < deg) mydiv = (float)max(__ldg(ptr + cent + 1) - __ldg(ptr + cent), 1);
    const int nent = load_ent(nb_, ne_);
    int ab, ae;
    load_ptr(row + 2 * GW, ab, ae);
    float old[V];
#pragma unroll
    for (int j = 0; j < V; ++j) old[j] = 0.f;
    typename TD::elem* o = dst + (size_t)row * W + lane * V;
    if (!INC && (flags & FVGN_ADJ_ACCUMULATE)) DI::up(DI::ld(o), old);
    // ---- consume in CSR order
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      if (k < deg) {
        float x[V];
        SI::up(v[k], x);
        if (div_src) {
          const float dv = __shfl_sync(hmask, mydiv, k, LPR);
#pragma unroll
          for (int j = 0; j < V; ++j) x[j] /= dv;
        }
        add_vec<V>(acc, x);
      }
    }
    for (int t = NB; t < deg; ++t) {  // long rows: the rest, one at a time (entries beyond LPR straight from memory)
      const int c = (t < LPR) ? __shfl_sync(hmask, cent, t, LPR) : __ldg(ent + cb + t);
      const int r = INC ? (c >> 1) : c;
      const int coff = INC ? (c & 1) * W : 0;
      float x[V];
      SI::up(SI::ld(src + (size_t)r * ROW_LD + coff + lane * V), x);
      if (div_src) {
        const float dv = (float)max(__ldg(ptr + c + 1) - __ldg(ptr + c), 1);
#pragma unroll
        for (int j = 0; j < V; ++j) x[j] /= dv;
      }
      add_vec<V>(acc, x);
    }
    if (!INC && (flags & FVGN_ADJ_DIV_DST_BY_DEG)) {
      const float dd = (float)max(deg, 1);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] /= dd;
    }
    if (!INC && (flags & FVGN_ADJ_ACCUMULATE)) {
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = old[j] + acc[j];
    }
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
  const int rows_per_block = 256 / PipeCfg<W, TS>::LPR;
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
  if (src_type == FVGN_T_F32 && dst_type == FVGN_T_F16) return launch_adj<T_F32, T_F16>(src, ptr, nbr, dst, n_rows, width, flags, stream);
  if (src_type == FVGN_T_F16 && dst_type == FVGN_T_F32) return launch_adj<T_F16, T_F32>(src, ptr, nbr, dst, n_rows, width, flags, stream);
  if (src_type == FVGN_T_F16 && dst_type == FVGN_T_F16) return launch_adj<T_F16, T_F16>(src, ptr, nbr, dst, n_rows, width, flags, stream);
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
  if (src_type == FVGN_T_F32 && dst_type == FVGN_T_F16) return launch_inc<T_F32, T_F16>(src, ptr, code, dst, n_rows, width, stream);
  if (src_type == FVGN_T_F16 && dst_type == FVGN_T_F32) return launch_inc<T_F16, T_F32>(src, ptr, code, dst, n_rows, width, stream);
  if (src_type == FVGN_T_F16 && dst_type == FVGN_T_F16) return launch_inc<T_F16, T_F16>(src, ptr, code, dst, n_rows, width, stream);
#endif
  if (src_type == FVGN_T_F32 && dst_type == FVGN_T_F32) return launch_inc<T_F32, T_F32>(src, ptr, code, dst, n_rows, width, stream);
  return FVGN_ERR_UNSUPPORTED;
}
