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

// ------------------------------------------------------------------------------------------ block-staged kernels
// Round-1 ncu of the warp-pipelined version of these reductions: 132 warp instructions per row, issue slots 48 % busy at
// 25 % of the DRAM bandwidth -- the rows' own data path is ~30 instructions, the rest was index bookkeeping (row pointers
// and entries fetched per warp through a three-deep register pipeline, shuffles, 64-bit predicated address chains).
// Here the INDEX traffic is staged per CTA instead: a CTA owns blocks of RB = 128 consecutive rows; the row pointers of
// block b+2 and the entries of block b+1 arrive by cp.async into shared memory while block b is reduced, so inside a
// block a row costs two LDS for its pointers, one LDS + one LDG.128 per entry and the arithmetic.  A group of LPR lanes
// (16 B of the source row per lane) owns a row; R rows per group are in flight together.  Entries are accumulated
// strictly in CSR order in fp32: bit-identical to the one-shot kernels above.
template <int W, class TS> struct PipeCfg {
  static constexpr int LPR = W * (int)sizeof(typename TS::elem) / 16;   // lanes per row: 8, 16 or 32
  static_assert(LPR == 8 || LPR == 16 || LPR == 32, "row of 128, 256 or 512 bytes");
};
// Measured on a B200 (4 M rows, degree 4, tools/reduce_variants.py, profiles/r2e_reduce_variants.txt): these kernels are
// bound by L2 -> SM gather throughput (every source row is read by each of its ~4 neighbours: 4x the DRAM volume through
// L2) and want many resident warps rather than deep per-thread pipelines -- R = 1 row per group at 5 CTAs per SM
// (48 registers) beats R = 2 at 2-4 CTAs and R = 1 at 8 CTAs (spills).
#ifndef FVGN_PIPE_ROWS
#define FVGN_PIPE_ROWS 1
#endif
#ifndef FVGN_PIPE_MINB
#define FVGN_PIPE_MINB 5   // resident 256-thread blocks per SM the register allocation is held to
#endif
constexpr int RB = 128;    // rows per block
constexpr int ECAP = 1536; // entries of a block held in shared memory (the rest, if any, are read from global memory)

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// INC = false: entries are neighbour rows (src row = entry, W columns);  INC = true: entries are edge*2+role codes
// (src row = code >> 1 of a [E, 2W] array, column block = code & 1).  FL: FVGN_ADJ_* flags, compile time.
template <int W, class TS, class TD, bool INC, int FL, int R>
__global__ void __launch_bounds__(256, FVGN_PIPE_MINB) pipe_reduce_kernel(const typename TS::elem* __restrict__ src,
                                                                          const int32_t* __restrict__ ptr,
                                                                          const int32_t* __restrict__ ent,
                                                                          typename TD::elem* __restrict__ dst, int64_t n) {
  constexpr int LPR = PipeCfg<W, TS>::LPR;  // lanes per row
  constexpr int V = W / LPR;                // elements per lane: 8 (16-bit source) or 4 (fp32 source)
  constexpr int NB = 4;                     // gathers in flight per lane and row
  constexpr int G = 256 / LPR;              // row groups per CTA
  constexpr int ROW_LD = INC ? 2 * W : W;
  constexpr bool DIV_SRC = !INC && (FL & FVGN_ADJ_DIV_SRC_BY_DEG), DIV_DST = !INC && (FL & FVGN_ADJ_DIV_DST_BY_DEG),
                 ACCUM = !INC && (FL & FVGN_ADJ_ACCUMULATE);
  typedef LaneIO<TS, V> SI;
  typedef LaneIO<TD, V> DI;
  __shared__ int s_ptr[3][RB + 1];
  __shared__ int s_ent[2][ECAP];
  const int tid = threadIdx.x, lane = tid & (LPR - 1), g = tid / LPR;
  const int64_t nblk = (n + RB - 1) / RB;
  auto fetch_ptr = [&](int64_t blk, int buf) {      // row pointers of block blk (clamped: rows past n repeat ptr[n])
    if (blk < nblk && tid <= RB) {
      const int64_t row = blk * RB + tid;
      cp_async4(&s_ptr[buf][tid], ptr + (row < n ? row : n));
    }
  };
  auto fetch_ent = [&](int64_t blk, int pbuf, int ebuf) {   // entries of block blk, whose pointers are in s_ptr[pbuf]
    if (blk < nblk) {
      const int e0 = s_ptr[pbuf][0];
      const int ne = min(s_ptr[pbuf][RB] - e0, ECAP);
      for (int i = tid; i < ne; i += 256) cp_async4(&s_ent[ebuf][i], ent + e0 + i);
    }
  };
  int64_t blk = blockIdx.x;
  fetch_ptr(blk, 0);
  fetch_ptr(blk + gridDim.x, 1);
  cp_async_commit_wait_all();
  __syncthreads();
  fetch_ent(blk, 0, 0);
  cp_async_commit_wait_all();
  __syncthreads();
  for (uint32_t it = 0; blk < nblk; blk += gridDim.x, ++it) {
    const int pb = it % 3, eb = it & 1;
    fetch_ptr(blk + 2 * (int64_t)gridDim.x, (it + 2) % 3);
    fetch_ent(blk + gridDim.x, (it + 1) % 3, eb ^ 1);
    const int* sp = s_ptr[pb];
    const int* se = s_ent[eb];
    const int e0 = sp[0];
    const int64_t row0 = blk * RB;
    const int nr = (int)min((int64_t)RB, n - row0);
    auto entry = [&](int t) { return (t - e0 < ECAP) ? se[t - e0] : __ldg(ent + t); };
    auto src_of = [&](int code) {
      const int r = INC ? (code >> 1) : code;
      const int coff = INC ? (code & 1) * W : 0;
      return src + (size_t)r * ROW_LD + coff + lane * V;
    };
#pragma unroll 1
    for (int rr = g; rr < nr; rr += G * R) {
      int b_[R], deg[R], c[R][NB];
      typename SI::raw v[R][NB];
      typename DI::raw oldraw[R];
      float dv[R][NB];
      // ---- issue phase: every gather (and the transposed-mean degrees / the old destination rows) of the R rows
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = rr + r * G;
        b_[r] = 0; deg[r] = 0;
        if (row < nr) { b_[r] = sp[row]; deg[r] = sp[row + 1] - b_[r]; }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          if (k < deg[r]) {
            c[r][k] = entry(b_[r] + k);
            v[r][k] = SI::ld(src_of(c[r][k]));
          }
        }
        if (DIV_SRC) {
#pragma unroll
          for (int k = 0; k < NB; ++k)
            if (k < deg[r]) dv[r][k] = (float)max(__ldg(ptr + c[r][k] + 1) - __ldg(ptr + c[r][k]), 1);
        }
        if (ACCUM && row < nr) oldraw[r] = DI::ld(dst + (size_t)(row0 + row) * W + lane * V);
      }
      // ---- consume in CSR order
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = rr + r * G;
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          if (k < deg[r]) {
            float x[V];
            SI::up(v[r][k], x);
            if (DIV_SRC) {
#pragma unroll
              for (int j = 0; j < V; ++j) x[j] /= dv[r][k];
            }
            add_vec<V>(acc, x);
          }
        }
        for (int t0 = NB; t0 < deg[r]; t0 += NB) {  // rows longer than NB entries: the rest, NB at a time
          typename SI::raw w[NB];
          int cc[NB];
#pragma unroll
          for (int k = 0; k < NB; ++k) {
            cc[k] = 0;
            if (t0 + k < deg[r]) {
              cc[k] = entry(b_[r] + t0 + k);
              w[k] = SI::ld(src_of(cc[k]));
            }
          }
#pragma unroll
          for (int k = 0; k < NB; ++k) {
            if (t0 + k < deg[r]) {
              float x[V];
              SI::up(w[k], x);
              if (DIV_SRC) {
                const float d = (float)max(__ldg(ptr + cc[k] + 1) - __ldg(ptr + cc[k]), 1);
#pragma unroll
                for (int j = 0; j < V; ++j) x[j] /= d;
              }
              add_vec<V>(acc, x);
            }
          }
        }
        if (DIV_DST) {
          const float dd = (float)max(deg[r], 1);
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] /= dd;
        }
        if (row < nr) {
          if (ACCUM) {
            float old[V];
            DI::up(oldraw[r], old);
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = old[j] + acc[j];
          }
          DI::st(dst + (size_t)(row0 + row) * W + lane * V, acc);
        }
      }
    }
    cp_async_commit_wait_all();   // the next block's entries and the pointers of the one after have landed
    __syncthreads();              // ... for every thread; and everybody is done with this block's buffers
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
template <int W, class TS, class TD, bool INC, int FL>
static void launch_pipe_fl(const typename TS::elem* s, const int32_t* ptr, const int32_t* ent, typename TD::elem* o, int64_t n_rows,
                           void* stream) {
  auto kern = pipe_reduce_kernel<W, TS, TD, INC, FL, FVGN_PIPE_ROWS>;
  static unsigned cap_grid[FVGN_MAX_DEV] = {0};  // per instantiation and device
  const int dev = fvgn_cur_device();
  if (cap_grid[dev] == 0) cap_grid[dev] = pipe_grid(kern, (int64_t)1 << 40);
  const int64_t want = (n_rows + RB - 1) / RB;
  const unsigned grid = (unsigned)(want < cap_grid[dev] ? want : cap_grid[dev]);
  kern<<<grid, 256, 0, (cudaStream_t)stream>>>(s, ptr, ent, o, n_rows);
}
template <int W, class TS, class TD, bool INC>
static void launch_pipe(const typename TS::elem* s, const int32_t* ptr, const int32_t* ent, typename TD::elem* o, int64_t n_rows,
                        int flags, void* stream) {
  switch (INC ? 0 : (flags & 7)) {
    case 0: launch_pipe_fl<W, TS, TD, INC, 0>(s, ptr, ent, o, n_rows, stream); break;
    case 1: launch_pipe_fl<W, TS, TD, INC, INC ? 0 : 1>(s, ptr, ent, o, n_rows, stream); break;
    case 2: launch_pipe_fl<W, TS, TD, INC, INC ? 0 : 2>(s, ptr, ent, o, n_rows, stream); break;
    case 4: launch_pipe_fl<W, TS, TD, INC, INC ? 0 : 4>(s, ptr, ent, o, n_rows, stream); break;
    default: launch_pipe_fl<W, TS, TD, INC, INC ? 0 : 7>(s, ptr, ent, o, n_rows, stream); break;   // any other combination
  }
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
