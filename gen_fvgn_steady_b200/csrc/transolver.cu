// Transolver_block kernels (reference: src/FVMmodel/Models/GraphTransolver/GraphTransolver.py:25-169).
//
// The block is  x -> [in_project_fx | in_project_x] -> slice softmax -> per-graph slice tokens -> token attention ->
// de-slice -> to_out (+x) -> LayerNorm -> Linear-GELU-Linear (+residual).  The four dense projections are fvgn_gemm_tf32
// calls (csrc/gemm_tf32.cu) in the tensor-core modes, fp32 library GEMMs in the parity mode; everything the reference does with broadcast products + torch_scatter (its [N,8,32,16]
// temporary, GraphTransolver.py:62-88) and with separate elementwise launches is fused here:
//
//   ts_slice_kernel<true>   in_project_slice + /graph_temperature + softmax (:59-61) and the per-graph token sums
//                           slice_norm / slice_token (:62-72) as deterministic per-chunk partials
//   ts_slice_kernel<false>  the same token sum for an arbitrary value tile (backward: d out_slice_token)
//   ts_token_attention_*    attention among the 32 slice tokens of a (graph, head) (:72-81) and its autograd
//   ts_deslice_kernel       out_x[n,h,:] = sum_g sw[n,h,g] tok[b(n),h,g,:]   (:83-90)
//   ts_slice_bwd_kernel     autograd of slice + de-slice w.r.t. the projections, in_project_slice and graph_temperature
//   ts_res_ln_*             y = a + bias + residual ; z = LayerNorm(y)       (:163-169, to_out bias + ln_2) and backward
//   ts_bias_gelu_*          h = GELU(hpre + b) (:105,124) and backward ; ts_bias_res: out = a + bias + residual
//
// Sizes are the reference's: 8 heads x 16 channels, 32 slices (TransFVGN_v1.py / _v2.py: num_heads=8, slice_num=32).
// Thread mappings avoid warp shuffles: per-(node, head) work is done by one thread on a padded shared-memory tile
// (row stride 260 / 132 floats: lane n touches banks 4n..4n+3, conflict-free 16-B accesses), cross-node sums walk the
// tile in row order, so every result is bit-reproducible for a given chunk table.
#include "common.cuh"

#ifndef FVGN_EMU
#include <cuda_bf16.h>
#endif

namespace {

constexpr int TS_HEADS = 8;
constexpr int TS_DH = 16;
constexpr int TS_G = 32;
constexpr int TS_TILE = 32;                                // nodes per tile (= lanes of the per-head warp)
constexpr int TS_RS = 260;                                 // padded row stride of a [32][256] tile
constexpr int TS_RS1 = 132;                                // padded row stride of a [32][128] tile
constexpr int TS_TOK = TS_HEADS * TS_G * TS_DH;            // 4096 token numerators per graph
constexpr int TS_TOKW = TS_TOK + TS_HEADS * TS_G;          // + 256 norms
constexpr int TS_PARAMW = TS_G * TS_DH + TS_G + TS_HEADS + 256;  // dWs | dbs | dT | colsum(dP)
constexpr int TS_NT = 256;

// Tile prefetch: 16-byte cp.async copies into the padded shared-memory tile (rows past the chunk end are zero-filled), so
// the next tile's global loads are in flight while the current tile is computed.  Under the CPU emulator (tests only)
// the same call is a plain copy and commit / wait are no-ops: the double-buffer logic is identical.
#if defined(FVGN_EMU) || defined(TS_NO_ASYNC)
#define TS_ASYNC 0
#else
#define TS_ASYNC 1
#include <cuda_pipeline.h>
#endif

template <int W, int RS>
__device__ __forceinline__ void ts_prefetch_tile(float* s, const float* __restrict__ g, int64_t ld, int64_t row0, int64_t r1) {
  constexpr int V = W / 4;
  for (int i = threadIdx.x; i < TS_TILE * V; i += TS_NT) {
    const int r = i / V, c4 = i % V;
    float* dst = s + r * RS + c4 * 4;
    if (row0 + r < r1) {
      const float* src = g + (size_t)(row0 + r) * ld + c4 * 4;
#if TS_ASYNC
      __pipeline_memcpy_async(dst, src, 16);
#else
      st4(dst, ld4(src));
#endif
    } else {
      st4(dst, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}
__device__ __forceinline__ void ts_prefetch_commit() {
#if TS_ASYNC
  __pipeline_commit();
#endif
}
__device__ __forceinline__ void ts_prefetch_wait() {
#if TS_ASYNC
  __pipeline_wait_prior(0);
#endif
}

template <int W, int RS>
__device__ __forceinline__ void ts_store_tile(const float* s, float* __restrict__ g, int64_t ld, int64_t row0, int64_t r1) {
  constexpr int V = W / 4;
  for (int i = threadIdx.x; i < TS_TILE * V; i += TS_NT) {
    const int r = i / V, c4 = i % V;
    if (row0 + r < r1) st4(g + (size_t)(row0 + r) * ld + c4 * 4, ld4(s + r * RS + c4 * 4));
  }
}

// 16-wide fp32 vectors are kept as 8 float2 so the inner products / axpys issue as packed FFMA2 (sm_100: the 3-register
// scalar FFMA runs at half rate; the packed form restores the full fp32 rate and halves the issue slots).
#ifdef FVGN_EMU
static inline float2 fma2(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
#else
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#endif

__device__ __forceinline__ void ld16(float2* dst, const float* src) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = ld4(src + q * 4);
    dst[q * 2 + 0] = make_float2(v.x, v.y);
    dst[q * 2 + 1] = make_float2(v.z, v.w);
  }
}
__device__ __forceinline__ void st16(float* dst, const float2* src) {
#pragma unroll
  for (int q = 0; q < 4; ++q) st4(dst + q * 4, make_float4(src[q * 2].x, src[q * 2].y, src[q * 2 + 1].x, src[q * 2 + 1].y));
}
__device__ __forceinline__ void zero16(float2* a) {
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = make_float2(0.f, 0.f);
}
// acc += s * x
__device__ __forceinline__ void axpy16(float2* acc, float s, const float2* x) {
  const float2 s2 = make_float2(s, s);
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = fma2(s2, x[k], acc[k]);
}
// c + x . w as four independent chains
__device__ __forceinline__ float dot16(const float2* x, const float2* w, float c) {
  float2 p = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);  // (seeding p.x with c makes ptxas spill: keep c outside)
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    p = fma2(x[k], w[k], p);
    q = fma2(x[k + 1], w[k + 1], q);
  }
  return c + ((p.x + p.y) + (q.x + q.y));
}

// ----------------------------------------------------------------------------------------------------------------
// FROM_P : P[N,256] = [fx_mid | x_mid] -> sw[N,256] (written) and partial[chunk] = (sum_n sw (x) fx_mid | sum_n sw)
// !FROM_P: sw[N,256] (read), V[N,128]  ->                  partial[chunk] = (sum_n sw (x) V      | sum_n sw)
template <bool FROM_P>
struct SliceSmem {
  static constexpr int VRS = FROM_P ? TS_RS : TS_RS1;         // row stride of the value tile
  static constexpr int VT = TS_TILE * VRS;                     // floats of one value tile
  static constexpr int ST = TS_TILE * TS_RS;                   // floats of one sw tile
  static constexpr int NSW = FROM_P ? 1 : 2;                   // sw tile: produced in place (1) or prefetched (2 buffers)
  static constexpr size_t bytes = sizeof(float) * (2 * VT + NSW * ST + TS_G * TS_DH + TS_G);
};

template <bool FROM_P>
__global__ void __launch_bounds__(TS_NT) ts_slice_kernel(const float* __restrict__ P, float* sw, const float* __restrict__ Ws,
                                                         const float* __restrict__ bs, const float* __restrict__ temp,
                                                         const int32_t* __restrict__ chunks, float* __restrict__ partial) {
  using L = SliceSmem<FROM_P>;
  FVGN_DYN_SMEM(smem);
  float* Vb = reinterpret_cast<float*>(smem);  // 2 x [32][VRS]  (FROM_P: fx|xm ; else the value tile V)
  float* Sb = Vb + 2 * L::VT;                  // NSW x [32][260]
  float* Wss = Sb + L::NSW * L::ST;            // [32][16]
  float* bss = Wss + TS_G * TS_DH;             // [32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t r0 = chunks[blockIdx.x * 3 + 1], r1 = chunks[blockIdx.x * 3 + 2];
  float T = 1.f;
  if (FROM_P) {
    for (int i = tid; i < TS_G * TS_DH; i += TS_NT) Wss[i] = Ws[i];
    if (tid < TS_G) bss[tid] = bs[tid];
    T = temp[warp];
  }
  float2 acc[8];
  zero16(acc);
  float nrm = 0.f;
  if (r0 < r1) {
    if (FROM_P) {
      ts_prefetch_tile<256, TS_RS>(Vb, P, 256, r0, r1);
    } else {
      ts_prefetch_tile<128, TS_RS1>(Vb, P, 128, r0, r1);
      ts_prefetch_tile<256, TS_RS>(Sb, sw, 256, r0, r1);
    }
  }
  ts_prefetch_commit();
  int buf = 0;
  for (int64_t row0 = r0; row0 < r1; row0 += TS_TILE, buf ^= 1) {
    ts_prefetch_wait();
    __syncthreads();  // tile `buf` has landed; every thread is done with the previous tile (its buffers may be refilled)
    if (row0 + TS_TILE < r1) {
      if (FROM_P) {
        ts_prefetch_tile<256, TS_RS>(Vb + (buf ^ 1) * L::VT, P, 256, row0 + TS_TILE, r1);
      } else {
        ts_prefetch_tile<128, TS_RS1>(Vb + (buf ^ 1) * L::VT, P, 128, row0 + TS_TILE, r1);
        ts_prefetch_tile<256, TS_RS>(Sb + (buf ^ 1) * L::ST, sw, 256, row0 + TS_TILE, r1);
      }
    }
    ts_prefetch_commit();
    const float* Vs = Vb + buf * L::VT;
    float* sws = FROM_P ? Sb : Sb + buf * L::ST;
    if (FROM_P) {
      // thread = (head = warp, node = lane)
      float2 xm[8];
      float l[TS_G];
      ld16(xm, Vs + lane * TS_RS + 128 + warp * 16);
#pragma unroll
      for (int g = 0; g < TS_G; ++g) {  // in_project_slice (GraphTransolver.py:60)
        float2 w[8];
        ld16(w, Wss + g * 16);
        l[g] = dot16(xm, w, bss[g]);
      }
      float m = -INFINITY;
#pragma unroll
      for (int g = 0; g < TS_G; ++g) { l[g] = l[g] / T; m = fmaxf(m, l[g]); }
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int g = 0; g < TS_G; g += 2) {
        l[g] = expf(l[g] - m);
        l[g + 1] = expf(l[g + 1] - m);
        s0 += l[g];
        s1 += l[g + 1];
      }
      const float s = s0 + s1;
      const bool valid = row0 + lane < r1;
#pragma unroll
      for (int g = 0; g < TS_G; ++g) l[g] = valid ? l[g] / s : 0.f;
      float* dst = sws + lane * TS_RS + warp * TS_G;
#pragma unroll
      for (int g4 = 0; g4 < TS_G / 4; ++g4) st4(dst + g4 * 4, make_float4(l[g4 * 4], l[g4 * 4 + 1], l[g4 * 4 + 2], l[g4 * 4 + 3]));
      __syncthreads();
      ts_store_tile<256, TS_RS>(sws, sw, 256, row0, r1);
    }
    // thread = (head = warp, slice = lane): token sums in row order
#pragma unroll 4
    for (int n = 0; n < TS_TILE; ++n) {
      const float sv = sws[n * TS_RS + warp * TS_G + lane];
      float2 f[8];
      ld16(f, Vs + n * L::VRS + warp * 16);
      axpy16(acc, sv, f);
      nrm += sv;
    }
  }
  float* out = partial + (size_t)blockIdx.x * TS_TOKW;
  st16(out + (warp * TS_G + lane) * 16, acc);
  out[TS_TOK + warp * TS_G + lane] = nrm;
}

// ----------------------------------------------------------------------------------------------------------------
// out[n, h*16+d] = sum_g sw[n,h,g] tok[seg % tok_mod, h, g, d]
__global__ void __launch_bounds__(TS_NT) ts_deslice_kernel(const float* __restrict__ sw, const float* __restrict__ tok,
                                                           int64_t tok_ld, int tok_mod, const int32_t* __restrict__ chunks,
                                                           float* __restrict__ out) {
  FVGN_DYN_SMEM(smem);
  float* Sb = reinterpret_cast<float*>(smem);   // 2 x [32][260]
  float* Ts = Sb + 2 * TS_TILE * TS_RS;         // [8][32][16]
  float* outs = Ts + TS_TOK;                    // [32][132]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int seg = chunks[blockIdx.x * 3 + 0];
  const int64_t r0 = chunks[blockIdx.x * 3 + 1], r1 = chunks[blockIdx.x * 3 + 2];
  const float* tp = tok + (size_t)(seg % tok_mod) * tok_ld;
  for (int i = tid; i < TS_TOK / 4; i += TS_NT) st4(Ts + i * 4, ld4(tp + i * 4));
  if (r0 < r1) ts_prefetch_tile<256, TS_RS>(Sb, sw, 256, r0, r1);
  ts_prefetch_commit();
  int buf = 0;
  for (int64_t row0 = r0; row0 < r1; row0 += TS_TILE, buf ^= 1) {
    ts_prefetch_wait();
    __syncthreads();
    if (row0 + TS_TILE < r1) ts_prefetch_tile<256, TS_RS>(Sb + (buf ^ 1) * TS_TILE * TS_RS, sw, 256, row0 + TS_TILE, r1);
    ts_prefetch_commit();
    const float* srow = Sb + buf * TS_TILE * TS_RS + lane * TS_RS + warp * TS_G;
    float2 o[8];
    zero16(o);
#pragma unroll
    for (int g4 = 0; g4 < TS_G / 4; ++g4) {
      const float4 s4 = ld4(srow + g4 * 4);
      const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 t[8];
        ld16(t, Ts + (warp * TS_G + g4 * 4 + j) * 16);
        axpy16(o, sv[j], t);
      }
    }
    st16(outs + lane * TS_RS1 + warp * 16, o);
    __syncthreads();
    ts_store_tile<128, TS_RS1>(outs, out, 128, row0, r1);
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Backward of slice + de-slice for one chunk.  Inputs: P = [fx|xm], sw, dox = d out_x, tok_out (forward tokens after
// attention), d_tok = [d slice_token numerators | d slice_norm] (zero for chunks of graphs this rank does not own).
// Outputs: dP[N,256] = [d fx | d xm] and partial[chunk] = (dWs[32,16] | dbs[32] | d graph_temperature[8] | colsum dP[256]).
constexpr int TS_BSET = 2 * TS_TILE * TS_RS + TS_TILE * TS_RS1;  // floats of one input tile set: P | sw | d out_x

__global__ void __launch_bounds__(TS_NT, 1) ts_slice_bwd_kernel(const float* __restrict__ P, const float* __restrict__ sw,
                                                                const float* __restrict__ dox, const float* __restrict__ tok_out,
                                                                const float* __restrict__ d_tok, int tok_mod,
                                                                const float* __restrict__ Ws, const float* __restrict__ bs,
                                                                const float* __restrict__ temp, const int32_t* __restrict__ chunks,
                                                                float* __restrict__ dP, float* __restrict__ partial) {
  FVGN_DYN_SMEM(smem);
  float* Bb = reinterpret_cast<float*>(smem);  // 2 x { Ps [32][260] | sws [32][260] | dxs [32][132] }
  float* OTs = Bb + 2 * TS_BSET;               // [8][32][16]
  float* DNs = OTs + TS_TOK;                   // [8][32][16]
  float* dns = DNs + TS_TOK;                   // [8][32]
  float* Wss = dns + TS_HEADS * TS_G;          // [32][16]
  float* bss = Wss + TS_G * TS_DH;             // [32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int seg = chunks[blockIdx.x * 3 + 0];
  const int64_t r0 = chunks[blockIdx.x * 3 + 1], r1 = chunks[blockIdx.x * 3 + 2];
  const bool own = seg < tok_mod;
  const float* otp = tok_out + (size_t)(seg % tok_mod) * TS_TOK;
  const float* dtp = d_tok + (size_t)(seg % tok_mod) * TS_TOKW;
  for (int i = tid; i < TS_TOK / 4; i += TS_NT) {
    st4(OTs + i * 4, ld4(otp + i * 4));
    st4(DNs + i * 4, own ? ld4(dtp + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f));
  }
  dns[tid] = own ? dtp[TS_TOK + tid] : 0.f;
  for (int i = tid; i < TS_G * TS_DH; i += TS_NT) Wss[i] = Ws[i];
  if (tid < TS_G) bss[tid] = bs[tid];
  const float T = temp[warp];
  float2 accW[8];
  zero16(accW);
  float accb = 0.f, accT = 0.f, acccol = 0.f;
  if (r0 < r1) {
    ts_prefetch_tile<256, TS_RS>(Bb, P, 256, r0, r1);
    ts_prefetch_tile<256, TS_RS>(Bb + TS_TILE * TS_RS, sw, 256, r0, r1);
    ts_prefetch_tile<128, TS_RS1>(Bb + 2 * TS_TILE * TS_RS, dox, 128, r0, r1);
  }
  ts_prefetch_commit();
  int buf = 0;
  for (int64_t row0 = r0; row0 < r1; row0 += TS_TILE, buf ^= 1) {
    ts_prefetch_wait();
    __syncthreads();  // tile set `buf` has landed; every thread is done with the previous tile set
    if (row0 + TS_TILE < r1) {
      float* nb = Bb + (buf ^ 1) * TS_BSET;
      ts_prefetch_tile<256, TS_RS>(nb, P, 256, row0 + TS_TILE, r1);
      ts_prefetch_tile<256, TS_RS>(nb + TS_TILE * TS_RS, sw, 256, row0 + TS_TILE, r1);
      ts_prefetch_tile<128, TS_RS1>(nb + 2 * TS_TILE * TS_RS, dox, 128, row0 + TS_TILE, r1);
    }
    ts_prefetch_commit();
    float* Ps = Bb + buf * TS_BSET;
    float* sws = Ps + TS_TILE * TS_RS;
    float* dxs = sws + TS_TILE * TS_RS;
    {  // thread = (head = warp, node = lane); sw stays in the tile (16-B reads), d sw / d logits live in registers
      float ds[TS_G];
      float2 a[8], b[8], f[8];
      float* srow = sws + lane * TS_RS + warp * TS_G;
      float* frow = Ps + lane * TS_RS + warp * 16;
      float* xrow = dxs + lane * TS_RS1 + warp * 16;
      ld16(a, xrow);  // d out_x
      ld16(b, frow);  // fx_mid
      zero16(f);
      float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
      for (int g4 = 0; g4 < TS_G / 4; ++g4) {
        const float4 s4 = ld4(srow + g4 * 4);
        const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int g = g4 * 4 + j;
          float2 o[8], dn[8];
          ld16(o, OTs + (warp * TS_G + g) * 16);
          ld16(dn, DNs + (warp * TS_G + g) * 16);
          // d sw[g] = d_norm[g] + d out_x . tok_out[g] + fx . d_num[g]
          const float v = dot16(a, o, dns[warp * TS_G + g]) + dot16(b, dn, 0.f);
          axpy16(f, sv[j], dn);  // d fx_mid[d] = sum_g sw[g] d_num[g,d]
          ds[g] = v;
          if (j & 1) dot1 += sv[j] * v; else dot0 += sv[j] * v;
        }
      }
      const float dot = dot0 + dot1;
      st16(frow, f);  // overwrites fx in the tile
      // softmax backward (scaled logits z = l / T), then through /T and in_project_slice
      ld16(b, Ps + lane * TS_RS + 128 + warp * 16);  // x_mid
      zero16(a);
      float t0 = 0.f, t1 = 0.f;
#pragma unroll
      for (int g4 = 0; g4 < TS_G / 4; ++g4) {
        const float4 s4 = ld4(srow + g4 * 4);
        const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
        float dl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int g = g4 * 4 + j;
          float2 w[8];
          ld16(w, Wss + g * 16);
          const float lg = dot16(b, w, bss[g]);
          const float dz = sv[j] * (ds[g] - dot);
          if (j & 1) t1 += dz * lg; else t0 += dz * lg;
          dl[j] = dz / T;
          axpy16(a, dl[j], w);
        }
        st4(srow + g4 * 4, make_float4(dl[0], dl[1], dl[2], dl[3]));  // d logits (pre-temperature) for the dWs sum below
      }
      accT += t0 + t1;
      st16(xrow, a);  // d x_mid
    }
    __syncthreads();
    // thread = (head = warp, slice = lane): dWs[g,c] += dl[n,h,g] x_mid[n,h,c], dbs[g] += dl
#pragma unroll 4
    for (int n = 0; n < TS_TILE; ++n) {
      const float dl = sws[n * TS_RS + warp * TS_G + lane];
      float2 x[8];
      ld16(x, Ps + n * TS_RS + 128 + warp * 16);
      axpy16(accW, dl, x);
      accb += dl;
    }
    // dP tile = [d fx (Ps cols 0..127) | d xm (dxs)], column sums for the projection biases
    for (int i = tid; i < TS_TILE * 64; i += TS_NT) {
      const int r = i >> 6, c4 = i & 63;
      if (row0 + r < r1) {
        const float4 v = c4 < 32 ? ld4(Ps + r * TS_RS + c4 * 4) : ld4(dxs + r * TS_RS1 + (c4 - 32) * 4);
        st4(dP + (size_t)(row0 + r) * 256 + c4 * 4, v);
      }
    }
    {
      float c0 = 0.f, c1 = 0.f;
#pragma unroll 4
      for (int r = 0; r < TS_TILE; r += 2) {
        c0 += tid < 128 ? Ps[r * TS_RS + tid] : dxs[r * TS_RS1 + tid - 128];
        c1 += tid < 128 ? Ps[(r + 1) * TS_RS + tid] : dxs[(r + 1) * TS_RS1 + tid - 128];
      }
      acccol += c0 + c1;
    }
  }
  __syncthreads();
  // cross-warp (head) sums in fixed order
  float* red = Bb + TS_TILE * TS_RS;  // [8][32][16]   (sw tile of set 0)
  float* red2 = Bb;                   // [8][32] dbs partials, then [8][32] dT partials
  st16(red + (warp * TS_G + lane) * 16, accW);
  red2[warp * 32 + lane] = accb;
  red2[256 + warp * 32 + lane] = accT;
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * TS_PARAMW;
  for (int i = tid; i < TS_G * TS_DH; i += TS_NT) {
    float v = 0.f;
#pragma unroll
    for (int h = 0; h < TS_HEADS; ++h) v += red[h * TS_G * 16 + i];
    out[i] = v;
  }
  if (tid < TS_G) {
    float v = 0.f;
#pragma unroll
    for (int h = 0; h < TS_HEADS; ++h) v += red2[h * 32 + tid];
    out[TS_G * TS_DH + tid] = v;
  }
  if (tid < TS_HEADS) {
    float v = 0.f;
    for (int l = 0; l < 32; ++l) v += red2[256 + tid * 32 + l];
    const float Th = temp[tid];
    out[TS_G * TS_DH + TS_G + tid] = -v / (Th * Th);
  }
  out[TS_G * TS_DH + TS_G + TS_HEADS + tid] = acccol;
}

// ----------------------------------------------------------------------------------------------------------------
// Row-wise kernels on [n,128] rows: warp = row, lane = 4 columns; row sums go through a per-warp shared-memory slot.
#ifdef FVGN_EMU
#define TS_WARP_SYNC() emu::syncwarp()
#else
#define TS_WARP_SYNC() __syncwarp()
#endif

__device__ __forceinline__ float ts_sum32(const float* p) {
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 v = ld4(p + q * 4);
    s += v.x; s += v.y; s += v.z; s += v.w;
  }
  return s;
}

// y = a + bias + res ; z = LayerNorm(y) * gamma + beta ; stats[row] = (mean, rstd)       (eps = 1e-5, nn.LayerNorm default)
__global__ void __launch_bounds__(TS_NT) ts_res_ln_fwd_kernel(const float* __restrict__ a, const float* __restrict__ bias,
                                                              const float* __restrict__ res, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float* __restrict__ y,
                                                              float* __restrict__ z, float* __restrict__ stats, int64_t n) {
  __shared__ __align__(16) float red[8][4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 b4 = ld4(bias + lane * 4), g4 = ld4(gamma + lane * 4), be4 = ld4(beta + lane * 4);
  int it = 0;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n; row += (int64_t)gridDim.x * 8, ++it) {
    float4 v = add4(add4(ld4(a + row * 128 + lane * 4), b4), ld4(res + row * 128 + lane * 4));
    st4(y + row * 128 + lane * 4, v);
    float* r1 = red[warp][(it & 1) * 2];
    float* r2 = red[warp][(it & 1) * 2 + 1];
    r1[lane] = (v.x + v.y) + (v.z + v.w);
    TS_WARP_SYNC();
    const float mean = ts_sum32(r1) * (1.0f / 128.0f);
    v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
    r2[lane] = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    TS_WARP_SYNC();
    const float rstd = 1.0f / sqrtf(ts_sum32(r2) * (1.0f / 128.0f) + 1e-5f);
    st4(z + row * 128 + lane * 4, make_float4(v.x * rstd * g4.x + be4.x, v.y * rstd * g4.y + be4.y,
                                              v.z * rstd * g4.z + be4.z, v.w * rstd * g4.w + be4.w));
    if (lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
  }
}

// d_y = LayerNorm-backward(dz) + d_y_in ; partial[cta] = (dgamma[128] | dbeta[128] | colsum d_y[128] | colsum d_y_in[128])
__global__ void __launch_bounds__(TS_NT) ts_res_ln_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                                              const float* __restrict__ stats, const float* __restrict__ gamma,
                                                              const float* __restrict__ d_y_in, float* __restrict__ d_y,
                                                              float* __restrict__ partial, int64_t n) {
  __shared__ __align__(16) float red[8][4][32];
  __shared__ float redc[8][512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 g4 = ld4(gamma + lane * 4);
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag, ac = ag, ai = ag;
  int it = 0;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n; row += (int64_t)gridDim.x * 8, ++it) {
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    const float4 yv = ld4(y + row * 128 + lane * 4), d = ld4(dz + row * 128 + lane * 4);
    const float4 yh = make_float4((yv.x - mean) * rstd, (yv.y - mean) * rstd, (yv.z - mean) * rstd, (yv.w - mean) * rstd);
    const float4 g = make_float4(d.x * g4.x, d.y * g4.y, d.z * g4.z, d.w * g4.w);
    float* r1 = red[warp][(it & 1) * 2];
    float* r2 = red[warp][(it & 1) * 2 + 1];
    r1[lane] = (g.x + g.y) + (g.z + g.w);
    r2[lane] = (g.x * yh.x + g.y * yh.y) + (g.z * yh.z + g.w * yh.w);
    TS_WARP_SYNC();
    const float c1 = ts_sum32(r1) * (1.0f / 128.0f), c2 = ts_sum32(r2) * (1.0f / 128.0f);
    float4 o = make_float4(rstd * (g.x - c1 - yh.x * c2), rstd * (g.y - c1 - yh.y * c2), rstd * (g.z - c1 - yh.z * c2),
                           rstd * (g.w - c1 - yh.w * c2));
    if (d_y_in) {
      const float4 e = ld4(d_y_in + row * 128 + lane * 4);
      o = add4(o, e);
      ai = add4(ai, e);
    }
    st4(d_y + row * 128 + lane * 4, o);
    ag = add4(ag, make_float4(d.x * yh.x, d.y * yh.y, d.z * yh.z, d.w * yh.w));
    ab = add4(ab, d);
    ac = add4(ac, o);
  }
  st4(&redc[warp][lane * 4], ag);
  st4(&redc[warp][128 + lane * 4], ab);
  st4(&redc[warp][256 + lane * 4], ac);
  st4(&redc[warp][384 + lane * 4], ai);
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += TS_NT) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += redc[w][i];
    partial[(size_t)blockIdx.x * 512 + i] = s;
  }
}

// h = GELU(hpre + bias) on [n,256] rows (exact erf GELU, nn.GELU default)
__global__ void __launch_bounds__(TS_NT) ts_bias_gelu_fwd_kernel(const float* __restrict__ hpre, const float* __restrict__ bias,
                                                                 float* __restrict__ h, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * TS_NT + threadIdx.x;
  if (i >= n4) return;
  const float4 b = ld4(bias + (i & 63) * 4), v = ld4(hpre + i * 4);
  st4(h + i * 4, make_float4(gelu_exact(v.x + b.x), gelu_exact(v.y + b.y), gelu_exact(v.z + b.z), gelu_exact(v.w + b.w)));
}

// dhpre = dh * GELU'(hpre + bias) (dhpre may alias dh) ; partial[cta] = colsum dhpre [256]
__global__ void __launch_bounds__(TS_NT) ts_bias_gelu_bwd_kernel(const float* dh, const float* __restrict__ hpre,
                                                                 const float* __restrict__ bias, float* dhpre,
                                                                 float* __restrict__ partial, int64_t n) {
  __shared__ __align__(16) float red[4][256];
  const int c4 = threadIdx.x & 63, rsub = threadIdx.x >> 6;
  const float4 b = ld4(bias + c4 * 4);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = (int64_t)blockIdx.x * 4 + rsub; row < n; row += (int64_t)gridDim.x * 4) {
    const float4 v = ld4(hpre + row * 256 + c4 * 4), d = ld4(dh + row * 256 + c4 * 4);
    const float4 o = make_float4(d.x * gelu_grad(v.x + b.x), d.y * gelu_grad(v.y + b.y), d.z * gelu_grad(v.z + b.z),
                                 d.w * gelu_grad(v.w + b.w));
    st4(dhpre + row * 256 + c4 * 4, o);
    acc = add4(acc, o);
  }
  st4(&red[rsub][c4 * 4], acc);
  __syncthreads();
  partial[(size_t)blockIdx.x * 256 + threadIdx.x] =
      ((red[0][threadIdx.x] + red[1][threadIdx.x]) + red[2][threadIdx.x]) + red[3][threadIdx.x];
}

// out = a + bias + res on [n,128] rows (+ bf16 shadow of out)
__global__ void __launch_bounds__(TS_NT) ts_bias_res_kernel(const float* __restrict__ a, const float* __restrict__ bias,
                                                            const float* __restrict__ res, float* __restrict__ out,
                                                            uint16_t* __restrict__ outh, int outh_f16, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * TS_NT + threadIdx.x;
  if (i >= n4) return;
  const float4 v = add4(add4(ld4(a + i * 4), ld4(bias + (i & 31) * 4)), ld4(res + i * 4));
  st4(out + i * 4, v);
#ifndef FVGN_EMU
  if (outh) {
    uint32_t lo, hi;
    if (outh_f16) {
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v.y), "f"(v.x));
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.w), "f"(v.z));
    } else {
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v.y), "f"(v.x));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.w), "f"(v.z));
    }
    *reinterpret_cast<uint2*>(outh + i * 4) = make_uint2(lo, hi);
  }
#endif
}

constexpr size_t kSmemDeslice = sizeof(float) * (2 * TS_TILE * TS_RS + TS_TOK + TS_TILE * TS_RS1);
constexpr size_t kSmemBwd = sizeof(float) * (2 * TS_BSET + 2 * TS_TOK + TS_HEADS * TS_G + TS_G * TS_DH + TS_G);

template <class K>
int ts_set_smem(K kern, size_t bytes) {
#ifndef FVGN_EMU
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return FVGN_ERR_LAUNCH;
#endif
  return FVGN_OK;
}

constexpr int kRowCtas = 148 * 4;  // CTAs of the grid-stride row kernels (also their number of partial rows)

// ----------------------------------------------------------------------------------------------------------------
// Attention among the slice tokens of one (graph, head) (GraphTransolver.py:72-81): tok = num / (norm + 1e-5),
// q, k, v = tok Wq^T, tok Wk^T, tok Wv^T (16 x 16, shared by the heads), attn = softmax(q k^T * scale), out = attn v.
// One 32-thread CTA per (graph, head), thread = token g: its row of the 32 x 32 score matrix lives in registers; k, v (and in
// the backward q, d out, the score gradients) are exchanged through padded shared-memory tiles.  No shuffles, fixed
// summation order.  The backward recomputes the forward from the token record.
constexpr int TA_RS = 17;   // padded row stride of a [32][16] tile
constexpr int TA_RS2 = 33;  // padded row stride of a [32][32] tile

struct TokAttnFwd {   // what forward and backward both need, per thread (token g)
  float tok[TS_DH], q[TS_DH], k[TS_DH], v[TS_DH], attn[TS_G], inv;
};

// rec_b: the token record of this graph [4352]; W*: [16][16] in shared memory; ks / vs: [32][TA_RS] tiles (written here)
__device__ __forceinline__ void tok_attn_forward(const float* __restrict__ rec_b, int h, int g, const float* Wq, const float* Wk,
                                                 const float* Wv, float scale, float* ks, float* vs, TokAttnFwd& f) {
  const float* num = rec_b + ((size_t)h * TS_G + g) * TS_DH;
  f.inv = 1.0f / (rec_b[TS_TOK + h * TS_G + g] + 1e-5f);
  for (int i = 0; i < TS_DH; ++i) f.tok[i] = num[i] * f.inv;
  for (int o = 0; o < TS_DH; ++o) {
    float aq = 0.f, ak = 0.f, av = 0.f;
    for (int i = 0; i < TS_DH; ++i) {
      aq = fmaf(f.tok[i], Wq[o * TS_DH + i], aq);
      ak = fmaf(f.tok[i], Wk[o * TS_DH + i], ak);
      av = fmaf(f.tok[i], Wv[o * TS_DH + i], av);
    }
    f.q[o] = aq; f.k[o] = ak; f.v[o] = av;
    ks[g * TA_RS + o] = ak;
    vs[g * TA_RS + o] = av;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int j = 0; j < TS_G; ++j) {
    float d = 0.f;
    for (int i = 0; i < TS_DH; ++i) d = fmaf(f.q[i], ks[j * TA_RS + i], d);
    f.attn[j] = d * scale;
    m = fmaxf(m, f.attn[j]);
  }
  float sum = 0.f;
  for (int j = 0; j < TS_G; ++j) {
    f.attn[j] = expf(f.attn[j] - m);
    sum += f.attn[j];
  }
  const float rs = 1.0f / sum;
  for (int j = 0; j < TS_G; ++j) f.attn[j] *= rs;
}

__global__ void __launch_bounds__(TS_G) ts_token_attention_fwd_kernel(const float* __restrict__ rec, const float* __restrict__ wq,
                                                                      const float* __restrict__ wk, const float* __restrict__ wv,
                                                                      float scale, float* __restrict__ out) {
  __shared__ float Wq[TS_DH * TS_DH], Wk[TS_DH * TS_DH], Wv[TS_DH * TS_DH], ks[TS_G * TA_RS], vs[TS_G * TA_RS];
  const int g = threadIdx.x, b = blockIdx.x / TS_HEADS, h = blockIdx.x % TS_HEADS;
  for (int i = g; i < TS_DH * TS_DH; i += TS_G) { Wq[i] = wq[i]; Wk[i] = wk[i]; Wv[i] = wv[i]; }
  __syncthreads();
  TokAttnFwd f;
  tok_attn_forward(rec + (size_t)b * TS_TOKW, h, g, Wq, Wk, Wv, scale, ks, vs, f);
  float o[TS_DH];
  for (int i = 0; i < TS_DH; ++i) o[i] = 0.f;
  for (int j = 0; j < TS_G; ++j)
    for (int i = 0; i < TS_DH; ++i) o[i] = fmaf(f.attn[j], vs[j * TA_RS + i], o[i]);
  float* dst = out + (size_t)b * TS_TOK + ((size_t)h * TS_G + g) * TS_DH;
  for (int i = 0; i < TS_DH; ++i) dst[i] = o[i];
}

// d_rec[b] = gradient of the token record; wpart[(b, h)] = this CTA's (dWq | dWk | dWv) [3][16][16]
__global__ void __launch_bounds__(TS_G) ts_token_attention_bwd_kernel(const float* __restrict__ rec, const float* __restrict__ wq,
                                                                      const float* __restrict__ wk, const float* __restrict__ wv,
                                                                      float scale, const float* __restrict__ d_out,
                                                                      float* __restrict__ d_rec, float* __restrict__ wpart) {
  __shared__ float Wq[TS_DH * TS_DH], Wk[TS_DH * TS_DH], Wv[TS_DH * TS_DH];
  __shared__ float ks[TS_G * TA_RS], vs[TS_G * TA_RS], qs[TS_G * TA_RS], dos[TS_G * TA_RS], toks[TS_G * TA_RS];
  __shared__ float dqs[TS_G * TA_RS], dks[TS_G * TA_RS], dvs[TS_G * TA_RS];
  __shared__ float as_[TS_G * TA_RS2], dss[TS_G * TA_RS2];
  const int g = threadIdx.x, b = blockIdx.x / TS_HEADS, h = blockIdx.x % TS_HEADS;
  for (int i = g; i < TS_DH * TS_DH; i += TS_G) { Wq[i] = wq[i]; Wk[i] = wk[i]; Wv[i] = wv[i]; }
  __syncthreads();
  TokAttnFwd f;
  tok_attn_forward(rec + (size_t)b * TS_TOKW, h, g, Wq, Wk, Wv, scale, ks, vs, f);
  const float* dsrc = d_out + (size_t)b * TS_TOK + ((size_t)h * TS_G + g) * TS_DH;
  float dO[TS_DH];
  for (int i = 0; i < TS_DH; ++i) {
    dO[i] = dsrc[i];
    dos[g * TA_RS + i] = dO[i];
    qs[g * TA_RS + i] = f.q[i];
    toks[g * TA_RS + i] = f.tok[i];
  }
  // d attn[g][j] = dO[g] . v[j]; softmax backward; d dots = scale * attn * (d attn - sum_j attn d attn)
  float da[TS_G], dot = 0.f;
  for (int j = 0; j < TS_G; ++j) {
    float d = 0.f;
    for (int i = 0; i < TS_DH; ++i) d = fmaf(dO[i], vs[j * TA_RS + i], d);
    da[j] = d;
    dot = fmaf(f.attn[j], d, dot);
  }
  float dq[TS_DH];
  for (int i = 0; i < TS_DH; ++i) dq[i] = 0.f;
  for (int j = 0; j < TS_G; ++j) {
    const float ds = scale * f.attn[j] * (da[j] - dot);
    as_[g * TA_RS2 + j] = f.attn[j];
    dss[g * TA_RS2 + j] = ds;
    for (int i = 0; i < TS_DH; ++i) dq[i] = fmaf(ds, ks[j * TA_RS + i], dq[i]);
  }
  __syncthreads();
  // thread g now acts as key / value token j = g:  dk[j] = sum_g' ds[g'][j] q[g'],  dv[j] = sum_g' attn[g'][j] dO[g']
  float dk[TS_DH], dv[TS_DH];
  for (int i = 0; i < TS_DH; ++i) { dk[i] = 0.f; dv[i] = 0.f; }
  for (int r = 0; r < TS_G; ++r) {
    const float ds = dss[r * TA_RS2 + g], a = as_[r * TA_RS2 + g];
    for (int i = 0; i < TS_DH; ++i) {
      dk[i] = fmaf(ds, qs[r * TA_RS + i], dk[i]);
      dv[i] = fmaf(a, dos[r * TA_RS + i], dv[i]);
    }
  }
  // d tok[g][i] = sum_o dq[o] Wq[o][i] + dk[o] Wk[o][i] + dv[o] Wv[o][i]; then through tok = num / (norm + eps)
  float dn = 0.f;
  float* drow = d_rec + (size_t)b * TS_TOKW + ((size_t)h * TS_G + g) * TS_DH;
  for (int i = 0; i < TS_DH; ++i) {
    float t = 0.f;
    for (int o = 0; o < TS_DH; ++o)
      t = fmaf(dq[o], Wq[o * TS_DH + i], fmaf(dk[o], Wk[o * TS_DH + i], fmaf(dv[o], Wv[o * TS_DH + i], t)));
    drow[i] = t * f.inv;
    dn = fmaf(t, f.tok[i], dn);
  }
  d_rec[(size_t)b * TS_TOKW + TS_TOK + h * TS_G + g] = -dn * f.inv;
  for (int i = 0; i < TS_DH; ++i) {
    dqs[g * TA_RS + i] = dq[i];
    dks[g * TA_RS + i] = dk[i];
    dvs[g * TA_RS + i] = dv[i];
  }
  __syncthreads();
  // weight gradients of this (graph, head): dW[o][i] = sum_g d{q,k,v}[g][o] tok[g][i]; 8 of the 256 entries per thread
  float* wp = wpart + (size_t)blockIdx.x * (3 * TS_DH * TS_DH);
  for (int e = g; e < TS_DH * TS_DH; e += TS_G) {
    const int o = e / TS_DH, i = e % TS_DH;
    float sq = 0.f, sk = 0.f, sv = 0.f;
    for (int r = 0; r < TS_G; ++r) {
      const float t = toks[r * TA_RS + i];
      sq = fmaf(dqs[r * TA_RS + o], t, sq);
      sk = fmaf(dks[r * TA_RS + o], t, sk);
      sv = fmaf(dvs[r * TA_RS + o], t, sv);
    }
    wp[e] = sq;
    wp[TS_DH * TS_DH + e] = sk;
    wp[2 * TS_DH * TS_DH + e] = sv;
  }
}

}  // namespace

extern "C" int fvgn_ts_row_partials(int64_t n) {
  const int64_t want = (n + 31) / 32;
  return (int)(want < 1 ? 1 : (want < kRowCtas ? want : kRowCtas));
}

extern "C" int fvgn_ts_slice_forward(const float* P, const float* Ws, const float* bs, const float* temp, const int32_t* chunks,
                                     int32_t nchunks, float* sw, float* partial, void* stream) {
  if (nchunks <= 0) return FVGN_OK;
  if (!P || !Ws || !bs || !temp || !chunks || !sw || !partial) return FVGN_ERR_NULL;
  if (!fvgn_aligned16(P) || !fvgn_aligned16(sw) || !fvgn_aligned16(partial)) return FVGN_ERR_ALIGN;
  auto kern = ts_slice_kernel<true>;
  if (ts_set_smem(kern, SliceSmem<true>::bytes) != FVGN_OK) return FVGN_ERR_LAUNCH;
  FVGN_LAUNCH(kern, (unsigned)nchunks, TS_NT, SliceSmem<true>::bytes, stream, P, sw, Ws, bs, temp, chunks, partial);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_ts_accumulate(const float* sw, const float* V, const int32_t* chunks, int32_t nchunks, float* partial,
                                  void* stream) {
  if (nchunks <= 0) return FVGN_OK;
  if (!sw || !V || !chunks || !partial) return FVGN_ERR_NULL;
  if (!fvgn_aligned16(V) || !fvgn_aligned16(sw) || !fvgn_aligned16(partial)) return FVGN_ERR_ALIGN;
  auto kern = ts_slice_kernel<false>;
  if (ts_set_smem(kern, SliceSmem<false>::bytes) != FVGN_OK) return FVGN_ERR_LAUNCH;
  FVGN_LAUNCH(kern, (unsigned)nchunks, TS_NT, SliceSmem<false>::bytes, stream, V, const_cast<float*>(sw), (const float*)nullptr,
              (const float*)nullptr, (const float*)nullptr, chunks, partial);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_ts_token_attention_forward(const float* rec, const float* wq, const float* wk, const float* wv, float scale,
                                               int32_t nb, float* tok_out, void* stream) {
  if (nb <= 0) return FVGN_OK;
  if (!rec || !wq || !wk || !wv || !tok_out) return FVGN_ERR_NULL;
  FVGN_LAUNCH(ts_token_attention_fwd_kernel, (unsigned)(nb * TS_HEADS), TS_G, 0, stream, rec, wq, wk, wv, scale, tok_out);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_ts_token_attention_backward(const float* rec, const float* wq, const float* wk, const float* wv, float scale,
                                                int32_t nb, const float* d_tok_out, float* d_rec, float* w_partial, void* stream) {
  if (nb <= 0) return FVGN_OK;
  if (!rec || !wq || !wk || !wv || !d_tok_out || !d_rec || !w_partial) return FVGN_ERR_NULL;
  FVGN_LAUNCH(ts_token_attention_bwd_kernel, (unsigned)(nb * TS_HEADS), TS_G, 0, stream, rec, wq, wk, wv, scale, d_tok_out, d_rec,
              w_partial);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_ts_deslice(const float* sw, const float* tok, int64_t tok_ld, int32_t tok_mod, const int32_t* chunks,
                               int32_t nchunks, float* out, void* stream) {
  if (nchunks <= 0) return FVGN_OK;
  if (!sw || !tok || !chunks || !out) return FVGN_ERR_NULL;
  if (tok_mod < 1 || tok_ld < TS_TOK || (tok_ld & 3)) return FVGN_ERR_SHAPE;
  if (!fvgn_aligned16(sw) || !fvgn_aligned16(tok) || !fvgn_aligned16(out)) return FVGN_ERR_ALIGN;
  if (ts_set_smem(ts_deslice_kernel, kSmemDeslice) != FVGN_OK) return FVGN_ERR_LAUNCH;
  FVGN_LAUNCH(ts_deslice_kernel, (unsigned)nchunks, TS_NT, kSmemDeslice, stream, sw, tok, tok_ld, (int)tok_mod, chunks, out);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_ts_slice_backward(const float* P, const float* sw, const float* d_out, const float* tok_out,
                                      const float* d_tok, int32_t tok_mod, const float* Ws, const float* bs, const float* temp,
                                      const int32_t* chunks, int32_t nchunks, float* dP, float* partial, void* stream) {
  if (nchunks <= 0) return FVGN_OK;
  if (!P || !sw || !d_out || !tok_out || !d_tok || !Ws || !bs || !temp || !chunks || !dP || !partial) return FVGN_ERR_NULL;
  if (tok_mod < 1) return FVGN_ERR_SHAPE;
  if (!fvgn_aligned16(P) || !fvgn_aligned16(sw) || !fvgn_aligned16(d_out) || !fvgn_aligned16(tok_out) ||
      !fvgn_aligned16(d_tok) || !fvgn_aligned16(dP))
    return FVGN_ERR_ALIGN;
  if (ts_set_smem(ts_slice_bwd_kernel, kSmemBwd) != FVGN_OK) return FVGN_ERR_LAUNCH;
  FVGN_LAUNCH(ts_slice_bwd_kernel, (unsigned)nchunks, TS_NT, kSmemBwd, stream, P, sw, d_out, tok_out, d_tok, (int)tok_mod, Ws,
              bs, temp, chunks, dP, partial);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_ts_residual_ln_forward(const float* a, const float* bias, const float* res, const float* gamma,
                                           const float* beta, float* y, float* z, float* stats, int64_t n, void* stream) {
  if (n <= 0) return FVGN_OK;
  if (!a || !bias || !res || !gamma || !beta || !y || !z || !stats) return FVGN_ERR_NULL;
  FVGN_LAUNCH(ts_res_ln_fwd_kernel, (unsigned)fvgn_ts_row_partials(n), TS_NT, 0, stream, a, bias, res, gamma, beta, y, z, stats,
              n);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_ts_residual_ln_backward(const float* dz, const float* y, const float* stats, const float* gamma,
                                            const float* d_y_in, float* d_y, float* partial, int64_t n, void* stream) {
  if (n <= 0) return FVGN_OK;
  if (!dz || !y || !stats || !gamma || !d_y || !partial) return FVGN_ERR_NULL;
  FVGN_LAUNCH(ts_res_ln_bwd_kernel, (unsigned)fvgn_ts_row_partials(n), TS_NT, 0, stream, dz, y, stats, gamma, d_y_in, d_y,
              partial, n);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_ts_bias_gelu_forward(const float* hpre, const float* bias, float* h, int64_t n, void* stream) {
  if (n <= 0) return FVGN_OK;
  if (!hpre || !bias || !h) return FVGN_ERR_NULL;
  const int64_t n4 = n * 64;
  FVGN_LAUNCH_SEQ(ts_bias_gelu_fwd_kernel, (unsigned)((n4 + TS_NT - 1) / TS_NT), TS_NT, 0, stream, hpre, bias, h, n4);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_ts_bias_gelu_backward(const float* dh, const float* hpre, const float* bias, float* dhpre, float* partial,
                                          int64_t n, void* stream) {
  if (n <= 0) return FVGN_OK;
  if (!dh || !hpre || !bias || !dhpre || !partial) return FVGN_ERR_NULL;
  FVGN_LAUNCH(ts_bias_gelu_bwd_kernel, (unsigned)fvgn_ts_row_partials(n), TS_NT, 0, stream, dh, hpre, bias, dhpre, partial, n);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_ts_bias_residual(const float* a, const float* bias, const float* res, float* out, void* outh,
                                     int32_t outh_type, int64_t n, void* stream) {
  if (n <= 0) return FVGN_OK;
  if (!a || !bias || !res || !out) return FVGN_ERR_NULL;
#ifdef FVGN_EMU
  if (outh) return FVGN_ERR_UNSUPPORTED;
#endif
  const int64_t n4 = n * 32;
  FVGN_LAUNCH_SEQ(ts_bias_res_kernel, (unsigned)((n4 + TS_NT - 1) / TS_NT), TS_NT, 0, stream, a, bias, res, out,
                  reinterpret_cast<uint16_t*>(outh), (int)(outh_type == FVGN_T_F16), n4);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}
