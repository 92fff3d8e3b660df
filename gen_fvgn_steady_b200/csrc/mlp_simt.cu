// Fused 3-layer MLP blocks, fp32 SIMT path (FVGN_PREC_FP32): the parity mode (rel 1e-5 vs the
// reference's CPU fp32).  build_mlp (EPD.py:10-33): Linear-GELU-Linear-GELU-Linear [+LayerNorm 1e-5].
//
// One CTA = a tile of TM=32 rows; the input row is assembled in shared memory straight from its
// sources (gathers by edge endpoint, concatenations and the relative edge features of
// importer.py:54-78 are never materialised in HBM), the three GEMMs are chained through shared
// memory, LayerNorm + residual are the epilogue.  Backward recomputes the hidden activations from
// the saved block inputs, runs dgrad/wgrad on the same tile and accumulates weight gradients into
// per-CTA partial buffers (static tile->CTA map => deterministic), reduced by a second kernel.
#include "common.cuh"

namespace {

constexpr int TM = 32;      // rows per tile
constexpr int NT = 256;     // threads per CTA
constexpr int LDH = 132;    // padded leading dimension of the [TM][128] activation buffers
constexpr int KC = 16;      // k-chunk of the staged weight tile

template <int MODE> struct Cfg;
template <> struct Cfg<FVGN_MLP_EDGE> { static constexpr int K1 = 384, NOUT = 128; static constexpr bool LN = true; };
template <> struct Cfg<FVGN_MLP_NODE> { static constexpr int K1 = 192, NOUT = 128; static constexpr bool LN = true; };
template <> struct Cfg<FVGN_MLP_ENC_NODE> { static constexpr int K1 = 12, NOUT = 128; static constexpr bool LN = true; };
template <> struct Cfg<FVGN_MLP_ENC_EDGE> { static constexpr int K1 = 15, NOUT = 128; static constexpr bool LN = true; };
template <> struct Cfg<FVGN_MLP_DEC> { static constexpr int K1 = 128, NOUT = 3; static constexpr bool LN = false; };

__host__ __device__ constexpr int k1pad(int k1) { return (k1 + KC - 1) / KC * KC; }
__host__ __device__ constexpr int ldx_of(int k1) { return k1pad(k1) + 4; }

__host__ __device__ inline int64_t param_count(int k1, int nout, bool ln) {
  return (int64_t)128 * k1 + 128 + 128 * 128 + 128 + (int64_t)nout * 128 + nout + (ln ? 256 : 0);
}

// ---------------------------------------------------------------------------------------------
// acc[2][8] = sum_k In[r][k] * Wt[k][c]   (rows r = ty*2+{0,1}; cols c = tx*4+{0..3} and 64+tx*4+{0..3})
// TRANS  : Wt[k][c] = Wg[c*ldw + k]          (y = x W^T, W is [128][K])
// !TRANS : Wt[k][c] = Wg[k*ldw + c0 + c]     (dX = dZ W, k over the 128 output features), zero where c0+c >= cmax
// In is a shared-memory tile zero-padded to a multiple of KC columns.
template <bool TRANS>
__device__ __forceinline__ void gemm_tile(const float* In, int ldin, int K, const float* __restrict__ Wg, int ldw, int c0,
                                          int cmax, float* Ws, float (&acc)[2][8]) {
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += KC) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * NT;
      float v;
      if (TRANS) {
        const int c = idx >> 4, kk = idx & 15;
        // K may be the zero-padded width of the input tile (layer 1 of the encoders: 12 / 15 -> 16): never read past a
        // weight row -- the pad columns multiply zeros, but 0 x (NaN bits past the end of the tensor) is NaN
        v = (k0 + kk < K && k0 + kk < ldw) ? Wg[(size_t)c * ldw + k0 + kk] : 0.f;
        Ws[kk * LDH + c] = v;
      } else {
        const int kk = idx >> 7, c = idx & 127;
        v = (c0 + c < cmax) ? Wg[(size_t)(k0 + kk) * ldw + c0 + c] : 0.f;
        Ws[kk * LDH + c] = v;
      }
    }
    __syncthreads();
    const float* a0p = In + (ty * 2) * ldin + k0;
    const float* a1p = a0p + ldin;
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float a0 = a0p[kk], a1 = a1p[kk];
      const float4 w0 = ld4(Ws + kk * LDH + tx * 4);
      const float4 w1 = ld4(Ws + kk * LDH + 64 + tx * 4);
      acc[0][0] = fmaf(a0, w0.x, acc[0][0]); acc[0][1] = fmaf(a0, w0.y, acc[0][1]);
      acc[0][2] = fmaf(a0, w0.z, acc[0][2]); acc[0][3] = fmaf(a0, w0.w, acc[0][3]);
      acc[0][4] = fmaf(a0, w1.x, acc[0][4]); acc[0][5] = fmaf(a0, w1.y, acc[0][5]);
      acc[0][6] = fmaf(a0, w1.z, acc[0][6]); acc[0][7] = fmaf(a0, w1.w, acc[0][7]);
      acc[1][0] = fmaf(a1, w0.x, acc[1][0]); acc[1][1] = fmaf(a1, w0.y, acc[1][1]);
      acc[1][2] = fmaf(a1, w0.z, acc[1][2]); acc[1][3] = fmaf(a1, w0.w, acc[1][3]);
      acc[1][4] = fmaf(a1, w1.x, acc[1][4]); acc[1][5] = fmaf(a1, w1.y, acc[1][5]);
      acc[1][6] = fmaf(a1, w1.z, acc[1][6]); acc[1][7] = fmaf(a1, w1.w, acc[1][7]);
    }
  }
}

__device__ __forceinline__ int col_of(int tx, int j) { return (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

// hidden layer epilogue: Z = acc + b ; H = gelu(Z)
__device__ __forceinline__ void store_hidden(const float (&acc)[2][8], const float* __restrict__ bias, float* Z, float* Hh) {
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = col_of(tx, j);
      const float z = acc[i][j] + bias[c];
      if (Z) Z[(ty * 2 + i) * LDH + c] = z;
      Hh[(ty * 2 + i) * LDH + c] = gelu_exact(z);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// Assemble the input tile into Xs[TM][LDX] (zero padded in rows >= rows and cols >= K1).
template <int MODE>
__device__ __forceinline__ void load_input_tile(const fvgn_mlp_desc& d, int64_t row0, float* Xs) {
  constexpr int K1 = Cfg<MODE>::K1, LDX = ldx_of(K1), KP = k1pad(K1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (MODE == FVGN_MLP_EDGE) {
    for (int r = warp; r < TM; r += NT / 32) {
      const int64_t row = row0 + r;
      float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
      if (row < d.rows) {
        const int s = d.idx_s[row], rr = d.idx_r[row];
        a = ld4(d.in0 + (size_t)s * 128 + lane * 4);
        b = ld4(d.in0 + (size_t)rr * 128 + lane * 4);
        c = ld4(d.in1 + (size_t)row * 128 + lane * 4);
      }
      st4(Xs + r * LDX + lane * 4, a);
      st4(Xs + r * LDX + 128 + lane * 4, b);
      st4(Xs + r * LDX + 256 + lane * 4, c);
    }
  } else if (MODE == FVGN_MLP_NODE) {
    for (int r = warp; r < TM; r += NT / 32) {
      const int64_t row = row0 + r;
      float4 x = make_float4(0, 0, 0, 0);
      float2 a = make_float2(0, 0);
      if (row < d.rows) {
        a = *reinterpret_cast<const float2*>(d.in0 + (size_t)row * 64 + lane * 2);
        x = ld4(d.in1 + (size_t)row * 128 + lane * 4);
      }
      *reinterpret_cast<float2*>(Xs + r * LDX + lane * 2) = a;
      st4(Xs + r * LDX + 64 + lane * 4, x);
    }
  } else if (MODE == FVGN_MLP_DEC) {
    for (int r = warp; r < TM; r += NT / 32) {
      const int64_t row = row0 + r;
      float4 x = make_float4(0, 0, 0, 0);
      if (row < d.rows) x = ld4(d.in0 + (size_t)row * 128 + lane * 4);
      st4(Xs + r * LDX + lane * 4, x);
    }
  } else if (MODE == FVGN_MLP_ENC_NODE) {
    for (int idx = tid; idx < TM * KP; idx += NT) {
      const int r = idx / KP, k = idx % KP;
      const int64_t row = row0 + r;
      Xs[r * LDX + k] = (row < d.rows && k < K1) ? d.in0[(size_t)row * 12 + k] : 0.f;
    }
  } else {  // ENC_EDGE: [xn[s]-xn[r] (12) | pos[s]-pos[r] (2) | ||dpos|| (1)]  importer.py:54-78
    for (int idx = tid; idx < TM * KP; idx += NT) {
      const int r = idx / KP, k = idx % KP;
      const int64_t row = row0 + r;
      float v = 0.f;
      if (row < d.rows && k < K1) {
        const int s = d.idx_s[row], rr = d.idx_r[row];
        if (k < 12) {
          v = d.in0[(size_t)s * 12 + k] - d.in0[(size_t)rr * 12 + k];
        } else {
          const float dx = d.in1[(size_t)s * 2] - d.in1[(size_t)rr * 2];
          const float dy = d.in1[(size_t)s * 2 + 1] - d.in1[(size_t)rr * 2 + 1];
          v = (k == 12) ? dx : (k == 13) ? dy : sqrtf(dx * dx + dy * dy);
        }
      }
      Xs[r * LDX + k] = v;
    }
  }
}

// =============================================================================================
// forward
template <int MODE>
__global__ void __launch_bounds__(NT) mlp_fwd_kernel(const fvgn_mlp_desc d) {
  constexpr int K1 = Cfg<MODE>::K1, LDX = ldx_of(K1), KP = k1pad(K1), NOUT = Cfg<MODE>::NOUT;
  FVGN_DYN_SMEM(smem_raw);
  float* Xs = reinterpret_cast<float*>(smem_raw);
  float* Ha = Xs + TM * LDX;
  float* Hb = Ha + TM * LDH;
  float* Ws = Hb + TM * LDH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, ty = tid >> 4, tx = tid & 15;
  const int64_t ntiles = (d.rows + TM - 1) / TM;
  float acc[2][8];
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TM;
    __syncthreads();
    load_input_tile<MODE>(d, row0, Xs);
    gemm_tile<true>(Xs, LDX, KP, d.w1, K1, 0, 0, Ws, acc);
    store_hidden(acc, d.b1, nullptr, Ha);
    gemm_tile<true>(Ha, LDH, 128, d.w2, 128, 0, 0, Ws, acc);
    store_hidden(acc, d.b2, nullptr, Hb);
    if (NOUT == 128) {
      gemm_tile<true>(Hb, LDH, 128, d.w3, 128, 0, 0, Ws, acc);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = col_of(tx, j);
          Ha[(ty * 2 + i) * LDH + c] = acc[i][j] + d.b3[c];
        }
      __syncthreads();
      // LayerNorm(128, eps 1e-5) + residual; one warp per row, 4 columns per lane
      const float4 g = ld4(d.ln_g + lane * 4), be = ld4(d.ln_b + lane * 4);
      for (int r = warp; r < TM; r += NT / 32) {
        const int64_t row = row0 + r;
        float4 v = ld4(Ha + r * LDH + lane * 4);
        const float mu = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / 128.0f);
        v.x -= mu; v.y -= mu; v.z -= mu; v.w -= mu;
        const float var = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w) * (1.0f / 128.0f);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        float4 yv = make_float4(v.x * rstd * g.x + be.x, v.y * rstd * g.y + be.y, v.z * rstd * g.z + be.z,
                                v.w * rstd * g.w + be.w);
        if (row < d.rows) {
          if (d.out) st4(d.out + (size_t)row * 128 + lane * 4, yv);
          if (MODE == FVGN_MLP_EDGE && d.out_res) {
            st4(d.out_res + (size_t)row * 128 + lane * 4, add4(ld4(Xs + r * LDX + 256 + lane * 4), yv));
          } else if (MODE == FVGN_MLP_NODE && d.out_res) {
            st4(d.out_res + (size_t)row * 128 + lane * 4, add4(ld4(Xs + r * LDX + 64 + lane * 4), yv));
          }
        }
      }
    } else {  // decoder: 3 outputs per row, no LN
      __syncthreads();
      if (tid < TM * 3) {
        const int r = tid / 3, o = tid % 3;
        const int64_t row = row0 + r;
        float s = 0.f;
        for (int k = 0; k < 128; ++k) s = fmaf(Hb[r * LDH + k], d.w3[o * 128 + k], s);
        if (row < d.rows) d.out[(size_t)row * 3 + o] = s + d.b3[o];
      }
    }
  }
}

// =============================================================================================
// backward
// P[o*ldp + i0+i] += sum_r A[r][o] * Bm[r][i0+i]  for o<128, i<128, i0+i<imax  (thread: 8 o x 8 i)
__device__ __forceinline__ void wgrad_tile(const float* A, int lda, const float* Bm, int ldb, int i0, int imax, float* P,
                                           int ldp) {
  const int tid = threadIdx.x, to = tid >> 4, ti = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  const int ia = i0 + ti * 4, ib = i0 + 64 + ti * 4;
  for (int r = 0; r < TM; ++r) {
    const float4 a0 = ld4(A + r * lda + to * 8), a1 = ld4(A + r * lda + to * 8 + 4);
    float bv[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      bv[q] = (ia + q < imax) ? Bm[r * ldb + ia + q] : 0.f;
      bv[4 + q] = (ib + q < imax) ? Bm[r * ldb + ib + q] : 0.f;
    }
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    float* prow = P + (size_t)(to * 8 + a) * ldp;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (ia + q < imax) prow[ia + q] += acc[a][q];
      if (ib + q < imax) prow[ib + q] += acc[a][4 + q];
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(NT) mlp_bwd_kernel(const fvgn_mlp_desc d) {
  constexpr int K1 = Cfg<MODE>::K1, LDX = ldx_of(K1), KP = k1pad(K1), NOUT = Cfg<MODE>::NOUT;
  constexpr bool LN = Cfg<MODE>::LN;
  constexpr bool NEED_DX = (MODE == FVGN_MLP_EDGE || MODE == FVGN_MLP_NODE || MODE == FVGN_MLP_DEC);
  FVGN_DYN_SMEM(smem_raw);
  float* Xs = reinterpret_cast<float*>(smem_raw);
  float* Z1 = Xs + TM * LDX;
  float* Z2 = Z1 + TM * LDH;
  float* H1 = Z2 + TM * LDH;
  float* H2 = H1 + TM * LDH;
  float* D1 = H2 + TM * LDH;
  float* D2 = D1 + TM * LDH;
  float* Ws = D2 + TM * LDH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, ty = tid >> 4, tx = tid & 15;
  const int64_t ntiles = (d.rows + TM - 1) / TM;
  const int64_t PC = param_count(K1, NOUT, LN);
  const bool resid = !(d.flags & FVGN_MLP_NO_RESIDUAL);
  float* P = d.partials + (size_t)blockIdx.x * PC;
  float* Pw1 = P;
  float* Pb1 = Pw1 + 128 * K1;
  float* Pw2 = Pb1 + 128;
  float* Pb2 = Pw2 + 128 * 128;
  float* Pw3 = Pb2 + 128;
  float* Pb3 = Pw3 + NOUT * 128;
  float* Pg = Pb3 + NOUT;
  float* Pbeta = Pg + 128;
  for (int64_t i = tid; i < PC; i += NT) P[i] = 0.f;
  // register accumulators (persist over the tiles of this CTA)
  float db1 = 0.f, db2 = 0.f, db3 = 0.f;           // thread c < 128 owns column c (c < NOUT for db3)
  float dg[4] = {0, 0, 0, 0}, dbt[4] = {0, 0, 0, 0};  // LN: lane owns cols lane*4.., summed over this warp's rows
  float acc[2][8];
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TM;
    __syncthreads();
    // ---- recompute forward
    load_input_tile<MODE>(d, row0, Xs);
    gemm_tile<true>(Xs, LDX, KP, d.w1, K1, 0, 0, Ws, acc);
    store_hidden(acc, d.b1, Z1, H1);
    gemm_tile<true>(H1, LDH, 128, d.w2, 128, 0, 0, Ws, acc);
    store_hidden(acc, d.b2, Z2, H2);
    if (LN) {
      gemm_tile<true>(H2, LDH, 128, d.w3, 128, 0, 0, Ws, acc);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = col_of(tx, j);
          D1[(ty * 2 + i) * LDH + c] = acc[i][j] + d.b3[c];
        }
      __syncthreads();
      // ---- upstream gradient + LayerNorm backward -> dY in D2
      const float4 g = ld4(d.ln_g + lane * 4);
      for (int r = warp; r < TM; r += NT / 32) {
        const int64_t row = row0 + r;
        float4 go = make_float4(0, 0, 0, 0);
        if (row < d.rows) {
          go = ld4(d.d_out + (size_t)row * 128 + lane * 4);
          if (MODE == FVGN_MLP_EDGE && d.d_gather) {
            const int node = (lane < 16) ? d.idx_s[row] : d.idx_r[row];
            go = add4(go, ld4(d.d_gather + (size_t)node * 64 + (lane & 15) * 4));
          }
        }
        float4 v = ld4(D1 + r * LDH + lane * 4);
        const float mu = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / 128.0f);
        v.x -= mu; v.y -= mu; v.z -= mu; v.w -= mu;
        const float var = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w) * (1.0f / 128.0f);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float4 xh = make_float4(v.x * rstd, v.y * rstd, v.z * rstd, v.w * rstd);
        dg[0] += go.x * xh.x; dg[1] += go.y * xh.y; dg[2] += go.z * xh.z; dg[3] += go.w * xh.w;
        dbt[0] += go.x; dbt[1] += go.y; dbt[2] += go.z; dbt[3] += go.w;
        const float4 dx = make_float4(go.x * g.x, go.y * g.y, go.z * g.z, go.w * g.w);
        const float m1 = warp_sum(dx.x + dx.y + dx.z + dx.w) * (1.0f / 128.0f);
        const float m2 = warp_sum(dx.x * xh.x + dx.y * xh.y + dx.z * xh.z + dx.w * xh.w) * (1.0f / 128.0f);
        st4(D2 + r * LDH + lane * 4, make_float4(rstd * (dx.x - m1 - xh.x * m2), rstd * (dx.y - m1 - xh.y * m2),
                                                 rstd * (dx.z - m1 - xh.z * m2), rstd * (dx.w - m1 - xh.w * m2)));
      }
      __syncthreads();
      // ---- layer 3: dW3 += dY^T H2 ; db3 ; dZ2 = (dY W3) * gelu'(Z2) -> D1
      wgrad_tile(D2, LDH, H2, LDH, 0, 128, Pw3, 128);
      if (tid < 128) {
        float s = 0.f;
        for (int r = 0; r < TM; ++r) s += D2[r * LDH + tid];
        db3 += s;
      }
      gemm_tile<false>(D2, LDH, 128, d.w3, 128, 0, 128, Ws, acc);
    } else {
      // decoder: dY [TM][3] straight from d_out; stash it in D2[r][0..2]
      __syncthreads();
      if (tid < TM * 3) {
        const int r = tid / 3, o = tid % 3;
        const int64_t row = row0 + r;
        D2[r * LDH + o] = (row < d.rows) ? d.d_out[(size_t)row * 3 + o] : 0.f;
      }
      __syncthreads();
      for (int idx = tid; idx < 3 * 128; idx += NT) {
        const int o = idx >> 7, i = idx & 127;
        float s = 0.f;
        for (int r = 0; r < TM; ++r) s = fmaf(D2[r * LDH + o], H2[r * LDH + i], s);
        Pw3[idx] += s;
      }
      if (tid < 3) {
        float s = 0.f;
        for (int r = 0; r < TM; ++r) s += D2[r * LDH + tid];
        db3 += s;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = col_of(tx, j), r = ty * 2 + i;
          acc[i][j] = D2[r * LDH + 0] * d.w3[c] + D2[r * LDH + 1] * d.w3[128 + c] + D2[r * LDH + 2] * d.w3[256 + c];
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = col_of(tx, j), r = ty * 2 + i;
        D1[r * LDH + c] = acc[i][j] * gelu_grad(Z2[r * LDH + c]);
      }
    __syncthreads();
    // ---- layer 2: dW2 += dZ2^T H1 ; db2 ; dZ1 = (dZ2 W2) * gelu'(Z1) -> D2
    wgrad_tile(D1, LDH, H1, LDH, 0, 128, Pw2, 128);
    if (tid < 128) {
      float s = 0.f;
      for (int r = 0; r < TM; ++r) s += D1[r * LDH + tid];
      db2 += s;
    }
    gemm_tile<false>(D1, LDH, 128, d.w2, 128, 0, 128, Ws, acc);
    __syncthreads();  // every warp is done reading D2 (dY) in the LN/decoder branch and wgrad3
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = col_of(tx, j), r = ty * 2 + i;
        D2[r * LDH + c] = acc[i][j] * gelu_grad(Z1[r * LDH + c]);
      }
    __syncthreads();
    // ---- layer 1: dW1 += dZ1^T X ; db1 ; dX = dZ1 W1
    for (int i0 = 0; i0 < K1; i0 += 128) wgrad_tile(D2, LDH, Xs, LDX, i0, K1, Pw1, K1);
    if (tid < 128) {
      float s = 0.f;
      for (int r = 0; r < TM; ++r) s += D2[r * LDH + tid];
      db1 += s;
    }
    if (NEED_DX) {
      for (int i0 = 0; i0 < K1; i0 += 128) {
        gemm_tile<false>(D2, LDH, 128, d.w1, K1, i0, K1, Ws, acc);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int64_t row = row0 + ty * 2 + i;
          if (row >= d.rows) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = i0 + h * 64 + tx * 4;  // first of 4 consecutive input columns
            if (c >= K1) continue;
            float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
            if (MODE == FVGN_MLP_EDGE) {
              if (c < 256) {
                st4(d.d_in0 + (size_t)row * 256 + c, v);
              } else {
                if (resid) v = add4(v, ld4(d.d_out + (size_t)row * 128 + (c - 256)));
                st4(d.d_in1 + (size_t)row * 128 + (c - 256), v);
              }
            } else if (MODE == FVGN_MLP_NODE) {
              if (c < 64) {
                st4(d.d_in0 + (size_t)row * 64 + c, v);
              } else {
                if (resid) v = add4(v, ld4(d.d_out + (size_t)row * 128 + (c - 64)));
                st4(d.d_in1 + (size_t)row * 128 + (c - 64), v);
              }
            } else {
              st4(d.d_in0 + (size_t)row * 128 + c, v);
            }
          }
        }
      }
    }
  }
  // ---- flush register accumulators
  __syncthreads();
  if (tid < 128) {
    Pb1[tid] = db1;
    Pb2[tid] = db2;
    if (tid < NOUT) Pb3[tid] = db3;
  }
  if (LN) {
    float* red = Ws;  // [8 warps][128] x 2
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      red[warp * 128 + lane * 4 + q] = dg[q];
      red[1024 + warp * 128 + lane * 4 + q] = dbt[q];
    }
    __syncthreads();
    if (tid < 128) {
      float s = 0.f, t = 0.f;
      for (int w = 0; w < NT / 32; ++w) {
        s += red[w * 128 + tid];
        t += red[1024 + w * 128 + tid];
      }
      Pg[tid] = s;
      Pbeta[tid] = t;
    }
  }
}

__global__ void __launch_bounds__(256) partial_reduce_kernel(const float* __restrict__ partials, int n_partials, int64_t pc,
                                                             float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= pc) return;
  float s = 0.f;
  for (int g = 0; g < n_partials; ++g) s += partials[(size_t)g * pc + i];
  out[i] = s;
}

template <int MODE> constexpr size_t fwd_smem() { return sizeof(float) * (TM * ldx_of(Cfg<MODE>::K1) + 2 * TM * LDH + KC * LDH); }
template <int MODE> constexpr size_t bwd_smem() { return sizeof(float) * (TM * ldx_of(Cfg<MODE>::K1) + 6 * TM * LDH + KC * LDH); }

constexpr int kMaxCtasFwd = 148 * 2;
constexpr int kMaxCtasBwd = 148;

template <int MODE>
int launch_fwd(const fvgn_mlp_desc& d, void* stream) {
  const int64_t ntiles = (d.rows + TM - 1) / TM;
  const unsigned grid = (unsigned)(ntiles < kMaxCtasFwd ? ntiles : kMaxCtasFwd);
  auto kern = mlp_fwd_kernel<MODE>;
#ifndef FVGN_EMU
  static bool attr_set[FVGN_MAX_DEV] = {false};  // the attribute is per device
  const int dev = fvgn_cur_device();
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<MODE>()) != cudaSuccess)
      return FVGN_ERR_LAUNCH;
    attr_set[dev] = true;
  }
#endif
  FVGN_LAUNCH(kern, grid, NT, fwd_smem<MODE>(), stream, d);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

template <int MODE>
int launch_bwd(const fvgn_mlp_desc& d, void* stream) {
  auto kern = mlp_bwd_kernel<MODE>;
#ifndef FVGN_EMU
  static bool attr_set[FVGN_MAX_DEV] = {false};  // the attribute is per device
  const int dev = fvgn_cur_device();
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem<MODE>()) != cudaSuccess)
      return FVGN_ERR_LAUNCH;
    attr_set[dev] = true;
  }
#endif
  FVGN_LAUNCH(kern, (unsigned)d.n_partials, NT, bwd_smem<MODE>(), stream, d);
  FVGN_CHECK_LAUNCH();
  const int64_t pc = param_count(Cfg<MODE>::K1, Cfg<MODE>::NOUT, Cfg<MODE>::LN);
  FVGN_LAUNCH_SEQ(partial_reduce_kernel, (unsigned)((pc + 255) / 256), 256, 0, stream, d.partials, d.n_partials, pc,
                  d.d_params);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

}  // namespace

int fvgn_mlp_simt_partials(int64_t rows) {
  const int64_t ntiles = (rows + TM - 1) / TM;
  return (int)(ntiles < kMaxCtasBwd ? (ntiles < 1 ? 1 : ntiles) : kMaxCtasBwd);
}

int64_t fvgn_mlp_param_count_impl(int32_t mode) {
  switch (mode) {
    case FVGN_MLP_EDGE: return param_count(384, 128, true);
    case FVGN_MLP_NODE: return param_count(192, 128, true);
    case FVGN_MLP_ENC_NODE: return param_count(12, 128, true);
    case FVGN_MLP_ENC_EDGE: return param_count(15, 128, true);
    case FVGN_MLP_DEC: return param_count(128, 3, false);
  }
  return -1;
}

int fvgn_mlp_forward_simt(const fvgn_mlp_desc* d, void* stream) {
  switch (d->mode) {
    case FVGN_MLP_EDGE: return launch_fwd<FVGN_MLP_EDGE>(*d, stream);
    case FVGN_MLP_NODE: return launch_fwd<FVGN_MLP_NODE>(*d, stream);
    case FVGN_MLP_ENC_NODE: return launch_fwd<FVGN_MLP_ENC_NODE>(*d, stream);
    case FVGN_MLP_ENC_EDGE: return launch_fwd<FVGN_MLP_ENC_EDGE>(*d, stream);
    case FVGN_MLP_DEC: return launch_fwd<FVGN_MLP_DEC>(*d, stream);
  }
  return FVGN_ERR_UNSUPPORTED;
}

int fvgn_mlp_backward_simt(const fvgn_mlp_desc* d, void* stream) {
  switch (d->mode) {
    case FVGN_MLP_EDGE: return launch_bwd<FVGN_MLP_EDGE>(*d, stream);
    case FVGN_MLP_NODE: return launch_bwd<FVGN_MLP_NODE>(*d, stream);
    case FVGN_MLP_ENC_NODE: return launch_bwd<FVGN_MLP_ENC_NODE>(*d, stream);
    case FVGN_MLP_ENC_EDGE: return launch_bwd<FVGN_MLP_ENC_EDGE>(*d, stream);
    case FVGN_MLP_DEC: return launch_bwd<FVGN_MLP_DEC>(*d, stream);
  }
  return FVGN_ERR_UNSUPPORTED;
}
