// Backward of the fused 3-layer MLP blocks on tcgen05 tensor cores (FVGN_PREC_BF16).
//
// The fp32 weight-gradient accumulators of one block (128 x (K1+128+128) x 4 B = up to 327 KB) exceed the 256 KB of
// TMEM of one SM, so the backward is two persistent kernels that meet through a bf16 tile image of dZ1 in HBM:
//
//   kernel A (upper): per 128-row tile  recompute Z1,H1,Z2,H2,Y  ->  LayerNorm backward  ->  dW3 += dY^T H2,
//                     dH2 = dY W3, dZ2 = dH2 * gelu'(Z2), dW2 += dZ2^T H1, dH1 = dZ2 W2, dZ1 = dH1 * gelu'(Z1);
//                     dW2/dW3 live in TMEM for the whole kernel (256 columns), gelu' factors are parked in TMEM as bf16,
//                     activations live in three 32 KB shared-memory buffers in the canonical SWIZZLE_128B layout, which
//                     is read K-major by the recompute/dgrad MMAs and MN-major (K = rows) by the wgrad MMAs.
//                     dZ1 leaves as a pre-swizzled 32 KB tile image (one cp.async.bulk store).
//   kernel B (lower): per tile  dW1 += dZ1^T X (X re-gathered by the producer warps, 64 columns at a time),
//                     dX = dZ1 W1 -> scattered to the block inputs' gradients (+ residual gradient);
//                     dW1 lives in TMEM (up to 384 columns), W1 stays resident in shared memory.
// Bias / LayerNorm-parameter gradients are column sums done by the epilogue warps (shared-memory column sums of the
// bf16 tiles, warp transpose-reduce for the fp32 LayerNorm terms).  Every CTA owns one partial buffer; tile->CTA is
// static, so the final reduction (partial_reduce) is deterministic.
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int NTHREADS = 288;
constexpr int BUF_BYTES = 2 * KB_BYTES;  // one [128 x 128] bf16 activation tile

template <int MODE> struct BCfg;
template <> struct BCfg<FVGN_MLP_EDGE> { static constexpr int K1 = 384, K1P = 384, NOUT = 128; static constexpr bool LN = true, DX = true; };
template <> struct BCfg<FVGN_MLP_NODE> { static constexpr int K1 = 192, K1P = 192, NOUT = 128; static constexpr bool LN = true, DX = true; };
template <> struct BCfg<FVGN_MLP_ENC_NODE> { static constexpr int K1 = 12, K1P = 16, NOUT = 128; static constexpr bool LN = true, DX = false; };
template <> struct BCfg<FVGN_MLP_ENC_EDGE> { static constexpr int K1 = 15, K1P = 16, NOUT = 128; static constexpr bool LN = true, DX = false; };
template <> struct BCfg<FVGN_MLP_DEC> { static constexpr int K1 = 128, K1P = 128, NOUT = 3; static constexpr bool LN = false, DX = true; };

__host__ __device__ inline int64_t pcount(int k1, int nout, bool ln) {
  return (int64_t)128 * k1 + 128 + 128 * 128 + 128 + (int64_t)nout * 128 + nout + (ln ? 256 : 0);
}

constexpr uint32_t IDESC_KK = make_idesc(128, 0, 0);
constexpr uint32_t IDESC_MM = make_idesc(128, 1, 1);
constexpr uint32_t IDESC_KM = make_idesc(128, 0, 1);
constexpr uint32_t IDESC_MM64 = make_idesc(64, 1, 1);
constexpr uint32_t IDESC_KM64 = make_idesc(64, 0, 1);

// write 16 consecutive columns [c0, c0+16) of row `row` as bf16 into a [128x128] swizzled tile
__device__ __forceinline__ void store_tile16(uint8_t* buf, int row, int c0, const uint32_t (&w)[8]) {
  const int kb = c0 >> 6, chunk = (c0 & 63) >> 3;
  *reinterpret_cast<uint4*>(buf + kb * KB_BYTES + sw128_off(row, chunk)) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(buf + kb * KB_BYTES + sw128_off(row, chunk + 1)) = make_uint4(w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ void load_tile16(const uint8_t* buf, int row, int c0, uint32_t (&w)[8]) {
  const int kb = c0 >> 6, chunk = (c0 & 63) >> 3;
  const uint4 a = *reinterpret_cast<const uint4*>(buf + kb * KB_BYTES + sw128_off(row, chunk));
  const uint4 b = *reinterpret_cast<const uint4*>(buf + kb * KB_BYTES + sw128_off(row, chunk + 1));
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
// column sums of a bf16 tile: thread t owns column pair (2p, 2p+1), p = t & 63, over rows [64*(t>>6), +64)
__device__ __forceinline__ void tile_colsum(const uint8_t* buf, int t, float& s0, float& s1) {
  const int p = t & 63, r0 = (t >> 6) * 64;
  const int kb = (2 * p) >> 6, chunk = ((2 * p) & 63) >> 3, word = p & 3;
  const uint8_t* base = buf + kb * KB_BYTES + word * 4;
  float a = 0.f, b = 0.f;
#pragma unroll 8
  for (int r = r0; r < r0 + 64; ++r) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(base + sw128_off(r, chunk));
    a += bf16_lo(w);
    b += bf16_hi(w);
  }
  s0 += a;
  s1 += b;
}

// =============================================================================================== kernel A
template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) mlp_tc_bwd_a_kernel(const fvgn_mlp_desc d) {
  using C = BCfg<MODE>;
  constexpr int NKB1 = nkb1(C::K1P);
  constexpr int LASTK = (C::K1P - 64 * (NKB1 - 1)) / 16;
  constexpr int NSTAGE = 2;
  constexpr uint32_t DW2 = 0, DW3 = 128, WACC = 256, G1 = 384, G2 = 448;
  FVGN_DYN_SMEM(smem);
  uint8_t* w23 = smem;                              // W2 image | W3 image (64 KB)
  uint8_t* ring = w23 + 4 * KB_BYTES;               // NSTAGE x [X chunk | W1 K-block]
  uint8_t* bufH1 = ring + NSTAGE * 2 * KB_BYTES;
  uint8_t* bufH2 = bufH1 + BUF_BYTES;
  uint8_t* bufC = bufH2 + BUF_BYTES;
  float* sb1 = reinterpret_cast<float*>(bufC + BUF_BYTES);
  float* sb2 = sb1 + 128;
  float* sb3 = sb2 + 128;
  float* sg = sb3 + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sg + 128);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_W = 0, B_FULL = 1, B_EMPTY = 3, B_MMA = 5, B_EPI = 6, B_DO = 7, B_CFREE = 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (d.rows + TILE_M - 1) / TILE_M;
  const int64_t PC = pcount(C::K1, C::NOUT, C::LN);
  float* P = d.partials + (size_t)blockIdx.x * PC;
  float* Pb1 = P + 128 * C::K1;
  float* Pw2 = Pb1 + 128;
  float* Pb2 = Pw2 + 128 * 128;
  float* Pw3 = Pb2 + 128;
  float* Pb3 = Pw3 + C::NOUT * 128;
  float* Pg = Pb3 + C::NOUT;
  float* Pbeta = Pg + 128;
  uint8_t* dz_img = reinterpret_cast<uint8_t*>(d.workspace);
  const uint8_t* w_img = reinterpret_cast<const uint8_t*>(d.w_bf16);

  if (tid == 0) {
    mbar_init(BAR(B_W), 1);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(BAR(B_FULL + s), 5);  // 4 producer warps + the expect_tx arrival of the W1 K-block copy
      mbar_init(BAR(B_EMPTY + s), 1);
    }
    mbar_init(BAR(B_MMA), 1);
    mbar_init(BAR(B_EPI), 128);
    mbar_init(BAR(B_DO), 4);     // producers: upstream-gradient tile parked in bufC
    mbar_init(BAR(B_CFREE), 1);  // tcgen05.commit: bufC no longer read by the previous tile's MMAs
    fence_barrier_init();
  }
  for (int i = tid; i < 128; i += NTHREADS) {
    sb1[i] = d.b1[i];
    sb2[i] = d.b2[i];
    sb3[i] = (i < C::NOUT) ? d.b3[i] : 0.f;
    sg[i] = C::LN ? d.ln_g[i] : 1.f;
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      mbar_expect_tx(BAR(B_W), 4 * KB_BYTES);
      for (int i = 0; i < 4; ++i)
        bulk_g2s(smem_u32(w23 + i * KB_BYTES), w_img + (size_t)(NKB1 + i) * KB_BYTES, KB_BYTES, BAR(B_W));
      mbar_wait(BAR(B_W), 0);
      const uint32_t w2s = smem_u32(w23), w3s = w2s + 2 * KB_BYTES;
      const uint32_t h1s = smem_u32(bufH1), h2s = smem_u32(bufH2), cs = smem_u32(bufC);
      uint32_t it = 0, pe = 0;
      bool first = true;
      auto wait_epi = [&]() {
        mbar_wait(BAR(B_EPI), pe);
        pe ^= 1;
        tc_fence_after();
      };
      // D[acc] = A(K-major tile at a_s) * B(K-major weight image at b_s)^T over K = 128
      auto gemm_kk = [&](uint32_t acc, uint32_t a_s, uint32_t b_s) {
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem + acc, make_desc_k128(a_s + (k >> 2) * KB_BYTES) + 2 * (k & 3),
                  make_desc_k128(b_s + (k >> 2) * KB_BYTES) + 2 * (k & 3), IDESC_KK, k != 0);
      };
      // dW[acc] += A^T B with both tiles read MN-major (K = the 128 rows)
      auto gemm_wgrad = [&](uint32_t acc, uint32_t a_s, uint32_t b_s, bool fresh) {
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem + acc, make_desc_mn128(a_s, KB_BYTES) + 128 * k, make_desc_mn128(b_s, KB_BYTES) + 128 * k, IDESC_MM,
                  !(fresh && k == 0));
      };
      // dH[acc] = dZ(K-major tile) * W(weight image read MN-major: N = input feature, K = output feature)
      auto gemm_dgrad = [&](uint32_t acc, uint32_t a_s, uint32_t w_s) {
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem + acc, make_desc_k128(a_s + (k >> 2) * KB_BYTES) + 2 * (k & 3), make_desc_mn128(w_s, KB_BYTES) + 128 * k,
                  IDESC_KM, k != 0);
      };
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // R1: Z1 = X W1^T  (X chunk and W1 K-block arrive together in a ring stage)
        for (int kb = 0; kb < NKB1; ++kb, ++it) {
          const int s = it % NSTAGE;
          mbar_wait(BAR(B_FULL + s), (it / NSTAGE) & 1);
          tc_fence_after();
          const uint64_t ad = make_desc_k128(smem_u32(ring + s * 2 * KB_BYTES));
          const uint64_t bd = make_desc_k128(smem_u32(ring + s * 2 * KB_BYTES + KB_BYTES));
          const int nk = (kb == NKB1 - 1) ? LASTK : 4;
          for (int k = 0; k < nk; ++k) umma_ss(tmem + WACC, ad + 2 * k, bd + 2 * k, IDESC_KK, (kb | k) != 0);
          umma_commit(BAR(B_EMPTY + s));
        }
        umma_commit(BAR(B_MMA));
        wait_epi();                       // E1: H1 in bufH1, gelu'(Z1) in TMEM
        gemm_kk(WACC, h1s, w2s);          // R2
        umma_commit(BAR(B_MMA));
        wait_epi();                       // E2: H2 in bufH2
        gemm_kk(WACC, h2s, w3s);          // R3
        umma_commit(BAR(B_MMA));
        wait_epi();                       // E3: dY in bufC
        gemm_wgrad(DW3, cs, h2s, first);  // dW3 += dY^T H2
        gemm_dgrad(WACC, cs, w3s);        // dH2 = dY W3
        umma_commit(BAR(B_CFREE));
        umma_commit(BAR(B_MMA));
        wait_epi();                       // E4: dZ2 in bufH2
        gemm_wgrad(DW2, h2s, h1s, first); // dW2 += dZ2^T H1
        gemm_dgrad(WACC, h2s, w2s);       // dH1 = dZ2 W2
        umma_commit(BAR(B_MMA));
        wait_epi();                       // E5: dZ1 written, WACC drained
        first = false;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ============================================================ producers: X chunk (gather + bf16), W1 K-block (bulk copy)
    // and the tile of upstream gradients dO = d_out (+ gathered d_a1) as bf16 into bufC
    const int pw = warp - 4;
    uint32_t it = 0, tcount = 0;
    TileIdx idx;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int64_t row0 = tile * TILE_M;
      load_tile_idx<MODE>(d, row0, pw, lane, idx);
      for (int kb = 0; kb < NKB1; ++kb, ++it) {
        const int s = it % NSTAGE;
        uint8_t* stage = ring + s * 2 * KB_BYTES;
        mbar_wait(BAR(B_EMPTY + s), ((it / NSTAGE) & 1) ^ 1);
        if (pw == 0 && lane == 0) {
          mbar_expect_tx(BAR(B_FULL + s), KB_BYTES);
          bulk_g2s(smem_u32(stage + KB_BYTES), w_img + (size_t)kb * KB_BYTES, KB_BYTES, BAR(B_FULL + s));
        }
        produce_chunk<MODE>(d, row0, kb, stage, pw, lane, idx);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_FULL + s));
      }
      mbar_wait(BAR(B_CFREE), (tcount & 1) ^ 1);  // previous tile's dW3 / dH2 MMAs have retired
      if (C::LN) {
        const int seg = lane & 7;
#pragma unroll 1
        for (int kb = 0; kb < 2; ++kb) {
          float4 lo[8], hi[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t row = row0 + i * 16 + pw * 4 + (lane >> 3);
            lo[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            hi[i] = lo[i];
            if (row < d.rows) {
              const float* p = d.d_out + (size_t)row * 128 + kb * 64 + seg * 8;
              lo[i] = __ldg(reinterpret_cast<const float4*>(p));
              hi[i] = __ldg(reinterpret_cast<const float4*>(p + 4));
              if (MODE == FVGN_MLP_EDGE && d.d_gather) {
                const float* g = d.d_gather + (size_t)(kb == 0 ? idx.s[i] : idx.r[i]) * 64 + seg * 8;
                const float4 a = __ldg(reinterpret_cast<const float4*>(g)), b = __ldg(reinterpret_cast<const float4*>(g + 4));
                lo[i] = make_float4(lo[i].x + a.x, lo[i].y + a.y, lo[i].z + a.z, lo[i].w + a.w);
                hi[i] = make_float4(hi[i].x + b.x, hi[i].y + b.y, hi[i].z + b.z, hi[i].w + b.w);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rloc = i * 16 + pw * 4 + (lane >> 3);
            *reinterpret_cast<uint4*>(bufC + kb * KB_BYTES + sw128_off(rloc, seg)) =
                make_uint4(pack_bf16(lo[i].x, lo[i].y), pack_bf16(lo[i].z, lo[i].w), pack_bf16(hi[i].x, hi[i].y),
                           pack_bf16(hi[i].z, hi[i].w));
          }
        }
      } else {
        // decoder: dY = d_out[row, 0:3] zero-padded to 128 columns (one thread per row)
        const int rloc = pw * 32 + lane;
        const int64_t row = row0 + rloc;
        float a = 0.f, b = 0.f, c = 0.f;
        if (row < d.rows) {
          a = d.d_out[(size_t)row * 3];
          b = d.d_out[(size_t)row * 3 + 1];
          c = d.d_out[(size_t)row * 3 + 2];
        }
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(bufC + kb * KB_BYTES + sw128_off(rloc, ch)) =
                (kb == 0 && ch == 0) ? make_uint4(pack_bf16(a, b), pack_bf16(c, 0.f), 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_DO));
    }
  } else {
    // ============================================================ epilogue: thread <-> row of the tile
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int rloc = warp * 32 + lane;
    uint32_t pm = 0;
    auto wait_mma = [&]() {
      mbar_wait(BAR(B_MMA), pm);
      pm ^= 1;
      tc_fence_after();
    };
    auto done = [&]() {
      tc_fence_before();
      mbar_arrive(BAR(B_EPI));
    };
    float db1a = 0.f, db1b = 0.f, db2a = 0.f, db2b = 0.f, db3a = 0.f, db3b = 0.f, dbta = 0.f, dbtb = 0.f;
    float dgam[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dgam[i] = 0.f;
    bool store_pending = false;
    uint32_t pdo = 0;
    const uint32_t wacc = tmem + lane_base + WACC;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      // ---------------- E1 / E2: hidden activations (bf16, shared memory) + gelu' (bf16, TMEM)
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        wait_mma();
        if (layer == 0 && store_pending) {  // bufH1 still feeds the previous tile's dZ1 bulk store
          if (tid == 0) bulk_wait_read0();
          epi_bar_sync();
        }
        uint8_t* hb = layer == 0 ? bufH1 : bufH2;
        const float* bias = layer == 0 ? sb1 : sb2;
        const uint32_t gcol = tmem + lane_base + (layer == 0 ? G1 : G2);
        for_each_chunk16(wacc, [&](int c0, uint32_t (&r)[16]) {
          uint32_t hw[8], gw[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float h0, g0, h1, g1;
            gelu_tanh_pair(__uint_as_float(r[2 * j]) + bias[c0 + 2 * j], h0, g0);
            gelu_tanh_pair(__uint_as_float(r[2 * j + 1]) + bias[c0 + 2 * j + 1], h1, g1);
            hw[j] = pack_bf16(h0, h1);
            gw[j] = pack_bf16(g0, g1);
          }
          store_tile16(hb, rloc, c0, hw);
          tmem_st8(gcol + c0 / 2, gw);
        });
        tmem_wait_st();
        fence_proxy_async();
        done();
      }
      // ---------------- E3: LayerNorm backward -> dY (bf16) in bufC (the producers parked dO there)
      wait_mma();
      mbar_wait(BAR(B_DO), pdo);
      pdo ^= 1;
      if (C::LN) {
        tile_colsum(bufC, tid, dbta, dbtb);  // d beta = column sums of dO
        float sum = 0.f, sq = 0.f;
        for_each_chunk16(wacc, [&](int c0, uint32_t (&r)[16]) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float y = __uint_as_float(r[j]) + sb3[c0 + j];
            sum += y;
            sq = fmaf(y, y, sq);
          }
        });
        const float mean = sum * (1.0f / 128.0f);
        const float rstd = rsqrtf(fmaxf(sq * (1.0f / 128.0f) - mean * mean, 0.f) + 1e-5f);
        float m1 = 0.f, m2 = 0.f;
        // sweep A: row moments and d gamma
        for_each_chunk16(wacc, [&](int c0, uint32_t (&r)[16]) {
          uint32_t ow[8];
          float gx[16];
          load_tile16(bufC, rloc, c0, ow);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh0 = (__uint_as_float(r[2 * j]) + sb3[c0 + 2 * j] - mean) * rstd;
            const float xh1 = (__uint_as_float(r[2 * j + 1]) + sb3[c0 + 2 * j + 1] - mean) * rstd;
            const float o0 = bf16_lo(ow[j]), o1 = bf16_hi(ow[j]);
            const float dx0 = o0 * sg[c0 + 2 * j], dx1 = o1 * sg[c0 + 2 * j + 1];
            m1 += dx0 + dx1;
            m2 = fmaf(dx0, xh0, fmaf(dx1, xh1, m2));
            gx[2 * j] = o0 * xh0;
            gx[2 * j + 1] = o1 * xh1;
          }
          const float cg = warp_colsum16(gx, lane);
#pragma unroll
          for (int i = 0; i < 8; ++i) dgam[i] += (i == (c0 >> 4)) ? cg : 0.f;  // static register indexing
        });
        m1 *= (1.0f / 128.0f);
        m2 *= (1.0f / 128.0f);
        epi_bar_sync();  // every thread has finished reading the whole dO tile (d beta) before rows are overwritten
        // sweep B: dY = rstd * (dO*gamma - m1 - xhat*m2)
        for_each_chunk16(wacc, [&](int c0, uint32_t (&r)[16]) {
          uint32_t ow[8];
          load_tile16(bufC, rloc, c0, ow);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh0 = (__uint_as_float(r[2 * j]) + sb3[c0 + 2 * j] - mean) * rstd;
            const float xh1 = (__uint_as_float(r[2 * j + 1]) + sb3[c0 + 2 * j + 1] - mean) * rstd;
            const float y0 = rstd * (bf16_lo(ow[j]) * sg[c0 + 2 * j] - m1 - xh0 * m2);
            const float y1 = rstd * (bf16_hi(ow[j]) * sg[c0 + 2 * j + 1] - m1 - xh1 * m2);
            ow[j] = pack_bf16(y0, y1);
          }
          store_tile16(bufC, rloc, c0, ow);
        });
        fence_proxy_async();
        epi_bar_sync();
      }
      tile_colsum(bufC, tid, db3a, db3b);
      done();
      // ---------------- E4: dZ2 = dH2 * gelu'(Z2) -> bufH2
      wait_mma();
      for_each_chunk16(wacc, [&](int c0, uint32_t (&r)[16]) {
        uint32_t g[8], ow[8];
        tmem_ld8(tmem + lane_base + G2 + c0 / 2, g);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          ow[j] = pack_bf16(__uint_as_float(r[2 * j]) * bf16_lo(g[j]), __uint_as_float(r[2 * j + 1]) * bf16_hi(g[j]));
        store_tile16(bufH2, rloc, c0, ow);
      });
      fence_proxy_async();
      epi_bar_sync();
      tile_colsum(bufH2, tid, db2a, db2b);
      done();
      // ---------------- E5: dZ1 = dH1 * gelu'(Z1) -> bufH1 -> HBM tile image
      wait_mma();
      for_each_chunk16(wacc, [&](int c0, uint32_t (&r)[16]) {
        uint32_t g[8], ow[8];
        tmem_ld8(tmem + lane_base + G1 + c0 / 2, g);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          ow[j] = pack_bf16(__uint_as_float(r[2 * j]) * bf16_lo(g[j]), __uint_as_float(r[2 * j + 1]) * bf16_hi(g[j]));
        store_tile16(bufH1, rloc, c0, ow);
      });
      fence_proxy_async();
      epi_bar_sync();
      if (tid == 0) bulk_s2g(dz_img + (size_t)tile * BUF_BYTES, smem_u32(bufH1), BUF_BYTES);
      store_pending = true;
      tile_colsum(bufH1, tid, db1a, db1b);
      done();
    }
    // ---------------- flush: weight-gradient accumulators (TMEM) and the column sums -> this CTA's partial buffer
    if (tid == 0) bulk_wait0();
    epi_bar_sync();
    tc_fence_after();
    {
      const int o = rloc;  // TMEM lane = output feature
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + lane_base + DW2 + c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) Pw2[(size_t)o * 128 + c0 + j] = __uint_as_float(r[j]);
        tmem_ld16(tmem + lane_base + DW3 + c0, r);
        tmem_wait_ld();
        if (o < C::NOUT) {
#pragma unroll
          for (int j = 0; j < 16; ++j) Pw3[(size_t)o * 128 + c0 + j] = __uint_as_float(r[j]);
        }
      }
    }
    // bias sums: combine the two row halves through shared memory (bufC is free now)
    float* scr = reinterpret_cast<float*>(bufC);
    {
      const int p = tid & 63, half = tid >> 6;
      scr[half * 128 + 2 * p] = db1a; scr[half * 128 + 2 * p + 1] = db1b;
      scr[256 + half * 128 + 2 * p] = db2a; scr[256 + half * 128 + 2 * p + 1] = db2b;
      scr[512 + half * 128 + 2 * p] = db3a; scr[512 + half * 128 + 2 * p + 1] = db3b;
      scr[1280 + half * 128 + 2 * p] = dbta; scr[1280 + half * 128 + 2 * p + 1] = dbtb;
      if (C::LN && lane < 16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) scr[768 + warp * 128 + i * 16 + lane] = dgam[i];
      }
    }
    epi_bar_sync();
    {
      Pb1[tid] = scr[tid] + scr[128 + tid];
      Pb2[tid] = scr[256 + tid] + scr[384 + tid];
      if (tid < C::NOUT) Pb3[tid] = scr[512 + tid] + scr[640 + tid];
      if (C::LN) {
        Pg[tid] = scr[768 + tid] + scr[896 + tid] + scr[1024 + tid] + scr[1152 + tid];
        Pbeta[tid] = scr[1280 + tid] + scr[1408 + tid];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =============================================================================================== kernel B
constexpr int STG_LD = 20;                      // staging row: 16 columns + pad (floats)
constexpr int STG_BYTES = 4 * 32 * STG_LD * 4;  // 4 epilogue warps x 32 rows

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) mlp_tc_bwd_b_kernel(const fvgn_mlp_desc d) {
  using C = BCfg<MODE>;
  constexpr int NKB1 = nkb1(C::K1P);
  constexpr int NSTAGE = 3;
  constexpr uint32_t DW1 = 0, WACC = 384;  // dW1: NKB1 x 64 columns; two 64-column dX accumulators
  FVGN_DYN_SMEM(smem);
  uint8_t* w1 = smem;                               // resident W1 image (NKB1 K-blocks)
  uint8_t* dz = w1 + NKB1 * KB_BYTES;               // 2 x dZ1 tile
  uint8_t* ring = dz + 2 * BUF_BYTES;               // NSTAGE x X chunk
  float* stg = reinterpret_cast<float*>(ring + NSTAGE * KB_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stg) + STG_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_W = 0, B_XFULL = 1, B_XEMPTY = 4, B_DZFULL = 7, B_DZEMPTY = 9, B_AFULL = 11, B_AFREE = 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (d.rows + TILE_M - 1) / TILE_M;
  const int64_t PC = pcount(C::K1, C::NOUT, C::LN);
  float* Pw1 = d.partials + (size_t)blockIdx.x * PC;
  const uint8_t* dz_img = reinterpret_cast<const uint8_t*>(d.workspace);
  const uint8_t* w_img = reinterpret_cast<const uint8_t*>(d.w_bf16);
  const bool resid = !(d.flags & FVGN_MLP_NO_RESIDUAL);

  if (tid == 0) {
    mbar_init(BAR(B_W), 1);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(BAR(B_XFULL + s), 4);
      mbar_init(BAR(B_XEMPTY + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(B_DZFULL + s), 1);
      mbar_init(BAR(B_DZEMPTY + s), 1);
      mbar_init(BAR(B_AFULL + s), 1);
      mbar_init(BAR(B_AFREE + s), 128);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      if (C::DX) {
        mbar_expect_tx(BAR(B_W), NKB1 * KB_BYTES);
        for (int i = 0; i < NKB1; ++i) bulk_g2s(smem_u32(w1 + i * KB_BYTES), w_img + (size_t)i * KB_BYTES, KB_BYTES, BAR(B_W));
        mbar_wait(BAR(B_W), 0);
      }
      uint32_t it = 0, blk = 0, tcount = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
        const int zb = tcount & 1;
        const uint32_t dzs = smem_u32(dz + zb * BUF_BYTES);
        mbar_wait(BAR(B_DZFULL + zb), (tcount >> 1) & 1);
        tc_fence_after();
        // G1: dW1[:, 64 kb .. 64 kb + 64) += dZ1^T X_kb   (A = dZ1 MN-major, B = X chunk MN-major, K = rows)
        for (int kb = 0; kb < NKB1; ++kb, ++it) {
          const int s = it % NSTAGE;
          mbar_wait(BAR(B_XFULL + s), (it / NSTAGE) & 1);
          tc_fence_after();
          const uint32_t xs = smem_u32(ring + s * KB_BYTES);
          for (int k = 0; k < 8; ++k)
            umma_ss(tmem + DW1 + 64 * kb, make_desc_mn128(dzs, KB_BYTES) + 128 * k, make_desc_mn128(xs, KB_BYTES) + 128 * k,
                    IDESC_MM64, !(tcount == 0 && k == 0));
          umma_commit(BAR(B_XEMPTY + s));
        }
        // D1: dX[:, 64 j .. 64 j + 64) = dZ1 W1[:, block j]   (A = dZ1 K-major, B = W1 image block j read MN-major)
        if (C::DX) {
          for (int j = 0; j < NKB1; ++j, ++blk) {
            const int ab = blk & 1;
            mbar_wait(BAR(B_AFREE + ab), ((blk >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t wj = smem_u32(w1 + j * KB_BYTES);
            for (int k = 0; k < 8; ++k)
              umma_ss(tmem + WACC + 64 * ab, make_desc_k128(dzs + (k >> 2) * KB_BYTES) + 2 * (k & 3),
                      make_desc_mn128(wj, KB_BYTES) + 128 * k, IDESC_KM64, k != 0);
            umma_commit(BAR(B_AFULL + ab));
          }
        }
        umma_commit(BAR(B_DZEMPTY + zb));
      }
      // all MMAs retired before the epilogue reads dW1
      umma_commit(BAR(B_W));
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ============================================================ producers: dZ1 tile (bulk copy) + X chunks (gather)
    const int pw = warp - 4;
    uint32_t it = 0, tcount = 0;
    TileIdx idx;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int64_t row0 = tile * TILE_M;
      load_tile_idx<MODE>(d, row0, pw, lane, idx);
      if (pw == 0 && lane == 0) {
        const int zb = tcount & 1;
        mbar_wait(BAR(B_DZEMPTY + zb), ((tcount >> 1) & 1) ^ 1);
        mbar_expect_tx(BAR(B_DZFULL + zb), BUF_BYTES);
        bulk_g2s(smem_u32(dz + zb * BUF_BYTES), dz_img + (size_t)tile * BUF_BYTES, BUF_BYTES, BAR(B_DZFULL + zb));
      }
      for (int kb = 0; kb < NKB1; ++kb, ++it) {
        const int s = it % NSTAGE;
        mbar_wait(BAR(B_XEMPTY + s), ((it / NSTAGE) & 1) ^ 1);
        produce_chunk<MODE>(d, row0, kb, ring + s * KB_BYTES, pw, lane, idx);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_XFULL + s));
      }
    }
  } else {
    // ============================================================ epilogue
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int rloc = warp * 32 + lane;
    float* mystg = stg + warp * 32 * STG_LD;
    if (C::DX) {
      uint32_t blk = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * TILE_M;
        for (int j = 0; j < NKB1; ++j, ++blk) {
          const int ab = blk & 1;
          mbar_wait(BAR(B_AFULL + ab), (blk >> 1) & 1);
          tc_fence_after();
          for_each_chunk16<64>(tmem + lane_base + WACC + 64 * ab, [&](int c0, uint32_t (&r)[16]) {
            const int col0 = 64 * j + c0;  // first input column of this 16-column group
            // which destination does this group go to (uniform over the group: boundaries are multiples of 64)
            float* dst;
            int ld, colo;
            bool add_res = false;
            if (MODE == FVGN_MLP_EDGE) {
              if (col0 < 256) { dst = d.d_in0; ld = 256; colo = col0; }
              else { dst = d.d_in1; ld = 128; colo = col0 - 256; add_res = resid; }
            } else if (MODE == FVGN_MLP_NODE) {
              if (col0 < 64) { dst = d.d_in0; ld = 64; colo = col0; }
              else { dst = d.d_in1; ld = 128; colo = col0 - 64; add_res = resid; }
            } else {
              dst = d.d_in0; ld = 128; colo = col0;
            }
            // residual gradient rows: issue the loads before the staging round trip
            float4 g[4];
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
              const int64_t row = row0 + warp * 32 + pass * 8 + (lane >> 2);
              g[pass] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (add_res && row < d.rows)
                g[pass] = __ldg(reinterpret_cast<const float4*>(d.d_out + (size_t)row * 128 + colo + (lane & 3) * 4));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(mystg + lane * STG_LD + 4 * q) =
                  make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                              __uint_as_float(r[4 * q + 3]));
            __syncwarp();
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
              const int rr = pass * 8 + (lane >> 2), cc = (lane & 3) * 4;
              const int64_t row = row0 + warp * 32 + rr;
              if (row < d.rows) {
                float4 v = *reinterpret_cast<const float4*>(mystg + rr * STG_LD + cc);
                v = make_float4(v.x + g[pass].x, v.y + g[pass].y, v.z + g[pass].z, v.w + g[pass].w);
                *reinterpret_cast<float4*>(dst + (size_t)row * ld + colo + cc) = v;
              }
            }
            __syncwarp();
          });
          tc_fence_before();
          mbar_arrive(BAR(B_AFREE + ab));
        }
      }
    }
    // ---------------- flush dW1 (TMEM lane = output feature o, column = input feature i)
    mbar_wait(BAR(B_W), C::DX ? 1 : 0);
    tc_fence_after();
    {
      const int o = rloc;
#pragma unroll 1
      for (int c0 = 0; c0 < NKB1 * 64; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + lane_base + DW1 + c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < C::K1) Pw1[(size_t)o * C::K1 + c0 + j] = __uint_as_float(r[j]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

__global__ void __launch_bounds__(256) tc_partial_reduce_kernel(const float* __restrict__ partials, int n_partials, int64_t pc,
                                                                float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= pc) return;
  float s = 0.f;
  for (int g = 0; g < n_partials; ++g) s += partials[(size_t)g * pc + i];
  out[i] = s;
}

template <int MODE> constexpr int smem_a() { return 4 * KB_BYTES + 2 * 2 * KB_BYTES + 3 * BUF_BYTES + 4 * 512 + 256; }
template <int MODE> constexpr int smem_b() {
  return nkb1(BCfg<MODE>::K1P) * KB_BYTES + 2 * BUF_BYTES + 3 * KB_BYTES + STG_BYTES + 256;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int MODE>
int launch_tc_bwd(const fvgn_mlp_desc& d, void* stream) {
  using C = BCfg<MODE>;
  auto ka = mlp_tc_bwd_a_kernel<MODE>;
  auto kb = mlp_tc_bwd_b_kernel<MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a<MODE>()) != cudaSuccess) return FVGN_ERR_LAUNCH;
    if (cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b<MODE>()) != cudaSuccess) return FVGN_ERR_LAUNCH;
    attr_set = true;
  }
  const unsigned grid = (unsigned)d.n_partials;
  ka<<<grid, NTHREADS, smem_a<MODE>(), (cudaStream_t)stream>>>(d);
  FVGN_CHECK_LAUNCH();
  kb<<<grid, NTHREADS, smem_b<MODE>(), (cudaStream_t)stream>>>(d);
  FVGN_CHECK_LAUNCH();
  const int64_t pc = pcount(C::K1, C::NOUT, C::LN);
  tc_partial_reduce_kernel<<<(unsigned)((pc + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d.partials, d.n_partials, pc, d.d_params);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

}  // namespace

int fvgn_mlp_tc_partials(int32_t, int64_t rows) {
  const int64_t ntiles = (rows + TILE_M - 1) / TILE_M;
  const int n = num_sms();
  return (int)(ntiles < 1 ? 1 : (ntiles < n ? ntiles : n));
}

int64_t fvgn_mlp_tc_workspace_bytes(int32_t, int64_t rows) {
  const int64_t ntiles = (rows + TILE_M - 1) / TILE_M;
  return (ntiles < 1 ? 1 : ntiles) * (int64_t)BUF_BYTES;
}

int fvgn_mlp_backward_simt(const fvgn_mlp_desc* d, void* stream);

int fvgn_mlp_backward_tc(const fvgn_mlp_desc* d, void* stream) {
  if (!d->w_bf16 || !d->workspace) return FVGN_ERR_NULL;
  if ((((uintptr_t)d->w_bf16) & 15) != 0 || (((uintptr_t)d->workspace) & 1023) != 0) return FVGN_ERR_ALIGN;
  if (d->n_partials != fvgn_mlp_tc_partials(d->mode, d->rows)) return FVGN_ERR_SHAPE;
  if (d->rows == 0) {
    const int64_t pc = fvgn_mlp_param_count(d->mode);
    if (cudaMemsetAsync(d->d_params, 0, (size_t)pc * sizeof(float), (cudaStream_t)stream) != cudaSuccess) return FVGN_ERR_LAUNCH;
    return FVGN_OK;
  }
  switch (d->mode) {
    case FVGN_MLP_EDGE: return launch_tc_bwd<FVGN_MLP_EDGE>(*d, stream);
    case FVGN_MLP_NODE: return launch_tc_bwd<FVGN_MLP_NODE>(*d, stream);
    case FVGN_MLP_ENC_NODE: return launch_tc_bwd<FVGN_MLP_ENC_NODE>(*d, stream);
    case FVGN_MLP_ENC_EDGE: return launch_tc_bwd<FVGN_MLP_ENC_EDGE>(*d, stream);
    case FVGN_MLP_DEC: return launch_tc_bwd<FVGN_MLP_DEC>(*d, stream);
  }
  return FVGN_ERR_UNSUPPORTED;
}
