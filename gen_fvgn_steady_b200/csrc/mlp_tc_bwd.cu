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
#include <type_traits>

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
// =============================================================================================== kernel A
// Roles (448 threads): warps 0-7 epilogue (thread <-> (row, column half): warps w and w+4 share a TMEM lane quarter and
// split the 128 columns), warps 8-11 stage the upstream-gradient tile dO as bf16, warp 12 issues the MMAs, warp 13 is the
// bulk-copy loader (W2/W3 image once, then one 32 KB Z1 tile image per tile, double buffered).
// Layer 1 is NOT recomputed: the forward left Z1 (bf16) in HBM; E1 turns the tile into H1 (in place, shared memory) and
// gelu'(Z1) (TMEM).  Per tile:  E1 -> R2 -> E2 -> R3 -> E3 (LayerNorm backward) -> dW3/dH2 -> E4 -> dW2/dH1 -> E5.
// NEW = number of epilogue warps (8 or 16): the 128 columns of a row are split over NEW / 4 threads.  Measured on B200
// (2 M edges): 16 warps (80 registers per thread) run the kernel in the same 1.17 ms as 8 warps (128 registers) -- the
// epilogue stages are already issue bound with two warps per SM sub-partition, the rest of a tile's time is the
// serialised MMA / barrier hand-offs -- so 8 is the default.
#ifndef FVGN_BWD_A_EPI_WARPS
#define FVGN_BWD_A_EPI_WARPS 8
#endif
constexpr int A_NEW = FVGN_BWD_A_EPI_WARPS;
// FVGN_COLSUM_PACKED = 1 adds four rows of a bias-gradient column sum as packed bf16x2 before going to fp32 (fewer
// instructions, measured: no change of the golden-run gradient errors, 1 % faster); 0 = plain fp32 accumulation.
#ifndef FVGN_COLSUM_PACKED
#define FVGN_COLSUM_PACKED 1
#endif
#ifdef FVGN_TIMING
// debug build: cycle breakdown of kernel A's per-tile chain, summed over the tiles of CTA 0 (thread 0 only)
__device__ unsigned long long g_prof_a[16];
#define PROF_T(i)                                  \
  do {                                             \
    if (blockIdx.x == 0 && tid == 0) {             \
      const long long now_ = clock64();            \
      g_prof_a[i] += (unsigned long long)(now_ - tprev_); \
      tprev_ = now_;                               \
    }                                              \
  } while (0)
#else
#define PROF_T(i) do { } while (0)
#endif
constexpr int A_THREADS = (A_NEW + 6) * 32;
constexpr int A_EPI = A_NEW * 32;  // epilogue threads
template <int N> __device__ __forceinline__ void epi_bar_sync_n() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// column sums of a bf16 tile by the epilogue threads: thread t owns column pair (2p, 2p+1), p = t & 63, rows [ROWS*(t>>6), +ROWS).
// Accumulated in fp32 (or, FVGN_COLSUM_PACKED, four rows at a time as packed bf16x2).
template <int ROWS, class P>
__device__ __forceinline__ void tile_colsum_n(const uint8_t* buf, int t, float& s0, float& s1) {
  const int p = t & 63, r0 = (t >> 6) * ROWS;
  const int kb = (2 * p) >> 6, chunk = ((2 * p) & 63) >> 3, word = p & 3;
  const uint8_t* base = buf + kb * KB_BYTES + word * 4;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int r = r0; r < r0 + ROWS; r += 4) {
    const uint32_t w0 = *reinterpret_cast<const uint32_t*>(base + sw128_off(r, chunk));
    const uint32_t w1 = *reinterpret_cast<const uint32_t*>(base + sw128_off(r + 1, chunk));
    const uint32_t w2 = *reinterpret_cast<const uint32_t*>(base + sw128_off(r + 2, chunk));
    const uint32_t w3 = *reinterpret_cast<const uint32_t*>(base + sw128_off(r + 3, chunk));
#if FVGN_COLSUM_PACKED
    const uint32_t w = P::add2(P::add2(w0, w1), P::add2(w2, w3));  // two extra bf16 roundings per 4 rows
    a += P::lo(w);
    b += P::hi(w);
#else
    a += (P::lo(w0) + P::lo(w1)) + (P::lo(w2) + P::lo(w3));        // exact fp32 accumulation of the bf16 tile
    b += (P::hi(w0) + P::hi(w1)) + (P::hi(w2) + P::hi(w3));
#endif
  }
  s0 += a;
  s1 += b;
}

template <int MODE, class P>
__global__ void __launch_bounds__(A_THREADS, 1) mlp_tc_bwd_a_kernel(const fvgn_mlp_desc d) {
  using C = BCfg<MODE>;
  constexpr uint32_t IDESC_KK = make_idesc(P::FMT, 128, 0, 0), IDESC_MM = make_idesc(P::FMT, 128, 1, 1),
                     IDESC_KM = make_idesc(P::FMT, 128, 0, 1);
  constexpr uint32_t DW2 = 0, DW3 = 128, WACC = 256, G1 = 384, G2 = 448;
  constexpr int NEW = A_NEW;               // epilogue warps
  constexpr int CSPLIT = NEW / 4;          // threads per row (column groups)
  constexpr int COLS = 128 / CSPLIT;       // columns per epilogue thread
  constexpr int CS_ROWS = 128 * 64 / A_EPI;  // rows per thread in a tile column sum
  constexpr int RG = A_EPI / 64;           // row groups of the column sums
  constexpr int W_PROD = NEW, W_MMA = NEW + 4, W_LD = NEW + 5;
  auto epi_bar = [] { epi_bar_sync_n<A_EPI>(); };
  auto colsum = [](const uint8_t* buf, int t, float& s0, float& s1) { tile_colsum_n<CS_ROWS, P>(buf, t, s0, s1); };
  FVGN_DYN_SMEM(smem);
  uint8_t* w23 = smem;                              // W2 image | W3 image (64 KB)
  uint8_t* bufZ = w23 + 4 * KB_BYTES;               // 2 x [Z1 -> H1 -> dZ1] tile
  uint8_t* bufH2 = bufZ + 2 * BUF_BYTES;            // H2 -> dZ2
  uint8_t* bufC = bufH2 + BUF_BYTES;                // dO -> dY
  float* sb2 = reinterpret_cast<float*>(bufC + BUF_BYTES);
  float* sb3 = sb2 + 128;
  float* sg = sb3 + 128;
  float4* xch = reinterpret_cast<float4*>(sg + 128);   // [CSPLIT][128 rows] LayerNorm row sums exchanged between the column groups
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + CSPLIT * 128);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_W = 0, B_ZFULL = 1, B_ZEMPTY = 3, B_MMA = 5, B_EPI = 6, B_DO = 7, B_CFREE = 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (d.rows + TILE_M - 1) / TILE_M;
  const int64_t PC = pcount(C::K1, C::NOUT, C::LN);
  float* Pbase = d.partials + (size_t)blockIdx.x * PC;
  float* Pb1 = Pbase + 128 * C::K1;
  float* Pw2 = Pb1 + 128;
  float* Pb2 = Pw2 + 128 * 128;
  float* Pw3 = Pb2 + 128;
  float* Pb3 = Pw3 + C::NOUT * 128;
  float* Pg = Pb3 + C::NOUT;
  float* Pbeta = Pg + 128;
  uint8_t* dz_img = reinterpret_cast<uint8_t*>(d.workspace);
  const uint8_t* z_img = reinterpret_cast<const uint8_t*>(d.z1_img);
  const uint8_t* w_img = reinterpret_cast<const uint8_t*>(d.w_bf16);
  constexpr int NKB1 = nkb1(C::K1P);

  if (tid == 0) {
    mbar_init(BAR(B_W), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(B_ZFULL + s), 1);          // expect_tx arrival + the bytes of the bulk copy
      mbar_init(BAR(B_ZEMPTY + s), A_EPI + 1); // every epilogue thread (done reading) + thread 0 (bulk store has read it)
    }
    mbar_init(BAR(B_MMA), 1);
    mbar_init(BAR(B_EPI), A_EPI);
    mbar_init(BAR(B_DO), 4);     // producer warps: upstream-gradient tile parked in bufC
    mbar_init(BAR(B_CFREE), 2);  // tcgen05.commit (dW3 / dH2 MMAs retired) + epilogue (db3 column sums done)
    fence_barrier_init();
  }
  for (int i = tid; i < 128; i += A_THREADS) {
    sb2[i] = d.b2[i];
    sb3[i] = (i < C::NOUT) ? d.b3[i] : 0.f;
    sg[i] = C::LN ? d.ln_g[i] : 1.f;
  }
  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == W_LD) {
    // ============================================================ loader
    if (lane == 0) {
      mbar_expect_tx(BAR(B_W), 4 * KB_BYTES);
      for (int i = 0; i < 4; ++i)
        bulk_g2s(smem_u32(w23 + i * KB_BYTES), w_img + (size_t)(NKB1 + i) * KB_BYTES, KB_BYTES, BAR(B_W));
      uint32_t i = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
        const int zb = i & 1;
        mbar_wait(BAR(B_ZEMPTY + zb), ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(BAR(B_ZFULL + zb), BUF_BYTES);
        bulk_g2s(smem_u32(bufZ + zb * BUF_BYTES), z_img + (size_t)tile * BUF_BYTES, BUF_BYTES, BAR(B_ZFULL + zb));
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      mbar_wait(BAR(B_W), 0);
      const uint32_t w2s = smem_u32(w23), w3s = w2s + 2 * KB_BYTES;
      const uint32_t h2s = smem_u32(bufH2), cs = smem_u32(bufC);
      uint32_t pe = 0, i = 0;
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
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
        const uint32_t h1s = smem_u32(bufZ + (i & 1) * BUF_BYTES);
        wait_epi();                       // E1: H1 in bufZ, gelu'(Z1) in TMEM
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
        first = false;
      }
    }
    __syncwarp();
  } else if (warp >= W_PROD) {
    // ============================================================ producers: the tile of upstream gradients
    // dO = d_out (+ gathered d_a1) as bf16 into bufC.  The loads of the next tile are issued (and packed) before waiting
    // for bufC to be released, so their latency is hidden behind the current tile.
    const int pw = warp - W_PROD;
    uint32_t tcount = 0;
    TileIdx idx;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int64_t row0 = tile * TILE_M;
      if (C::LN) {
        const bool gath = MODE == FVGN_MLP_EDGE && (d.d_gather || d.d_gatherh);
        if (gath) load_tile_idx<MODE>(d, row0, pw, lane, idx);
        const int seg = lane & 7;
        // rows i0 .. i0+NR-1 of this thread's 8 (tile rows i*16 + pw*4 + lane/8), 64-column block kb -> packed bf16
        auto load_pack = [&](int kb, int i0, auto nr_tag, uint4* out) {
          constexpr int NR = decltype(nr_tag)::value;
          if (d.d_outh && (!gath || d.d_gatherh)) {
            // 16-bit gradient stream: the rows are the operand format already -- no conversion at all; the gathered d_a1 rows
            // are added as packed pairs (one rounding: both addends are 16-bit values)
#pragma unroll
            for (int j = 0; j < NR; ++j) {
              const int64_t row = row0 + (i0 + j) * 16 + pw * 4 + (lane >> 3);
              out[j] = make_uint4(0u, 0u, 0u, 0u);
              if (row < d.rows)
                out[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(d.d_outh) + (size_t)row * 256 + kb * 128 +
                                                              seg * 16));
            }
            if (gath) {
              uint4 gv[NR];
#pragma unroll
              for (int j = 0; j < NR; ++j) {
                const int64_t row = row0 + (i0 + j) * 16 + pw * 4 + (lane >> 3);
                const int gi = kb == 0 ? idx.sender(i0 + j, lane) : idx.receiver(i0 + j, lane);
                gv[j] = make_uint4(0u, 0u, 0u, 0u);
                if (row < d.rows)
                  gv[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(d.d_gatherh) + (size_t)gi * 128 + seg * 16));
              }
#pragma unroll
              for (int j = 0; j < NR; ++j)
                out[j] = make_uint4(P::add2(out[j].x, gv[j].x), P::add2(out[j].y, gv[j].y), P::add2(out[j].z, gv[j].z),
                                    P::add2(out[j].w, gv[j].w));
            }
            return;
          }
          float4 lo[NR], hi[NR];
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            const int64_t row = row0 + (i0 + j) * 16 + pw * 4 + (lane >> 3);
            lo[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            hi[j] = lo[j];
            if (row < d.rows && d.d_outh) {   // 16-bit gradient stream with an fp32 d_gather (stand-alone use)
              const uint4 w = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(d.d_outh) + (size_t)row * 256 +
                                                                   kb * 128 + seg * 16));
              lo[j] = make_float4(P::lo(w.x), P::hi(w.x), P::lo(w.y), P::hi(w.y));
              hi[j] = make_float4(P::lo(w.z), P::hi(w.z), P::lo(w.w), P::hi(w.w));
            } else if (row < d.rows && d.d_out) {   // neither (EDGE): no upstream gradient but the gathered d_a1
              const float* p = d.d_out + (size_t)row * 128 + kb * 64 + seg * 8;
              lo[j] = __ldg(reinterpret_cast<const float4*>(p));
              hi[j] = __ldg(reinterpret_cast<const float4*>(p + 4));
            }
          }
          if (gath) {
            int gi[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j) gi[j] = kb == 0 ? idx.sender(i0 + j, lane) : idx.receiver(i0 + j, lane);
            if (d.d_gatherh) {
              uint4 gv[NR];
#pragma unroll
              for (int j = 0; j < NR; ++j) {
                const int64_t row = row0 + (i0 + j) * 16 + pw * 4 + (lane >> 3);
                gv[j] = make_uint4(0u, 0u, 0u, 0u);
                if (row < d.rows)
                  gv[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(d.d_gatherh) + (size_t)gi[j] * 128 + seg * 16));
              }
#pragma unroll
              for (int j = 0; j < NR; ++j) {
                lo[j] = make_float4(lo[j].x + P::lo(gv[j].x), lo[j].y + P::hi(gv[j].x), lo[j].z + P::lo(gv[j].y),
                                    lo[j].w + P::hi(gv[j].y));
                hi[j] = make_float4(hi[j].x + P::lo(gv[j].z), hi[j].y + P::hi(gv[j].z), hi[j].z + P::lo(gv[j].w),
                                    hi[j].w + P::hi(gv[j].w));
              }
            } else {
#pragma unroll
              for (int j = 0; j < NR; ++j) {
                const int64_t row = row0 + (i0 + j) * 16 + pw * 4 + (lane >> 3);
                if (row < d.rows) {
                  const float* g = d.d_gather + (size_t)gi[j] * 64 + seg * 8;
                  const float4 a = __ldg(reinterpret_cast<const float4*>(g)), b = __ldg(reinterpret_cast<const float4*>(g + 4));
                  lo[j] = make_float4(lo[j].x + a.x, lo[j].y + a.y, lo[j].z + a.z, lo[j].w + a.w);
                  hi[j] = make_float4(hi[j].x + b.x, hi[j].y + b.y, hi[j].z + b.z, hi[j].w + b.w);
                }
              }
            }
          }
#pragma unroll
          for (int j = 0; j < NR; ++j)
            out[j] = make_uint4(pack16<P>(lo[j].x, lo[j].y), pack16<P>(lo[j].z, lo[j].w), pack16<P>(hi[j].x, hi[j].y),
                                pack16<P>(hi[j].z, hi[j].w));
        };
        auto park = [&](int kb, int i0, int n, const uint4* v) {
          for (int j = 0; j < n; ++j) {
            const int rloc = (i0 + j) * 16 + pw * 4 + (lane >> 3);
            *reinterpret_cast<uint4*>(bufC + kb * KB_BYTES + sw128_off(rloc, seg)) = v[j];
          }
        };
        if (A_NEW <= 8) {
          // registers to spare: both 64-column blocks of the next tile are loaded and packed before bufC is released
          uint4 pk[2][8];
          load_pack(0, 0, std::integral_constant<int, 8>{}, pk[0]);
          load_pack(1, 0, std::integral_constant<int, 8>{}, pk[1]);
          mbar_wait(BAR(B_CFREE), (tcount & 1) ^ 1);  // previous tile's dW3 / dH2 MMAs and db3 column sums are done
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) park(kb, 0, 8, pk[kb]);
        } else {
          // 80 registers per thread: block 0 is prefetched two rows at a time, block 1 follows the release in two halves
          uint4 pk0[8];
#pragma unroll
          for (int i0 = 0; i0 < 8; i0 += 2) load_pack(0, i0, std::integral_constant<int, 2>{}, pk0 + i0);
          mbar_wait(BAR(B_CFREE), (tcount & 1) ^ 1);
          park(0, 0, 8, pk0);
#pragma unroll
          for (int i0 = 0; i0 < 8; i0 += 4) {
            uint4 pk1[4];
            load_pack(1, i0, std::integral_constant<int, 4>{}, pk1);
            park(1, i0, 4, pk1);
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
        mbar_wait(BAR(B_CFREE), (tcount & 1) ^ 1);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(bufC + kb * KB_BYTES + sw128_off(rloc, ch)) =
                (kb == 0 && ch == 0) ? make_uint4(pack16<P>(a, b), pack16<P>(c, 0.f), 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_DO));
    }
  } else {
    // ============================================================ epilogue: thread <-> (row, column half)
    const int q = warp & 3, half = warp >> 2;  // TMEM lane quarter, column group
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int rloc = q * 32 + lane;
    const int cbase = COLS * half;  // first column of this thread's group
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
    float dgam[COLS / 16];
#pragma unroll
    for (int i = 0; i < COLS / 16; ++i) dgam[i] = 0.f;
    uint32_t pdo = 0, i = 0;
    const uint32_t wacc = tmem + lane_base + WACC + cbase;
    const uint32_t g1c = tmem + lane_base + G1 + cbase / 2, g2c = tmem + lane_base + G2 + cbase / 2;

#ifdef FVGN_TIMING
    long long tprev_ = clock64();
#endif
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int zb = i & 1;
      uint8_t* bz = bufZ + zb * BUF_BYTES;
      // ---------------- E1: Z1 (bf16, from the forward) -> H1 in place + gelu'(Z1) into TMEM
      mbar_wait(BAR(B_ZFULL + zb), (i >> 1) & 1);
      PROF_T(0);  // wait Z1
#pragma unroll 1
      for (int c0 = 0; c0 < COLS; c0 += 16) {
        uint32_t zw[8], hw[8], gw[8];
        load_tile16(bz, rloc, cbase + c0, zw);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#if FVGN_F32X2
          float2 h, g;
          gelu_tanh_pair2(unpack16<P>(zw[j]), h, g);
          hw[j] = pack16<P>(h);
          gw[j] = pack16<P>(g);
#else
          float h0, g0, h1, g1;
          gelu_tanh_pair(P::lo(zw[j]), h0, g0);
          gelu_tanh_pair(P::hi(zw[j]), h1, g1);
          hw[j] = pack16<P>(h0, h1);
          gw[j] = pack16<P>(g0, g1);
#endif
        }
        store_tile16(bz, rloc, cbase + c0, hw);
        tmem_st8(g1c + c0 / 2, gw);
      }
      tmem_wait_st();
      fence_proxy_async();
      done();
      if (i > 0 && tid == 0) {
        // the previous tile's dZ1 bulk store has finished reading its buffer: hand it back to the loader
        bulk_wait_read0();
        mbar_arrive(BAR(B_ZEMPTY + (zb ^ 1)));
      }
      PROF_T(1);  // E1
      // ---------------- E2: H2 (bf16, shared memory) + gelu'(Z2) (bf16, TMEM)
      wait_mma();
      PROF_T(2);  // wait R2
      for_each_chunk16<COLS>(wacc, [&](int c0, uint32_t (&r)[16]) {
        uint32_t hw[8], gw[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float4 b = *reinterpret_cast<const float4*>(sb2 + cbase + c0 + 2 * j);
#if FVGN_F32X2
          float2 ha, ga, hb, gb;
          gelu_tanh_pair2(__fadd2_rn(f2u(r[2 * j], r[2 * j + 1]), make_float2(b.x, b.y)), ha, ga);
          gelu_tanh_pair2(__fadd2_rn(f2u(r[2 * j + 2], r[2 * j + 3]), make_float2(b.z, b.w)), hb, gb);
          hw[j] = pack16<P>(ha);
          gw[j] = pack16<P>(ga);
          hw[j + 1] = pack16<P>(hb);
          gw[j + 1] = pack16<P>(gb);
#else
          float h0, g0, h1, g1, h2, g2, h3, g3;
          gelu_tanh_pair(__uint_as_float(r[2 * j]) + b.x, h0, g0);
          gelu_tanh_pair(__uint_as_float(r[2 * j + 1]) + b.y, h1, g1);
          gelu_tanh_pair(__uint_as_float(r[2 * j + 2]) + b.z, h2, g2);
          gelu_tanh_pair(__uint_as_float(r[2 * j + 3]) + b.w, h3, g3);
          hw[j] = pack16<P>(h0, h1);
          gw[j] = pack16<P>(g0, g1);
          hw[j + 1] = pack16<P>(h2, h3);
          gw[j + 1] = pack16<P>(g2, g3);
#endif
        }
        store_tile16(bufH2, rloc, cbase + c0, hw);
        tmem_st8(g2c + c0 / 2, gw);
      });
      tmem_wait_st();
      fence_proxy_async();
      done();
      PROF_T(3);  // E2
      // ---------------- E3: LayerNorm backward -> dY (bf16) in bufC (the producers parked dO there)
      wait_mma();
      PROF_T(4);  // wait R3
      mbar_wait(BAR(B_DO), pdo);
      PROF_T(5);  // wait dO
      pdo ^= 1;
      if (C::LN) {
        colsum(bufC, tid, dbta, dbtb);  // d beta = column sums of dO
        // sweep 1 (one pass, no dependence on the statistics): sum y, sum y^2, S1 = sum dO*gamma, S2 = sum dO*gamma*y
        float sum = 0.f, sq = 0.f, s1 = 0.f, s2 = 0.f;
#if FVGN_F32X2
        float2 sum2 = splat2(0.f), sq2 = splat2(0.f), s12 = splat2(0.f), s22 = splat2(0.f);
#endif
        for_each_chunk16<COLS>(wacc, [&](int c0, uint32_t (&r)[16]) {
          uint32_t ow[8];
          load_tile16(bufC, rloc, cbase + c0, ow);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const int c = cbase + c0 + 2 * j;
            const float4 b = *reinterpret_cast<const float4*>(sb3 + c);
            const float4 gm = *reinterpret_cast<const float4*>(sg + c);
#if FVGN_F32X2
            const float2 ya = __fadd2_rn(f2u(r[2 * j], r[2 * j + 1]), make_float2(b.x, b.y));
            const float2 yb = __fadd2_rn(f2u(r[2 * j + 2], r[2 * j + 3]), make_float2(b.z, b.w));
            const float2 da = __fmul2_rn(unpack16<P>(ow[j]), make_float2(gm.x, gm.y));
            const float2 db = __fmul2_rn(unpack16<P>(ow[j + 1]), make_float2(gm.z, gm.w));
            sum2 = __fadd2_rn(sum2, __fadd2_rn(ya, yb));
            sq2 = __ffma2_rn(ya, ya, __ffma2_rn(yb, yb, sq2));
            s12 = __fadd2_rn(s12, __fadd2_rn(da, db));
            s22 = __ffma2_rn(da, ya, __ffma2_rn(db, yb, s22));
            continue;
#endif
            const float y0 = __uint_as_float(r[2 * j]) + b.x, y1 = __uint_as_float(r[2 * j + 1]) + b.y;
            const float y2 = __uint_as_float(r[2 * j + 2]) + b.z, y3 = __uint_as_float(r[2 * j + 3]) + b.w;
            const float d0 = P::lo(ow[j]) * gm.x, d1 = P::hi(ow[j]) * gm.y;
            const float d2 = P::lo(ow[j + 1]) * gm.z, d3 = P::hi(ow[j + 1]) * gm.w;
            sum += (y0 + y1) + (y2 + y3);
            sq = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, sq))));
            s1 += (d0 + d1) + (d2 + d3);
            s2 = fmaf(d0, y0, fmaf(d1, y1, fmaf(d2, y2, fmaf(d3, y3, s2))));
          }
        });
#if FVGN_F32X2
        sum = sum2.x + sum2.y; sq = sq2.x + sq2.y; s1 = s12.x + s12.y; s2 = s22.x + s22.y;
#endif
        xch[half * 128 + rloc] = make_float4(sum, sq, s1, s2);
        epi_bar();  // also: every thread has finished reading the dO tile (d beta) before rows are overwritten
#pragma unroll
        for (int g = 1; g < CSPLIT; ++g) {
          const float4 o = xch[((half + g) % CSPLIT) * 128 + rloc];
          sum += o.x; sq += o.y; s1 += o.z; s2 += o.w;
        }
        const float mean = sum * (1.0f / 128.0f);
        const float rstd = rsqrtf(fmaxf(sq * (1.0f / 128.0f) - mean * mean, 0.f) + 1e-5f);
        const float nmr = -mean * rstd;
        const float m1 = s1 * (1.0f / 128.0f);                       // mean(dO*gamma)
        const float m2 = rstd * (s2 * (1.0f / 128.0f) - mean * m1);  // mean(dO*gamma*xhat)
        // sweep 2: dY = rstd * (dO*gamma - m1 - xhat*m2), d gamma += dO*xhat
        for_each_chunk16<COLS>(wacc, [&](int c0, uint32_t (&r)[16]) {
          uint32_t ow[8];
          float gx[16];
          load_tile16(bufC, rloc, cbase + c0, ow);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const int c = cbase + c0 + 2 * j;
            const float4 b = *reinterpret_cast<const float4*>(sb3 + c);
            const float4 gm = *reinterpret_cast<const float4*>(sg + c);
#if FVGN_F32X2
            {
              const float2 rs = splat2(rstd), nm = splat2(nmr), nm1 = splat2(-m1), nm2 = splat2(-m2);
              const float2 xa = __ffma2_rn(__fadd2_rn(f2u(r[2 * j], r[2 * j + 1]), make_float2(b.x, b.y)), rs, nm);
              const float2 xb = __ffma2_rn(__fadd2_rn(f2u(r[2 * j + 2], r[2 * j + 3]), make_float2(b.z, b.w)), rs, nm);
              const float2 oa = unpack16<P>(ow[j]), ob = unpack16<P>(ow[j + 1]);
              const float2 ga = __fmul2_rn(oa, xa), gb = __fmul2_rn(ob, xb);
              gx[2 * j] = ga.x; gx[2 * j + 1] = ga.y; gx[2 * j + 2] = gb.x; gx[2 * j + 3] = gb.y;
              // rstd * (o*gamma - m1 - xhat*m2) = rstd * fma(xhat, -m2, fma(o, gamma, -m1))
              const float2 ya = __fmul2_rn(rs, __ffma2_rn(xa, nm2, __ffma2_rn(oa, make_float2(gm.x, gm.y), nm1)));
              const float2 yb = __fmul2_rn(rs, __ffma2_rn(xb, nm2, __ffma2_rn(ob, make_float2(gm.z, gm.w), nm1)));
              ow[j] = pack16<P>(ya);
              ow[j + 1] = pack16<P>(yb);
              continue;
            }
#endif
            const float xh0 = fmaf(__uint_as_float(r[2 * j]) + b.x, rstd, nmr), xh1 = fmaf(__uint_as_float(r[2 * j + 1]) + b.y, rstd, nmr);
            const float xh2 = fmaf(__uint_as_float(r[2 * j + 2]) + b.z, rstd, nmr), xh3 = fmaf(__uint_as_float(r[2 * j + 3]) + b.w, rstd, nmr);
            const float o0 = P::lo(ow[j]), o1 = P::hi(ow[j]), o2 = P::lo(ow[j + 1]), o3 = P::hi(ow[j + 1]);
            gx[2 * j] = o0 * xh0; gx[2 * j + 1] = o1 * xh1; gx[2 * j + 2] = o2 * xh2; gx[2 * j + 3] = o3 * xh3;
            const float y0 = rstd * fmaf(-xh0, m2, fmaf(o0, gm.x, -m1)), y1 = rstd * fmaf(-xh1, m2, fmaf(o1, gm.y, -m1));
            const float y2 = rstd * fmaf(-xh2, m2, fmaf(o2, gm.z, -m1)), y3 = rstd * fmaf(-xh3, m2, fmaf(o3, gm.w, -m1));
            ow[j] = pack16<P>(y0, y1);
            ow[j + 1] = pack16<P>(y2, y3);
          }
          store_tile16(bufC, rloc, cbase + c0, ow);
          const float cg = warp_colsum16(gx, lane);
#pragma unroll
          for (int k = 0; k < COLS / 16; ++k) dgam[k] += (k == (c0 >> 4)) ? cg : 0.f;  // static register indexing
        });
        fence_proxy_async();
      }
      done();
      epi_bar();                        // the whole dY tile is written
      colsum(bufC, tid, db3a, db3b);    // overlaps the dW3 / dH2 MMAs
      PROF_T(6);  // E3 (+ db3 column sums)
      // ---------------- E4: dZ2 = dH2 * gelu'(Z2) -> bufH2
      wait_mma();
      PROF_T(7);  // wait dW3/dH2
      for_each_chunk16<COLS>(wacc, [&](int c0, uint32_t (&r)[16]) {
        uint32_t g[8], ow[8];
        tmem_ld8(g2c + c0 / 2, g);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)
#if FVGN_F32X2
          ow[j] = pack16<P>(__fmul2_rn(f2u(r[2 * j], r[2 * j + 1]), unpack16<P>(g[j])));
#else
          ow[j] = pack16<P>(__uint_as_float(r[2 * j]) * P::lo(g[j]), __uint_as_float(r[2 * j + 1]) * P::hi(g[j]));
#endif
        store_tile16(bufH2, rloc, cbase + c0, ow);
      });
      fence_proxy_async();
      done();
      epi_bar();                         // every thread is past its db3 column sums; the dZ2 tile is complete
      if (tid == 0) mbar_arrive(BAR(B_CFREE));
      colsum(bufH2, tid, db2a, db2b);    // overlaps the dW2 / dH1 MMAs
      PROF_T(8);  // E4 (+ db2 column sums)
      // ---------------- E5: dZ1 = dH1 * gelu'(Z1) -> bufZ (in place over H1) -> HBM tile image
      wait_mma();
      PROF_T(9);  // wait dW2/dH1
      for_each_chunk16<COLS>(wacc, [&](int c0, uint32_t (&r)[16]) {
        uint32_t g[8], ow[8];
        tmem_ld8(g1c + c0 / 2, g);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)
#if FVGN_F32X2
          ow[j] = pack16<P>(__fmul2_rn(f2u(r[2 * j], r[2 * j + 1]), unpack16<P>(g[j])));
#else
          ow[j] = pack16<P>(__uint_as_float(r[2 * j]) * P::lo(g[j]), __uint_as_float(r[2 * j + 1]) * P::hi(g[j]));
#endif
        store_tile16(bz, rloc, cbase + c0, ow);
      });
      fence_proxy_async();
      tc_fence_before();
      epi_bar();
      if (tid == 0) bulk_s2g(dz_img + (size_t)tile * BUF_BYTES, smem_u32(bz), BUF_BYTES);
      colsum(bz, tid, db1a, db1b);
      mbar_arrive(BAR(B_ZEMPTY + zb));  // this thread no longer reads the buffer (thread 0 adds the bulk store's release)
      PROF_T(10);  // E5 (+ db1 column sums)
#ifdef FVGN_TIMING
      if (blockIdx.x == 0 && tid == 0) g_prof_a[15] += 1;
#endif
    }
    // ---------------- flush: weight-gradient accumulators (TMEM) and the column sums -> this CTA's partial buffer
    if (tid == 0) bulk_wait0();
    epi_bar();
    tc_fence_after();
    {
      const int o = rloc;  // TMEM lane = output feature
#pragma unroll 1
      for (int c0 = 0; c0 < COLS; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + lane_base + DW2 + cbase + c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) Pw2[(size_t)o * 128 + cbase + c0 + j] = __uint_as_float(r[j]);
        tmem_ld16(tmem + lane_base + DW3 + cbase + c0, r);
        tmem_wait_ld();
        if (o < C::NOUT) {
#pragma unroll
          for (int j = 0; j < 16; ++j) Pw3[(size_t)o * 128 + cbase + c0 + j] = __uint_as_float(r[j]);
        }
      }
    }
    // bias sums: combine the RG row groups through shared memory (bufC is free now)
    float* scr = reinterpret_cast<float*>(bufC);
    constexpr int VS = RG * 128;  // floats per bias vector in the scratch: [RG row groups][128 columns]
    {
      const int p = tid & 63, rq = tid >> 6;
      scr[rq * 128 + 2 * p] = db1a; scr[rq * 128 + 2 * p + 1] = db1b;
      scr[VS + rq * 128 + 2 * p] = db2a; scr[VS + rq * 128 + 2 * p + 1] = db2b;
      scr[2 * VS + rq * 128 + 2 * p] = db3a; scr[2 * VS + rq * 128 + 2 * p + 1] = db3b;
      scr[3 * VS + rq * 128 + 2 * p] = dbta; scr[3 * VS + rq * 128 + 2 * p + 1] = dbtb;
      if (C::LN && lane < 16) {
        // dgam[k] of lane L: column cbase + 16 k + L, summed over this warp's 32 rows
#pragma unroll
        for (int k = 0; k < COLS / 16; ++k) scr[4 * VS + q * 128 + cbase + k * 16 + lane] = dgam[k];
      }
    }
    epi_bar();
    if (tid < 128) {
      float a1 = 0.f, a2 = 0.f, a3 = 0.f, ab = 0.f, ag = 0.f;
#pragma unroll
      for (int g = 0; g < RG; ++g) {
        a1 += scr[g * 128 + tid];
        a2 += scr[VS + g * 128 + tid];
        a3 += scr[2 * VS + g * 128 + tid];
        ab += scr[3 * VS + g * 128 + tid];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) ag += scr[4 * VS + g * 128 + tid];
      Pb1[tid] = a1;
      Pb2[tid] = a2;
      if (tid < C::NOUT) Pb3[tid] = a3;
      if (C::LN) {
        Pg[tid] = ag;
        Pbeta[tid] = ab;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =============================================================================================== kernel B
constexpr int STG_BYTES = 4 * WSTG_BYTES;  // one wide staging tile per epilogue warp

// KB0: first 64-column block of the layer-1 input this kernel handles.  0: everything.  4 (EDGE, node-level layer-1 path,
// mlp_tc_bwd_node.cu): only the e columns 256..383 -- dW1[:, 256:384] += dZ1^T e, d_e = dZ1 W1[:, 256:384] (+ residual
// gradient); the agg[s] | agg[r] columns are differentiated per node by mlp_tc_bwd_node_kernel, so no gathered operand
// chunks, no [E,256] gradient stream.
// GS: gradient streams of the residual path.  0 = fp32 in (d_out), fp32 out (d_in1); 1 = 16-bit in (d_outh), 16-bit out
// (d_in1h); 2 = decided at run time (mixed: the first / last block of a 16-bit chain).  Compile-time for the two pure cases so
// that the residual-row loads of the write-out keep their single predicated form (batched ahead of the accumulator wait).
template <int MODE, class P, int KB0, int GS = 0>
__global__ void __launch_bounds__(NTHREADS, 1) mlp_tc_bwd_b_kernel(const fvgn_mlp_desc d) {
  using C = BCfg<MODE>;
  constexpr uint32_t IDESC_MM64 = make_idesc(P::FMT, 64, 1, 1), IDESC_KM64 = make_idesc(P::FMT, 64, 0, 1);
  constexpr int NKB1 = nkb1(C::K1P);
  constexpr int NSTAGE = 3;
  constexpr uint32_t DW1 = 0, WACC = 384;  // dW1: NKB1 x 64 columns; two 64-column dX accumulators
  FVGN_DYN_SMEM(smem);
  uint8_t* w1 = smem;                               // resident W1 image (NKB1 K-blocks)
  uint8_t* dz = w1 + NKB1 * KB_BYTES;               // 2 x dZ1 tile
  uint8_t* ring = dz + 2 * BUF_BYTES;               // NSTAGE x X chunk
  uint8_t* stg = ring + NSTAGE * KB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + STG_BYTES);
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
  const bool in16 = GS == 2 ? d.d_outh != nullptr : GS == 1, out16 = GS == 2 ? d.d_in1h != nullptr : GS == 1;
  const bool resid = !(d.flags & FVGN_MLP_NO_RESIDUAL) && (in16 ? d.d_outh != nullptr : d.d_out != nullptr);
  const uint8_t* resid_h = in16 ? reinterpret_cast<const uint8_t*>(d.d_outh) : nullptr;   // 16-bit gradient stream (else fp32 d_out)

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
        for (int kb = KB0; kb < NKB1; ++kb, ++it) {
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
          for (int j = KB0; j < NKB1; ++j, ++blk) {
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
    // ============================================================ producers: dZ1 tile (bulk copy) + X chunks
    const int pw = warp - 4;
    uint32_t it = 0;
    // the dZ1 tile image of the CTA's t-th tile goes into buffer t & 1
    auto fetch_dz = [&](uint32_t tcount) {
      if (pw == 0 && lane == 0) {
        const int zb = tcount & 1;
        const int64_t tile = blockIdx.x + (int64_t)tcount * gridDim.x;
        mbar_wait(BAR(B_DZEMPTY + zb), ((tcount >> 1) & 1) ^ 1);
        mbar_expect_tx(BAR(B_DZFULL + zb), BUF_BYTES);
        bulk_g2s(smem_u32(dz + zb * BUF_BYTES), dz_img + (size_t)tile * BUF_BYTES, BUF_BYTES, BAR(B_DZFULL + zb));
      }
    };
    if constexpr (mode_has_shadow(MODE)) {
      produce_tiles_h<MODE, NKB1, KB0>(d, ntiles, pw, lane, [&](const ChunkRegs& c, uint32_t tcount, int kb) {
        if (kb == KB0) fetch_dz(tcount);
        const int s = it % NSTAGE;
        mbar_wait(BAR(B_XEMPTY + s), ((it / NSTAGE) & 1) ^ 1);
        store_chunk_h(ring + s * KB_BYTES, pw, lane, c);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_XFULL + s));
        ++it;
      });
    } else {
      TileIdx idx;
      uint32_t tcount = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
        const int64_t row0 = tile * TILE_M;
        fetch_dz(tcount);
        for (int kb = KB0; kb < NKB1; ++kb, ++it) {
          const int s = it % NSTAGE;
          mbar_wait(BAR(B_XEMPTY + s), ((it / NSTAGE) & 1) ^ 1);
          produce_chunk<MODE, P>(d, row0, kb, ring + s * KB_BYTES, pw, lane, idx);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_XFULL + s));
        }
      }
    }
  } else {
    // ============================================================ epilogue
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int rloc = warp * 32 + lane;
    uint8_t* mystg = stg + warp * WSTG_BYTES;
    const int orow = lane >> 3, oseg = lane & 7;  // write-out mapping: 4 rows x 8 x 16 B per warp instruction
    if (C::DX) {
      uint32_t blk = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t wrow0 = tile * TILE_M + warp * 32;
        if (resid && (MODE == FVGN_MLP_EDGE || MODE == FVGN_MLP_NODE) && wrow0 + lane < d.rows) {
          // the residual-gradient rows are read at the END of this tile (last two blocks): pull them into L2 now
          if (resid_h) {
            prefetch_l2(resid_h + (size_t)(wrow0 + lane) * 256);
            prefetch_l2(resid_h + (size_t)(wrow0 + lane) * 256 + 128);
          } else {
            const float* rp = d.d_out + (size_t)(wrow0 + lane) * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) prefetch_l2(rp + 32 * k);
          }
        }
        for (int j = KB0; j < NKB1; ++j, ++blk) {
          const int ab = blk & 1;
          const uint32_t tacc = tmem + lane_base + WACC + 64 * ab;
          const int col0 = 64 * j;  // first input column of this 64-column block
          // destination of the block (uniform: boundaries are multiples of 64)
          float* dst;
          int ld, colo;
          bool add_res = false, to_bf16 = false, in1_blk = false;
          if (MODE == FVGN_MLP_EDGE) {
            if (col0 < 256) { dst = d.d_in0; ld = 256; colo = col0; to_bf16 = d.d_in0h != nullptr; }
            else { dst = d.d_in1; ld = 128; colo = col0 - 256; add_res = resid; to_bf16 = out16; in1_blk = true; }
          } else if (MODE == FVGN_MLP_NODE) {
            if (col0 < 64) { dst = d.d_in0; ld = 64; colo = col0; to_bf16 = d.d_in0h != nullptr; }
            else { dst = d.d_in1; ld = 128; colo = col0 - 64; add_res = resid; to_bf16 = out16; in1_blk = true; }
          } else {
            dst = d.d_in0; ld = 128; colo = col0; to_bf16 = d.d_in0h != nullptr;   // DEC: d_x, optionally as 16-bit rows
          }
          // residual gradient of this block: both 32-column halves are requested before waiting for the accumulator
          float4 g[8], g1[8];
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const int64_t row = wrow0 + ps * 4 + orow;
            g[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
            g1[ps] = g[ps];
            if (add_res && row < d.rows) {
              // (16-bit residual rows are kept as raw bits and unpacked where they are used, so that these loads stay in
              //  flight behind the accumulator wait)
              if (to_bf16) {   // 16-bit destination: this lane's 8 columns of the 64-column block; residual from either stream
                if (resid_h) {
                  const uint4 w = __ldg(reinterpret_cast<const uint4*>(resid_h + (size_t)row * 256 + colo * 2 + oseg * 16));
                  g[ps] = make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
                } else {
                  const float* gp = d.d_out + (size_t)row * 128 + colo + oseg * 8;
                  g[ps] = __ldg(reinterpret_cast<const float4*>(gp));
                  g1[ps] = __ldg(reinterpret_cast<const float4*>(gp + 4));
                }
              } else if (resid_h) {   // fp32 destination, 16-bit residual: columns colo + oseg*4 .. +3 and + 32
                const uint2 w0 = __ldg(reinterpret_cast<const uint2*>(resid_h + (size_t)row * 256 + (colo + oseg * 4) * 2));
                const uint2 w1 = __ldg(reinterpret_cast<const uint2*>(resid_h + (size_t)row * 256 + (colo + 32 + oseg * 4) * 2));
                g[ps].x = __uint_as_float(w0.x); g[ps].y = __uint_as_float(w0.y);
                g1[ps].x = __uint_as_float(w1.x); g1[ps].y = __uint_as_float(w1.y);
              } else {
                const float* gp = d.d_out + (size_t)row * 128 + colo + oseg * 4;
                g[ps] = __ldg(reinterpret_cast<const float4*>(gp));
                g1[ps] = __ldg(reinterpret_cast<const float4*>(gp + 32));
              }
            }
          }
          mbar_wait(BAR(B_AFULL + ab), (blk >> 1) & 1);
          tc_fence_after();
          if (to_bf16) {
            // 64 columns -> 128 B of bf16 per row: one staging tile, one coalesced pass
            uint32_t r[32];
            float rs = 1.f;  // NODE: D^-1 of the transposed scatter_mean applied to this thread's row of d_a2
            if (MODE == FVGN_MLP_NODE && !in1_blk && d.d_in0_row_ptr != nullptr && wrow0 + lane < d.rows)
              rs = 1.f / (float)max(__ldg(d.d_in0_row_ptr + wrow0 + lane + 1) - __ldg(d.d_in0_row_ptr + wrow0 + lane), 1);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              tmem_ld32(tacc + 32 * half, r);
              tmem_wait_ld();
              if (MODE == FVGN_MLP_NODE && !in1_blk) {
#pragma unroll
                for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * rs);
              }
#pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<uint4*>(wstg_at(mystg, lane, half * 4 + k)) =
                    make_uint4(pack16<P>(__uint_as_float(r[8 * k]), __uint_as_float(r[8 * k + 1])),
                               pack16<P>(__uint_as_float(r[8 * k + 2]), __uint_as_float(r[8 * k + 3])),
                               pack16<P>(__uint_as_float(r[8 * k + 4]), __uint_as_float(r[8 * k + 5])),
                               pack16<P>(__uint_as_float(r[8 * k + 6]), __uint_as_float(r[8 * k + 7])));
            }
            tc_fence_before();
            mbar_arrive(BAR(B_AFREE + ab));
            __syncwarp();
            uint8_t* dh = reinterpret_cast<uint8_t*>(in1_blk ? d.d_in1h : d.d_in0h);
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
              const int rr = ps * 4 + orow;
              const int64_t row = wrow0 + rr;
              if (row < d.rows) {
                uint4 v = *reinterpret_cast<const uint4*>(wstg_at(mystg, rr, oseg));
                if (add_res && resid_h) {   // d_e / d_x = dX + residual gradient: both 16-bit rows, packed adds (no conversions)
                  const float4 ga = g[ps];
                  v = make_uint4(P::add2(v.x, __float_as_uint(ga.x)), P::add2(v.y, __float_as_uint(ga.y)),
                                 P::add2(v.z, __float_as_uint(ga.z)), P::add2(v.w, __float_as_uint(ga.w)));
                } else if (add_res) {       // fp32 residual gradient
                  const float4 ga = g[ps], gb = g1[ps];
                  v = make_uint4(pack16<P>(P::lo(v.x) + ga.x, P::hi(v.x) + ga.y), pack16<P>(P::lo(v.y) + ga.z, P::hi(v.y) + ga.w),
                                 pack16<P>(P::lo(v.z) + gb.x, P::hi(v.z) + gb.y), pack16<P>(P::lo(v.w) + gb.z, P::hi(v.w) + gb.w));
                }
                *reinterpret_cast<uint4*>(dh + (size_t)row * (ld * 2) + colo * 2 + oseg * 16) = v;
              }
            }
            __syncwarp();
          } else {
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
              uint32_t r[32];
              tmem_ld32(tacc + 32 * half, r);
              if (half == 1) {
#pragma unroll
                for (int ps = 0; ps < 8; ++ps) g[ps] = g1[ps];
              }
              tmem_wait_ld();
#pragma unroll
              for (int k = 0; k < 8; ++k)
                *reinterpret_cast<float4*>(wstg_at(mystg, lane, k)) =
                    make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]),
                                __uint_as_float(r[4 * k + 3]));
              if (half == 1) {
                tc_fence_before();
                mbar_arrive(BAR(B_AFREE + ab));
              }
              __syncwarp();
#pragma unroll
              for (int ps = 0; ps < 8; ++ps) {
                const int rr = ps * 4 + orow;
                const int64_t row = wrow0 + rr;
                if (row < d.rows) {
                  float4 v = *reinterpret_cast<const float4*>(wstg_at(mystg, rr, oseg));
                  float4 ga = g[ps];
                  if (add_res && resid_h) {
                    const uint32_t w0 = __float_as_uint(ga.x), w1 = __float_as_uint(ga.y);
                    ga = make_float4(P::lo(w0), P::hi(w0), P::lo(w1), P::hi(w1));
                  }
                  v = make_float4(v.x + ga.x, v.y + ga.y, v.z + ga.z, v.w + ga.w);
                  *reinterpret_cast<float4*>(dst + (size_t)row * ld + colo + 32 * half + oseg * 4) = v;
                }
              }
              __syncwarp();
            }
          }
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
        for (int j = 0; j < 16; ++j)   // columns below KB0 * 64 belong to the node-level kernel: zero here
          if (c0 + j < C::K1) Pw1[(size_t)o * C::K1 + c0 + j] = (c0 < KB0 * 64) ? 0.f : __uint_as_float(r[j]);
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

// out[i] = unscale * sum over the CTAs' partial buffers (fixed order: deterministic).  unscale (nullable device scalar) is
// 1 / S of the power-of-two gradient pre-scaling of the fp16 mode (ops.GradScaleFn): the multiplication is exact.
__global__ void __launch_bounds__(256) tc_partial_reduce_kernel(const float* __restrict__ partials, int n_partials, int64_t pc,
                                                                const float* __restrict__ unscale, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= pc) return;
  float s = 0.f;
  for (int g = 0; g < n_partials; ++g) s += partials[(size_t)g * pc + i];
  out[i] = unscale ? s * __ldg(unscale) : s;
}

template <int MODE> constexpr int smem_a() { return 4 * KB_BYTES + 4 * BUF_BYTES + 3 * 512 + (A_NEW / 4) * 2048 + 256; }
template <int MODE> constexpr int smem_b() {
  return nkb1(BCfg<MODE>::K1P) * KB_BYTES + 2 * BUF_BYTES + 3 * KB_BYTES + STG_BYTES + 256;
}

int num_sms() { return fvgn_num_sms(); }

}  // namespace
template <class P> int launch_tc_bwd_node(const fvgn_mlp_desc& d, void* stream);   // mlp_tc_bwd_node.cu
namespace {

template <int MODE, class P>
int launch_tc_bwd(const fvgn_mlp_desc& d, void* stream) {
  using C = BCfg<MODE>;
  auto ka = mlp_tc_bwd_a_kernel<MODE, P>;
  constexpr int KBN = (MODE == FVGN_MLP_EDGE) ? 4 : 0;   // EDGE with the node-level layer-1 path: e columns only
  constexpr bool GRAD = MODE == FVGN_MLP_EDGE || MODE == FVGN_MLP_NODE;   // modes with a residual gradient path
  const bool node_path = MODE == FVGN_MLP_EDGE && d.d_aggh != nullptr;
  // pure 16-bit: 16-bit destination and a 16-bit (or no) upstream gradient; pure fp32 likewise; anything else is mixed
  const bool out16 = d.d_in1h != nullptr;
  const int gs = !GRAD ? 0 : (out16 && (d.d_outh || !d.d_out)) ? 1 : (!out16 && (d.d_out || !d.d_outh)) ? 0 : 2;
  void (*kb)(const fvgn_mlp_desc) = mlp_tc_bwd_b_kernel<MODE, P, 0, 0>;
  void (*kbn)(const fvgn_mlp_desc) = mlp_tc_bwd_b_kernel<MODE, P, KBN, 0>;
  if (GRAD && gs == 1) { kb = mlp_tc_bwd_b_kernel<MODE, P, 0, GRAD ? 1 : 0>; kbn = mlp_tc_bwd_b_kernel<MODE, P, KBN, GRAD ? 1 : 0>; }
  if (GRAD && gs == 2) { kb = mlp_tc_bwd_b_kernel<MODE, P, 0, GRAD ? 2 : 0>; kbn = mlp_tc_bwd_b_kernel<MODE, P, KBN, GRAD ? 2 : 0>; }
  static bool attr_set[FVGN_MAX_DEV][3] = {{false}};  // the attribute is per device (and per kernel variant)
  const int dev = fvgn_cur_device();
  if (!attr_set[dev][gs]) {
    if (cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a<MODE>()) != cudaSuccess) return FVGN_ERR_LAUNCH;
    if (cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b<MODE>()) != cudaSuccess) return FVGN_ERR_LAUNCH;
    if (cudaFuncSetAttribute(kbn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b<MODE>()) != cudaSuccess) return FVGN_ERR_LAUNCH;
    attr_set[dev][gs] = true;
  }
  const unsigned grid = (unsigned)d.n_partials;
  ka<<<grid, A_THREADS, smem_a<MODE>(), (cudaStream_t)stream>>>(d);
  FVGN_CHECK_LAUNCH();
  if (node_path) kbn<<<grid, NTHREADS, smem_b<MODE>(), (cudaStream_t)stream>>>(d);
  else kb<<<grid, NTHREADS, smem_b<MODE>(), (cudaStream_t)stream>>>(d);
  FVGN_CHECK_LAUNCH();
  const int64_t pc = pcount(C::K1, C::NOUT, C::LN);
  tc_partial_reduce_kernel<<<(unsigned)((pc + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d.partials, d.n_partials, pc, d.grad_unscale,
                                                                                       d.d_params);
  FVGN_CHECK_LAUNCH();
  if (node_path) return launch_tc_bwd_node<P>(d, stream);   // d(agg) and dW1[:, 0:256] at node level (overwrites those columns)
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

#ifdef FVGN_TIMING
extern "C" int fvgn_debug_profile_a(unsigned long long* out16, int reset) {
  if (cudaMemcpyFromSymbol(out16, g_prof_a, sizeof(unsigned long long) * 16) != cudaSuccess) return FVGN_ERR_LAUNCH;
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_prof_a, z, sizeof(z));
  }
  return FVGN_OK;
}
#endif

int fvgn_mlp_backward_simt(const fvgn_mlp_desc* d, void* stream);

int fvgn_mlp_backward_tc(const fvgn_mlp_desc* d, void* stream) {
  if (!d->w_bf16 || !d->workspace || !d->z1_img) return FVGN_ERR_NULL;
  if ((((uintptr_t)d->w_bf16) & 15) != 0 || (((uintptr_t)d->workspace) & 1023) != 0 || (((uintptr_t)d->z1_img) & 1023) != 0)
    return FVGN_ERR_ALIGN;
  if (d->n_partials != fvgn_mlp_tc_partials(d->mode, d->rows)) return FVGN_ERR_SHAPE;
  if (d->rows == 0) {
    const int64_t pc = fvgn_mlp_param_count(d->mode);
    if (cudaMemsetAsync(d->d_params, 0, (size_t)pc * sizeof(float), (cudaStream_t)stream) != cudaSuccess) return FVGN_ERR_LAUNCH;
    return FVGN_OK;
  }
  const bool f16 = d->precision == FVGN_PREC_F16;
#define FVGN_TC_BWD_CASE(M) case M: return f16 ? launch_tc_bwd<M, PF16>(*d, stream) : launch_tc_bwd<M, PBF16>(*d, stream);
  switch (d->mode) {
    FVGN_TC_BWD_CASE(FVGN_MLP_EDGE)
    FVGN_TC_BWD_CASE(FVGN_MLP_NODE)
    FVGN_TC_BWD_CASE(FVGN_MLP_ENC_NODE)
    FVGN_TC_BWD_CASE(FVGN_MLP_ENC_EDGE)
    FVGN_TC_BWD_CASE(FVGN_MLP_DEC)
  }
#undef FVGN_TC_BWD_CASE
  return FVGN_ERR_UNSUPPORTED;
}
