// Fused 3-layer MLP blocks on 5th-gen tensor cores (FVGN_PREC_BF16 / FVGN_PREC_F16): tcgen05.mma kind::f16, 16-bit
// operands (bf16, or IEEE half = the 11-bit significand of TF32), fp32 accumulators in TMEM, fp32 bias / GELU /
// LayerNorm / residual, fp32 residual streams in HBM.
//
// Persistent kernel, one CTA per SM, tile = 128 rows (UMMA M=128, N=128, K=16):
//   warps 4-7  producers : gather the input rows (edge endpoints / concatenations / relative edge features),
//                          convert to bf16 and write 64-wide K chunks into a shared-memory ring in the canonical
//                          K-major SWIZZLE_128B UMMA layout
//   warp  8    MMA issuer: one elected thread issues tcgen05.mma; layer 1 reads A from the ring (SS), layers 2/3
//                          read A straight from TMEM (TS) where the epilogue left the GELU output as bf16
//   warps 0-3  epilogue  : tcgen05.ld accumulator -> +bias -> GELU -> bf16 -> tcgen05.st (next layer's A operand);
//                          last layer: LayerNorm (one thread owns one row: no cross-lane reduction) + residual,
//                          transposed through a small smem stage for coalesced stores
// All three weight matrices stay resident in shared memory as a pre-swizzled bf16 image (<= 160 KB), loaded once
// per CTA with cp.async.bulk; HBM traffic is the activation rows only.
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int NTHREADS = 416;  // 8 epilogue warps (two tile pipelines) + 4 producer warps + 1 MMA warp

template <int MODE> struct TCfg;
template <> struct TCfg<FVGN_MLP_EDGE> { static constexpr int K1P = 384, NOUT = 128, NSTAGE = 2; static constexpr bool LN = true; };
template <> struct TCfg<FVGN_MLP_NODE> { static constexpr int K1P = 192, NOUT = 128, NSTAGE = 4; static constexpr bool LN = true; };
template <> struct TCfg<FVGN_MLP_ENC_NODE> { static constexpr int K1P = 16, NOUT = 128, NSTAGE = 4; static constexpr bool LN = true; };
template <> struct TCfg<FVGN_MLP_ENC_EDGE> { static constexpr int K1P = 16, NOUT = 128, NSTAGE = 4; static constexpr bool LN = true; };
template <> struct TCfg<FVGN_MLP_DEC> { static constexpr int K1P = 128, NOUT = 3, NSTAGE = 4; static constexpr bool LN = false; };

constexpr int STG_BYTES = 8 * WSTG_BYTES;  // one wide staging tile per epilogue warp

template <int MODE> constexpr int smem_bytes() {
  return image_bytes(TCfg<MODE>::K1P) + TCfg<MODE>::NSTAGE * KB_BYTES + STG_BYTES + 5 * 512 + 256;
}

// ------------------------------------------------------------------------------------------ weight image
// image = [W1 K-blocks][W2: 2 K-blocks][W3: 2 K-blocks], each K-block 128 rows x 64 k, bf16, pre-swizzled
template <class P>
__global__ void pack_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ w3,
                                    int k1, int k1p, int nout, uint8_t* __restrict__ img) {
  const int nkb = nkb1(k1p);
  const int total = (nkb + 4) * 128 * 8;  // 16-byte chunks
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int kb = idx / (128 * 8), row = (idx / 8) % 128, chunk = idx % 8;
    uint32_t packed[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int kk = chunk * 8 + j * 2 + h;
        float x = 0.f;
        if (kb < nkb) {
          const int k = kb * 64 + kk;
          if (k < k1) x = w1[(size_t)row * k1 + k];
        } else if (kb < nkb + 2) {
          x = w2[(size_t)row * 128 + (kb - nkb) * 64 + kk];
        } else {
          if (row < nout) x = w3[(size_t)row * 128 + (kb - nkb - 2) * 64 + kk];
        }
        v[h] = x;
      }
      packed[j] = pack16<P>(v[0], v[1]);
    }
    *reinterpret_cast<uint4*>(img + (size_t)kb * KB_BYTES + sw128_off(row, chunk)) =
        make_uint4(packed[0], packed[1], packed[2], packed[3]);
  }
}

// ------------------------------------------------------------------------------------------ forward kernel
// Two tiles are in flight per CTA (pipelines P0 / P1: one epilogue warpgroup and two 128-column TMEM accumulators each).
// The bf16 A operand of layers 2/3 is written IN PLACE over the lower 64 columns of the accumulator it was computed
// from (column pair (2j,2j+1) -> column j, always behind the read front), so per tile  L1 -> T1, L2: T1[0:64] -> T2,
// L3: T2[0:64] -> T1, and the NEXT tile of the pipeline runs its layer 1 into T2 while the epilogue still drains T1.
// MMA issue order per tile pair:  L2(P0) L2(P1) L3(P0) L1'(P0) L3(P1) L1'(P1)   (L1' = layer 1 of the following pair),
// so loads, tensor work and the three epilogues of different tiles overlap.
//
// Memory side: layer-1 operands come from bf16 shadows (8 x 16 B per thread per chunk, held in registers two chunks
// ahead of the ring); every result leaves through a 4 KB swizzled staging tile per warp as 128-B row segments
// (Z1 image: one 4 KB bulk store per 64 columns; fp32 residual stream and bf16 shadows: 4 rows per warp instruction).
#ifdef FVGN_TIMING
// debug build: cycle breakdown of the forward epilogue chain of pipeline 0 (thread 0 of CTA 0), summed over its tiles
__device__ unsigned long long g_prof_f[16];
#define PROF_F(i)                                               \
  do {                                                          \
    if (blockIdx.x == 0 && tid == 0) {                          \
      const long long now_ = clock64();                         \
      g_prof_f[i] += (unsigned long long)(now_ - tprev_);       \
      tprev_ = now_;                                            \
    }                                                           \
  } while (0)
#else
#define PROF_F(i) do { } while (0)
#endif

template <int MODE, class P>
__global__ void __launch_bounds__(NTHREADS, 1) mlp_tc_fwd_kernel(const fvgn_mlp_desc d) {
  using C = TCfg<MODE>;
  constexpr uint32_t IDESC = make_idesc(P::FMT, 128, 0, 0);
  constexpr int NKB1 = nkb1(C::K1P);
  constexpr int LASTK = (C::K1P - 64 * (NKB1 - 1)) / 16;  // MMAs (K=16) in the last K-block of layer 1
  constexpr int NSTAGE = C::NSTAGE;
  FVGN_DYN_SMEM(smem);
  uint8_t* w_img = smem;
  uint8_t* ring = w_img + image_bytes(C::K1P);
  uint8_t* stg = ring + NSTAGE * KB_BYTES;
  float* sb1 = reinterpret_cast<float*>(stg + STG_BYTES);
  float* sb2 = sb1 + 128;
  float* sb3 = sb2 + 128;
  float* sg = sb3 + 128;
  float* sbeta = sg + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbeta + 128);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  // barrier map: weights | ring full | ring empty | per pipeline: layer-1 done, layer-2/3 done, A operand ready
  constexpr int B_W = 0, B_FULL = 1, B_EMPTY = 1 + NSTAGE, B_L1 = 1 + 2 * NSTAGE, B_L23 = B_L1 + 2, B_A = B_L1 + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (d.rows + TILE_M - 1) / TILE_M;
  // tiles of this CTA: blockIdx.x + i * gridDim.x, i = 0 .. ntl-1; tile i belongs to pipeline i & 1
  const int64_t ntl = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (tid == 0) {
    mbar_init(BAR(B_W), 1);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(BAR(B_FULL + s), 4);   // one elected lane per producer warp
      mbar_init(BAR(B_EMPTY + s), 1);  // tcgen05.commit
    }
    for (int p = 0; p < 2; ++p) {
      mbar_init(BAR(B_L1 + p), 1);
      mbar_init(BAR(B_L23 + p), 1);
      mbar_init(BAR(B_A + p), 128);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 128; i += NTHREADS) {
    sb1[i] = d.b1[i];
    sb2[i] = d.b2[i];
    sb3[i] = (i < C::NOUT) ? d.b3[i] : 0.f;
    sg[i] = C::LN ? d.ln_g[i] : 1.f;
    sbeta[i] = C::LN ? d.ln_b[i] : 0.f;
  }
  if (warp == 12) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 12) {
    // =============================================================== MMA issuer (+ weight loader)
    if (lane == 0) {
      constexpr uint32_t img_bytes = (uint32_t)image_bytes(C::K1P);
      mbar_expect_tx(BAR(B_W), img_bytes);
      for (uint32_t off = 0; off < img_bytes; off += KB_BYTES)
        bulk_g2s(smem_u32(w_img + off), reinterpret_cast<const uint8_t*>(d.w_bf16) + off, KB_BYTES, BAR(B_W));
      mbar_wait(BAR(B_W), 0);
      const uint32_t w1s = smem_u32(w_img), w2s = w1s + NKB1 * KB_BYTES, w3s = w2s + 2 * KB_BYTES;
      uint32_t it = 0;
      uint32_t pa[2] = {0, 0}, pl[2] = {0, 0};
      // accumulators of tile i: T1 = first target (layer 1, layer 3), T2 = layer 2 target; they swap every tile
      auto T1 = [&](int64_t i) { return tmem + 256 * (uint32_t)(i & 1) + 128 * (uint32_t)((i >> 1) & 1); };
      auto T2 = [&](int64_t i) { return tmem + 256 * (uint32_t)(i & 1) + 128 * (uint32_t)(((i >> 1) & 1) ^ 1); };
      // layers 2/3: A = bf16 image in the lower 64 columns of `src`, D = `dst`
      auto issue23 = [&](uint32_t dst, uint32_t src, uint32_t ws, int p) {
        pa[p] ^= 1;
        tc_fence_after();
        for (int k = 0; k < 8; ++k)
          umma_ts(dst, src + 8 * k, make_desc_k128(ws + (k >> 2) * KB_BYTES) + 2 * (k & 3), IDESC, k != 0);
        umma_commit(BAR(B_L23 + p));
        pl[p] ^= 1;
      };
      // EVENT-DRIVEN issue.  Per pipeline p (tiles p, p+2, ..) the program order is  L1 L2 L3 | L1 L2 L3 | ..; the layer-1
      // chunks of ALL tiles are consumed from the ring in tile order.  Instead of blocking on one barrier at a time (the
      // layer-2/3 MMAs of one pipeline then queue behind the ring waits of the other pipeline's layer 1: 23 % of an EDGE
      // tile's timeline in the round-1 profile), the issuing thread polls: whichever of {next layer-2/3 step of P0, of P1,
      // next ring chunk of the layer-1 stream} has its operands ready is issued.
      int64_t l23_tile[2] = {0, 1};   // tile whose layer 2 / 3 is next on pipeline p
      int l23_step[2] = {0, 0};       // 0: layer 2 next, 1: layer 3 next
      int64_t l1_tile = 0;            // layer-1 stream: tile and K-block to issue next
      int l1_kb = 0;
      bool l1_open = false;           // accumulator of l1_tile known to be free (layer 3 of tile l1_tile - 2 retired)
      while (l1_tile < ntl || l23_tile[0] < ntl || l23_tile[1] < ntl) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int64_t t = l23_tile[p];
          // layer 2/3 of tile t needs its layer 1 issued (program order) and the epilogue's bf16 A operand in TMEM
          if (t < ntl && (t < l1_tile) && mbar_test(BAR(B_A + p), pa[p])) {
            if (l23_step[p] == 0) {
              issue23(T2(t), T1(t), w2s, p);
              l23_step[p] = 1;
            } else {
              issue23(T1(t), T2(t), w3s, p);
              l23_step[p] = 0;
              l23_tile[p] = t + 2;
            }
          }
        }
        if (l1_tile < ntl) {
          const int p = (int)(l1_tile & 1);
          if (!l1_open) {
            // layer 1 of tile t targets T2 of tile t-2, whose lower half is the A operand of that tile's layer 3: it must
            // have been issued (l23_tile[p] moved past t-2) and retired
            if (l1_tile < 2) l1_open = true;
            else if (l23_tile[p] >= l1_tile && l23_step[p] == 0 && mbar_test(BAR(B_L23 + p), pl[p] ^ 1)) {
              tc_fence_after();
              l1_open = true;
            }
          }
          if (l1_open) {
            const int s = it % NSTAGE;
            if (mbar_test(BAR(B_FULL + s), (it / NSTAGE) & 1)) {
              tc_fence_after();
              const uint32_t acc = T1(l1_tile);
              const uint64_t ad = make_desc_k128(smem_u32(ring + s * KB_BYTES));
              const uint64_t bd = make_desc_k128(w1s + l1_kb * KB_BYTES);
              const int nk = (l1_kb == NKB1 - 1) ? LASTK : 4;
              for (int k = 0; k < nk; ++k) umma_ss(acc, ad + 2 * k, bd + 2 * k, IDESC, (l1_kb | k) != 0);
              umma_commit(BAR(B_EMPTY + s));
              ++it;
              if (++l1_kb == NKB1) {
                umma_commit(BAR(B_L1 + p));
                l1_kb = 0;
                ++l1_tile;
                l1_open = false;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // =============================================================== producers (4 warps)
    const int pw = warp - 8;
    uint32_t it = 0;
    if constexpr (mode_has_shadow(MODE)) {
      // chunks are held in registers (2-3 deep) ahead of the ring
      produce_tiles_h<MODE, NKB1>(d, ntiles, pw, lane, [&](const ChunkRegs& c, uint32_t, int) {
        const int s = it % NSTAGE;
        mbar_wait(BAR(B_EMPTY + s), ((it / NSTAGE) & 1) ^ 1);
        store_chunk_h(ring + s * KB_BYTES, pw, lane, c);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_FULL + s));
        ++it;
      });
    } else {
      TileIdx idx;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * TILE_M;
        for (int kb = 0; kb < NKB1; ++kb, ++it) {
          const int s = it % NSTAGE;
          mbar_wait(BAR(B_EMPTY + s), ((it / NSTAGE) & 1) ^ 1);
          produce_chunk<MODE, P>(d, row0, kb, ring + s * KB_BYTES, pw, lane, idx);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_FULL + s));
        }
      }
    }
  } else {
    // =============================================================== epilogue: pipeline p = warp / 4, thread <-> row
    const int p = warp >> 2, q = warp & 3;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int rloc = q * 32 + lane;
    uint8_t* mystg = stg + warp * WSTG_BYTES;
    const int orow = lane >> 3, oseg = lane & 7;  // write-out mapping: 4 rows x 8 x 16 B per warp instruction
    // residual row: the fp32 stream in1, or (FVGN_MLP_RESIDUAL_FROM_SHADOW: 16-bit latent streams) the 16-bit shadow in1h the
    // layer-1 operand was read from -- then the fp32 row is neither read nor (out_res == NULL) written
    const bool res16 = (MODE == FVGN_MLP_EDGE || MODE == FVGN_MLP_NODE) && (d.flags & FVGN_MLP_RESIDUAL_FROM_SHADOW);
    const float* resid = (MODE == FVGN_MLP_EDGE || MODE == FVGN_MLP_NODE) && !res16 ? d.in1 : nullptr;
    const uint16_t* resid_h = res16 ? reinterpret_cast<const uint16_t*>(d.in1h) : nullptr;
    const bool do_res = (resid && d.out_res) || (resid_h && (d.out_res || d.out_resh));
    bool stg_busy = false;  // a bulk store may still be reading the staging tile
    uint32_t ph1 = 0, ph23 = 0;
#ifdef FVGN_TIMING
    long long tprev_ = clock64();
#endif
    for (int64_t i = p; i < ntl; i += 2) {
      const int64_t tile = blockIdx.x + i * gridDim.x;
      const int64_t row0 = tile * TILE_M;
      const int64_t wrow0 = row0 + q * 32;  // first row of this warp
      const uint32_t t1 = tmem + lane_base + 256 * (uint32_t)p + 128 * (uint32_t)((i >> 1) & 1);
      const uint32_t t2 = tmem + lane_base + 256 * (uint32_t)p + 128 * (uint32_t)(((i >> 1) & 1) ^ 1);
      if (do_res && wrow0 + lane < d.rows) {
        // the fp32 residual row is first touched ~3 epilogue passes from now: pull it into L2 meanwhile
        if (resid_h) {
          const uint16_t* rp = resid_h + (size_t)(wrow0 + lane) * 128;
          prefetch_l2(rp);
          prefetch_l2(rp + 64);
        } else {
          const float* rp = resid + (size_t)(wrow0 + lane) * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) prefetch_l2(rp + 32 * k);
        }
      }
      // ---- hidden layers: +bias, GELU, bf16 written in place over the accumulator's lower 64 columns
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        if (layer == 0) {
          mbar_wait(BAR(B_L1 + p), ph1);
          ph1 ^= 1;
          PROF_F(0);  // wait layer 1 (producers + MMA)
        } else {
          mbar_wait(BAR(B_L23 + p), ph23);
          ph23 ^= 1;
          PROF_F(2);  // wait layer 2
        }
        tc_fence_after();
        const uint32_t acc = layer == 0 ? t1 : t2;
        const float* bias = layer == 0 ? sb1 : sb2;
        // layer 1 only: Z1 is rounded to bf16 (what the backward reads back) and stored as a pre-swizzled tile image;
        // this warp's 32 rows x 64 columns are 4 KB contiguous in the image (same XOR pattern as the staging tile)
        uint8_t* zimg = (layer == 0 && d.z1_img)
                            ? reinterpret_cast<uint8_t*>(d.z1_img) + (size_t)tile * (2 * KB_BYTES) + q * WSTG_BYTES : nullptr;
        for_each_chunk16(acc, [&](int c0, uint32_t (&r)[16]) {
          uint32_t w[8];
#if FVGN_F32X2
          float2 z[8];   // z[j] = columns (2j, 2j+1)
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float4 b = *reinterpret_cast<const float4*>(bias + c0 + 2 * j);
            z[j] = __fadd2_rn(f2u(r[2 * j], r[2 * j + 1]), make_float2(b.x, b.y));
            z[j + 1] = __fadd2_rn(f2u(r[2 * j + 2], r[2 * j + 3]), make_float2(b.z, b.w));
          }
#else
          float z[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(bias + c0 + j);
            z[j] = __uint_as_float(r[j]) + b.x;
            z[j + 1] = __uint_as_float(r[j + 1]) + b.y;
            z[j + 2] = __uint_as_float(r[j + 2]) + b.z;
            z[j + 3] = __uint_as_float(r[j + 3]) + b.w;
          }
#endif
          if (layer == 0) {
            uint32_t zw[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#if FVGN_F32X2
              zw[j] = pack16<P>(z[j]);
              z[j] = unpack16<P>(zw[j]);
#else
              zw[j] = pack16<P>(z[2 * j], z[2 * j + 1]);
              z[2 * j] = P::lo(zw[j]);
              z[2 * j + 1] = P::hi(zw[j]);
#endif
            }
            if (zimg) {
              const int ch = (c0 & 63) >> 3;
              if (ch == 0 && stg_busy) {  // previous bulk store must have finished reading the staging tile
                if (lane == 0) bulk_wait_read0();
                __syncwarp();
                stg_busy = false;
              }
              *reinterpret_cast<uint4*>(wstg_at(mystg, lane, ch)) = make_uint4(zw[0], zw[1], zw[2], zw[3]);
              *reinterpret_cast<uint4*>(wstg_at(mystg, lane, ch + 1)) = make_uint4(zw[4], zw[5], zw[6], zw[7]);
              if (ch == 6) {  // 64 columns staged: one 4 KB bulk store
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) bulk_s2g(zimg + (c0 >> 6) * KB_BYTES, smem_u32(mystg), WSTG_BYTES);
                stg_busy = true;
              }
            }
          }
#pragma unroll
#if FVGN_F32X2
          for (int j = 0; j < 8; ++j) w[j] = pack16<P>(gelu_tanh2(z[j]));
#else
          for (int j = 0; j < 8; ++j) w[j] = pack16<P>(gelu_tanh(z[2 * j]), gelu_tanh(z[2 * j + 1]));
#endif
          tmem_st8(acc + c0 / 2, w);  // in place: always behind the columns still to be read (and the one in flight)
        });
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(BAR(B_A + p));
        if (layer == 0) PROF_F(1); else PROF_F(3);  // epilogue of layer 1 / layer 2
      }
      // ---- output layer (accumulated into T1)
      mbar_wait(BAR(B_L23 + p), ph23);
      ph23 ^= 1;
      PROF_F(4);  // wait layer 3
      tc_fence_after();
      const uint32_t acc = t1;
      if (C::LN) {
        // LayerNorm statistics in one pass over the accumulator (sum / sum of squares in fp32)
        float sum = 0.f, sq = 0.f;
#if FVGN_F32X2
        float2 sum2 = splat2(0.f), sq2 = splat2(0.f);
#endif
        for_each_chunk16(acc, [&](int c0, uint32_t (&r)[16]) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(sb3 + c0 + j);
#if FVGN_F32X2
            const float2 ya = __fadd2_rn(f2u(r[j], r[j + 1]), make_float2(b.x, b.y));
            const float2 yb = __fadd2_rn(f2u(r[j + 2], r[j + 3]), make_float2(b.z, b.w));
            sum2 = __fadd2_rn(sum2, __fadd2_rn(ya, yb));
            sq2 = __ffma2_rn(ya, ya, __ffma2_rn(yb, yb, sq2));
            continue;
#endif
            const float y0 = __uint_as_float(r[j]) + b.x, y1 = __uint_as_float(r[j + 1]) + b.y;
            const float y2 = __uint_as_float(r[j + 2]) + b.z, y3 = __uint_as_float(r[j + 3]) + b.w;
            sum += (y0 + y1) + (y2 + y3);
            sq = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, sq))));
          }
        });
#if FVGN_F32X2
        sum = sum2.x + sum2.y;
        sq = sq2.x + sq2.y;
#endif
        const float mean = sum * (1.0f / 128.0f);
        const float rstd = rsqrtf(fmaxf(sq * (1.0f / 128.0f) - mean * mean, 0.f) + 1e-5f);
        PROF_F(5);  // LayerNorm statistics
        if (stg_busy) {
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
          stg_busy = false;
        }
        uint16_t* outh = reinterpret_cast<uint16_t*>(d.outh);
        uint16_t* out_resh = reinterpret_cast<uint16_t*>(d.out_resh);
        // The normalisation is applied AFTER the transposition through the staging tile: in the write-out mapping a lane owns 4
        // fixed columns of 8 rows, so gamma / beta / b3 are three 16-byte reads per 32-column group (instead of 24 in the
        // thread-per-row mapping) and the accumulator chunk leaves its registers at once; the statistics of the 8 rows a lane
        // serves are fetched once per tile.  Same operations in the same order per element: results unchanged.
        float nmr[8], rsr[8];
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
          nmr[ps] = -__shfl_sync(0xffffffffu, mean, ps * 4 + orow);
          rsr[ps] = __shfl_sync(0xffffffffu, rstd, ps * 4 + orow);
        }
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(acc + c0, r);
          // residual: 8 passes x (4 rows x 128 B); issued before the normalisation math
          float4 xr[8];
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const int64_t row = wrow0 + ps * 4 + orow;
            xr[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (do_res && row < d.rows) {
              if (resid_h) {   // raw bits now, unpacked where the row is used: the load latency stays behind the LayerNorm math
                const uint2 h = __ldg(reinterpret_cast<const uint2*>(resid_h + (size_t)row * 128 + c0 + oseg * 4));
                xr[ps].x = __uint_as_float(h.x);
                xr[ps].y = __uint_as_float(h.y);
              } else {
                xr[ps] = __ldg(reinterpret_cast<const float4*>(resid + (size_t)row * 128 + c0 + oseg * 4));
              }
            }
          }
          const float4 b = *reinterpret_cast<const float4*>(sb3 + c0 + oseg * 4);
          const float4 gm = *reinterpret_cast<const float4*>(sg + c0 + oseg * 4);
          const float4 bt = *reinterpret_cast<const float4*>(sbeta + c0 + oseg * 4);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(wstg_at(mystg, lane, j >> 2)) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
          __syncwarp();
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const int rr = ps * 4 + orow;
            const int64_t row = wrow0 + rr;
            if (row < d.rows) {
              const float4 a = *reinterpret_cast<const float4*>(wstg_at(mystg, rr, oseg));
              float4 y;
#if FVGN_F32X2
              {  // ((acc + b) - mean) * rstd * gamma + beta, same operation order as the scalar form
                const float2 nm = splat2(nmr[ps]), rs = splat2(rsr[ps]);
                const float2 ya = __ffma2_rn(__fmul2_rn(__fadd2_rn(__fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)), nm), rs),
                                             make_float2(gm.x, gm.y), make_float2(bt.x, bt.y));
                const float2 yb = __ffma2_rn(__fmul2_rn(__fadd2_rn(__fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w)), nm), rs),
                                             make_float2(gm.z, gm.w), make_float2(bt.z, bt.w));
                y = make_float4(ya.x, ya.y, yb.x, yb.y);
              }
#else
              y.x = (a.x + b.x + nmr[ps]) * rsr[ps] * gm.x + bt.x;
              y.y = (a.y + b.y + nmr[ps]) * rsr[ps] * gm.y + bt.y;
              y.z = (a.z + b.z + nmr[ps]) * rsr[ps] * gm.z + bt.z;
              y.w = (a.w + b.w + nmr[ps]) * rsr[ps] * gm.w + bt.w;
#endif
              const size_t o = (size_t)row * 128 + c0 + oseg * 4;
              if (d.out) *reinterpret_cast<float4*>(d.out + o) = y;
              if (outh) *reinterpret_cast<uint2*>(outh + o) = make_uint2(pack16<P>(y.x, y.y), pack16<P>(y.z, y.w));
              if (do_res) {
                float4 x = xr[ps];
                if (resid_h) {
                  const uint32_t h0 = __float_as_uint(x.x), h1 = __float_as_uint(x.y);
                  x = make_float4(P::lo(h0), P::hi(h0), P::lo(h1), P::hi(h1));
                }
                const float4 t = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
                if (d.out_res) *reinterpret_cast<float4*>(d.out_res + o) = t;
                if (out_resh) *reinterpret_cast<uint2*>(out_resh + o) = make_uint2(pack16<P>(t.x, t.y), pack16<P>(t.z, t.w));
              }
            }
          }
          __syncwarp();
        }
      } else {
        // decoder: 3 outputs per row, no LayerNorm
        uint32_t r[16];
        tmem_ld16(acc, r);
        tmem_wait_ld();
        const int64_t row = row0 + rloc;
        if (row < d.rows) {
#pragma unroll
          for (int j = 0; j < 3; ++j) d.out[(size_t)row * 3 + j] = __uint_as_float(r[j]) + sb3[j];
        }
      }
      tc_fence_before();
      PROF_F(6);  // normalise + residual + write-out
#ifdef FVGN_TIMING
      if (blockIdx.x == 0 && tid == 0) g_prof_f[15] += 1;
#endif
    }
    if (lane == 0) bulk_wait0();  // the last Z1 bulk store must be done before shared memory is released
  }
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int MODE, class P>
int launch_tc_fwd(const fvgn_mlp_desc& d, void* stream) {
  auto kern = mlp_tc_fwd_kernel<MODE, P>;
  static bool attr_set[FVGN_MAX_DEV] = {false};  // the attribute is per device
  const int dev = fvgn_cur_device();
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<MODE>()) != cudaSuccess)
      return FVGN_ERR_LAUNCH;
    attr_set[dev] = true;
  }
  const int num_sms = fvgn_num_sms();
  const int64_t ntiles = (d.rows + TILE_M - 1) / TILE_M;
  const unsigned grid = (unsigned)(ntiles < num_sms ? ntiles : num_sms);
  kern<<<grid, NTHREADS, smem_bytes<MODE>(), (cudaStream_t)stream>>>(d);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

}  // namespace

#ifdef FVGN_TIMING
extern "C" int fvgn_debug_profile_f(unsigned long long* out16, int reset) {
  if (cudaMemcpyFromSymbol(out16, g_prof_f, sizeof(unsigned long long) * 16) != cudaSuccess) return FVGN_ERR_LAUNCH;
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_prof_f, z, sizeof(z));
  }
  return FVGN_OK;
}
#endif

int fvgn_mlp_forward_tc(const fvgn_mlp_desc* d, void* stream) {
  if (!d->w_bf16) return FVGN_ERR_NULL;
  if ((((uintptr_t)d->w_bf16) & 15) != 0) return FVGN_ERR_ALIGN;
  const bool f16 = d->precision == FVGN_PREC_F16;
#define FVGN_TC_FWD_CASE(M) case M: return f16 ? launch_tc_fwd<M, PF16>(*d, stream) : launch_tc_fwd<M, PBF16>(*d, stream);
  switch (d->mode) {
    FVGN_TC_FWD_CASE(FVGN_MLP_EDGE)
    FVGN_TC_FWD_CASE(FVGN_MLP_NODE)
    FVGN_TC_FWD_CASE(FVGN_MLP_ENC_NODE)
    FVGN_TC_FWD_CASE(FVGN_MLP_ENC_EDGE)
    FVGN_TC_FWD_CASE(FVGN_MLP_DEC)
  }
#undef FVGN_TC_FWD_CASE
  return FVGN_ERR_UNSUPPORTED;
}

int64_t fvgn_mlp_tc_packed_bytes(int32_t mode) {
  switch (mode) {
    case FVGN_MLP_EDGE: return image_bytes(384);
    case FVGN_MLP_NODE: return image_bytes(192);
    case FVGN_MLP_ENC_NODE: return image_bytes(16);
    case FVGN_MLP_ENC_EDGE: return image_bytes(16);
    case FVGN_MLP_DEC: return image_bytes(128);
  }
  return -1;
}

int fvgn_mlp_tc_pack(int32_t mode, int32_t precision, const float* w1, const float* w2, const float* w3, void* packed,
                     void* stream) {
  int k1, k1p, nout = 128;
  switch (mode) {
    case FVGN_MLP_EDGE: k1 = 384; k1p = 384; break;
    case FVGN_MLP_NODE: k1 = 192; k1p = 192; break;
    case FVGN_MLP_ENC_NODE: k1 = 12; k1p = 16; break;
    case FVGN_MLP_ENC_EDGE: k1 = 15; k1p = 16; break;
    case FVGN_MLP_DEC: k1 = 128; k1p = 128; nout = 3; break;
    default: return FVGN_ERR_UNSUPPORTED;
  }
  if (!w1 || !w2 || !w3 || !packed) return FVGN_ERR_NULL;
  if (precision == FVGN_PREC_F16)
    pack_weights_kernel<PF16><<<64, 256, 0, (cudaStream_t)stream>>>(w1, w2, w3, k1, k1p, nout, reinterpret_cast<uint8_t*>(packed));
  else
    pack_weights_kernel<PBF16><<<64, 256, 0, (cudaStream_t)stream>>>(w1, w2, w3, k1, k1p, nout, reinterpret_cast<uint8_t*>(packed));
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}
