// Node-level part of the edge MLP's layer-1 backward (tensor-core modes; fvgn_mlp_desc.d_aggh != NULL).
//
// The first 256 input columns of the edge MLP are agg[senders] | agg[receivers] (blocks.py:101-107).  By linearity
//     d(agg)[i]      = sum_{f: s_f = i} dZ1[f] W1[:, 0:128] + sum_{f: r_f = i} dZ1[f] W1[:, 128:256]
//                    = U_s[i] W1[:, 0:128] + U_r[i] W1[:, 128:256],        U_s[i] = sum_{f: s_f = i} dZ1[f],  U_r likewise
//     dW1[:, 0:128]  = sum_f dZ1[f]^T agg[s_f] = U_s^T agg,                dW1[:, 128:256] = U_r^T agg
// so this part of the backward runs on N node rows instead of E edge rows: the [E,256] gradient stream
// d(agg[s]) | d(agg[r]) (512 B per edge written by kernel B and read back by the incidence reduction), the four gathered
// operand chunks of kernel B's weight gradient and four of its six dgrad blocks disappear.
//
// One persistent CTA per SM, 128-node tiles:
//   warps 4-11  gather : U_s / U_r rows = incidence sums of the dZ1 tile images kernel A wrote (16 lanes x 16 B per dZ1
//                        row, fp32 accumulation in CSR order = the order of the reference's scatter), rounded to the
//                        16-bit operand format into two canonical SWIZZLE_128B tiles; the node's aggh rows are copied
//                        into a third tile.  Row pointers / incidence codes of the following tiles arrive by cp.async.
//   warp 12     MMA    : d(agg) tile = U_s W1[blocks 0,1] + U_r W1[blocks 2,3]  (A K-major, weight image read MN-major,
//                        as the dgrad of kernel A), dW1a += U_s^T agg, dW1b += U_r^T agg (both tiles read MN-major, K =
//                        the 128 nodes, as the wgrad of kernel A); accumulators in TMEM (2 x 128 + 2 x 128 columns).
//   warps 0-3   epilogue: d(agg) accumulator -> 16-bit -> coalesced 128-B row segments of d_aggh [N,128].
// Deterministic: static tile -> CTA map, one weight-gradient partial per CTA, fixed-order reduction.
#include "tc_common.cuh"

using namespace tc;

namespace {

#ifndef FVGN_NODE_GATHER_WARPS
#define FVGN_NODE_GATHER_WARPS 16
#endif
constexpr int N_GWARPS = FVGN_NODE_GATHER_WARPS;
constexpr int N_THREADS = (4 + N_GWARPS + 1) * 32;   // 4 epilogue + gather + 1 MMA warps
constexpr int N_GATHER = N_GWARPS * 32;              // gather threads
constexpr int N_GROUPS = N_GATHER / 16;              // 16-lane groups
constexpr int W_MMA = 4 + N_GWARPS;
constexpr int BUF_BYTES = 2 * KB_BYTES;   // one [128 x 128] 16-bit tile
constexpr int NECAP = 1024;               // incidence entries of a tile staged in shared memory (rest: global)
constexpr int R = 2;                      // nodes per 16-lane group in flight
constexpr int NB = 4;                     // gathers in flight per node

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void gather_bar() { asm volatile("bar.sync 2, %0;" ::"n"(N_GATHER) : "memory"); }
// 16-byte asynchronous copy global -> shared; src_bytes = 0 writes zeros
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem),
               "r"(src_bytes)
               : "memory");
}

constexpr int smem_node() { return 4 * KB_BYTES + 3 * BUF_BYTES + 4 * WSTG_BYTES + 3 * 132 * 4 + 2 * NECAP * 4 + 256; }

// Two-kernel variant (fvgn_mlp_desc.node_ws != NULL): the incidence sums are formed by a separate, many-CTA kernel that is
// bound by memory rather than by one CTA's gather latency, and written as ready-made operand tile images:
// node tile t -> [U_s kb0 | U_s kb1 | U_r kb0 | U_r kb1], 4 x 16 KB at node_ws + t * 64 KB.
template <class P, int R, int MINB>
__global__ void __launch_bounds__(256, MINB) dz_incidence_kernel(const uint8_t* __restrict__ dz_img, const int32_t* __restrict__ ptr,
                                                                  const int32_t* __restrict__ ent, uint8_t* __restrict__ u_img, int64_t n) {
  constexpr int RB = TILE_M, ECAP = 1536, NB = 4, G = 16;
  __shared__ int s_ptr[3][RB + 1];
  __shared__ int s_ent[2][ECAP];
  const int tid = threadIdx.x, l16 = tid & 15, g = tid >> 4;
  const int kb = l16 >> 3, ch = l16 & 7;
  const int64_t nblk = (n + RB - 1) / RB;
  auto fetch_ptr = [&](int64_t blk, int buf) {
    if (blk < nblk && tid <= RB) {
      const int64_t row = blk * RB + tid;
      cp_async4(&s_ptr[buf][tid], ptr + (row < n ? row : n));
    }
  };
  auto fetch_ent = [&](int64_t blk, int pbuf, int ebuf) {
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
    const int nr = (int)min((int64_t)RB, n - blk * RB);
    auto entry = [&](int t) { return (t - e0 < ECAP) ? se[t - e0] : __ldg(ent + t); };
    auto dz_row = [&](int f) {   // this lane's 16 bytes of row f of the dZ1 tile images (8 columns)
      return reinterpret_cast<const uint4*>(dz_img + (size_t)(f >> 7) * BUF_BYTES + kb * KB_BYTES + sw128_off(f & 127, ch));
    };
    uint8_t* ut = u_img + (size_t)blk * (2 * BUF_BYTES);
#pragma unroll 1
    for (int r0 = g; r0 < RB; r0 += G * R) {   // rows past the last node are written as zeros (they are MMA operands)
      int b[R], deg[R], c[R][NB];
      uint4 v[R][NB];
#pragma unroll
      for (int r = 0; r < R; ++r) {   // issue phase: the first NB gathers of R rows
        const int rr = r0 + r * G;
        b[r] = 0; deg[r] = 0;
        if (rr < nr) { b[r] = sp[rr]; deg[r] = sp[rr + 1] - b[r]; }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          c[r][k] = 0;
          if (k < deg[r]) {
            c[r][k] = entry(b[r] + k);
            v[r][k] = __ldg(dz_row(c[r][k] >> 1));
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int rr = r0 + r * G;
        float as[8], ar[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { as[j] = 0.f; ar[j] = 0.f; }
        // role mask as a multiplier (adding 0 * x is exact): the two 16-lane groups of a warp never diverge
        auto add = [&](const uint4& w, int code) {
          const float x[8] = {P::lo(w.x), P::hi(w.x), P::lo(w.y), P::hi(w.y), P::lo(w.z), P::hi(w.z), P::lo(w.w), P::hi(w.w)};
          const float mr = (float)(code & 1), ms = 1.f - mr;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            as[j] = fmaf(ms, x[j], as[j]);
            ar[j] = fmaf(mr, x[j], ar[j]);
          }
        };
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k < deg[r]) add(v[r][k], c[r][k]);
        for (int t0 = NB; t0 < deg[r]; t0 += NB) {
          uint4 w[NB];
          int cc[NB];
#pragma unroll
          for (int k = 0; k < NB; ++k) {
            cc[k] = 0;
            if (t0 + k < deg[r]) {
              cc[k] = entry(b[r] + t0 + k);
              w[k] = __ldg(dz_row(cc[k] >> 1));
            }
          }
#pragma unroll
          for (int k = 0; k < NB; ++k)
            if (t0 + k < deg[r]) add(w[k], cc[k]);
        }
        if (rr < RB) {
          const uint32_t off = kb * KB_BYTES + sw128_off(rr, ch);
          *reinterpret_cast<uint4*>(ut + off) =
              make_uint4(pack16<P>(as[0], as[1]), pack16<P>(as[2], as[3]), pack16<P>(as[4], as[5]), pack16<P>(as[6], as[7]));
          *reinterpret_cast<uint4*>(ut + BUF_BYTES + off) =
              make_uint4(pack16<P>(ar[0], ar[1]), pack16<P>(ar[2], ar[3]), pack16<P>(ar[4], ar[5]), pack16<P>(ar[6], ar[7]));
        }
      }
    }
    cp_async_commit_wait_all();
    __syncthreads();
  }
}

template <class P, bool IMG>
__global__ void __launch_bounds__(N_THREADS, 1) mlp_tc_bwd_node_kernel(const fvgn_mlp_desc d) {
  constexpr uint32_t IDESC_KM = make_idesc(P::FMT, 128, 0, 1), IDESC_MM = make_idesc(P::FMT, 128, 1, 1);
  constexpr uint32_t DWA = 0, DWB = 128, ACC = 256;
  FVGN_DYN_SMEM(smem);
  uint8_t* w_img = smem;                        // W1 image blocks 0..3 (input columns 0..255)
  uint8_t* us = w_img + 4 * KB_BYTES;           // U_s tile
  uint8_t* ur = us + BUF_BYTES;                 // U_r tile
  uint8_t* ag = ur + BUF_BYTES;                 // aggh tile
  uint8_t* stg = ag + BUF_BYTES;
  int* s_ptr = reinterpret_cast<int*>(stg + 4 * WSTG_BYTES);   // [3][132]
  int* s_ent = s_ptr + 3 * 132;                                 // [2][NECAP]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ent + 2 * NECAP);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_W = 0, B_UFULL = 1, B_UFREE = 2, B_ACCFULL = 3, B_ACCFREE = 5, B_UTX = 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n = d.n_nodes;
  const int64_t ntiles = (n + TILE_M - 1) / TILE_M;
  const uint8_t* dz_img = reinterpret_cast<const uint8_t*>(d.workspace);
  const uint8_t* aggh = reinterpret_cast<const uint8_t*>(d.in0h);

  if (tid == 0) {
    mbar_init(BAR(B_W), 1);
    mbar_init(BAR(B_UFULL), 1);
    mbar_init(BAR(B_UFREE), 1);
    mbar_init(BAR(B_UTX), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(B_ACCFULL + s), 1);
      mbar_init(BAR(B_ACCFREE + s), 128);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == W_MMA) {
    // =============================================================== MMA issuer (+ weight loader)
    if (lane == 0) {
      mbar_expect_tx(BAR(B_W), 4 * KB_BYTES);
      for (int i = 0; i < 4; ++i)
        bulk_g2s(smem_u32(w_img + i * KB_BYTES), reinterpret_cast<const uint8_t*>(d.w_bf16) + (size_t)i * KB_BYTES, KB_BYTES, BAR(B_W));
      mbar_wait(BAR(B_W), 0);
      const uint32_t ws = smem_u32(w_img), uss = smem_u32(us), urs = smem_u32(ur), ags = smem_u32(ag);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int ab = it & 1;
        mbar_wait(BAR(B_UFULL), it & 1);
        if (IMG) mbar_wait(BAR(B_UTX), it & 1);
        mbar_wait(BAR(B_ACCFREE + ab), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem + ACC + 128 * ab;
        // d(agg) = U_s W1[:, 0:128] + U_r W1[:, 128:256]   (A: U tile K-major; B: weight image read MN-major)
        for (int k = 0; k < 8; ++k)
          umma_ss(acc, make_desc_k128(uss + (k >> 2) * KB_BYTES) + 2 * (k & 3), make_desc_mn128(ws, KB_BYTES) + 128 * k, IDESC_KM,
                  k != 0);
        for (int k = 0; k < 8; ++k)
          umma_ss(acc, make_desc_k128(urs + (k >> 2) * KB_BYTES) + 2 * (k & 3),
                  make_desc_mn128(ws + 2 * KB_BYTES, KB_BYTES) + 128 * k, IDESC_KM, 1);
        umma_commit(BAR(B_ACCFULL + ab));
        // dW1a += U_s^T agg, dW1b += U_r^T agg   (both tiles read MN-major, K = the 128 nodes of the tile)
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem + DWA, make_desc_mn128(uss, KB_BYTES) + 128 * k, make_desc_mn128(ags, KB_BYTES) + 128 * k, IDESC_MM,
                  !(it == 0 && k == 0));
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem + DWB, make_desc_mn128(urs, KB_BYTES) + 128 * k, make_desc_mn128(ags, KB_BYTES) + 128 * k, IDESC_MM,
                  !(it == 0 && k == 0));
        umma_commit(BAR(B_UFREE));
      }
      umma_commit(BAR(B_W));   // every MMA has retired: the epilogue may read dW1a / dW1b
    }
    __syncwarp();
  } else if (warp >= 4) {
    // =============================================================== gather: U_s, U_r, agg tiles
    const int tg = tid - 128;                 // 0 .. N_GATHER-1
    const int grp = tg >> 4, l16 = tg & 15;   // groups of 16 lanes; lane <-> 16-byte chunk of a 256-B row
    const int kb = l16 >> 3, ch = l16 & 7;
    auto fetch_ptr = [&](int64_t tile, int buf) {
      if (tile < ntiles && tg <= TILE_M) {
        const int64_t row = tile * TILE_M + tg;
        cp_async4(&s_ptr[buf * 132 + tg], d.inc_ptr + (row < n ? row : n));
      }
    };
    auto fetch_ent = [&](int64_t tile, int pbuf, int ebuf) {
      if (tile < ntiles) {
        const int e0 = s_ptr[pbuf * 132];
        const int ne = min(s_ptr[pbuf * 132 + TILE_M] - e0, NECAP);
        for (int i = tg; i < ne; i += N_GATHER) cp_async4(&s_ent[ebuf * NECAP + i], d.inc_code + e0 + i);
      }
    };
    int64_t tile = blockIdx.x;
    if (!IMG) {
      fetch_ptr(tile, 0);
      fetch_ptr(tile + gridDim.x, 1);
      cp_async_commit_wait_all();
      gather_bar();
      fetch_ent(tile, 0, 0);
      cp_async_commit_wait_all();
      gather_bar();
    }
    for (uint32_t it = 0; tile < ntiles; tile += gridDim.x, ++it) {
      const int pb = it % 3, eb = it & 1;
      if (!IMG) {
        fetch_ptr(tile + 2 * (int64_t)gridDim.x, (it + 2) % 3);
        fetch_ent(tile + gridDim.x, (it + 1) % 3, eb ^ 1);
      }
      const int* sp = s_ptr + pb * 132;
      const int* se = s_ent + eb * NECAP;
      const int e0 = sp[0];
      const int64_t row0 = tile * TILE_M;
      const int nr = (int)min((int64_t)TILE_M, n - row0);
      auto entry = [&](int t) { return (t - e0 < NECAP) ? se[t - e0] : __ldg(d.inc_code + t); };
      auto dz_row = [&](int f) {   // this lane's 16 bytes of row f of the dZ1 tile images (8 columns)
        return reinterpret_cast<const uint4*>(dz_img + (size_t)(f >> 7) * BUF_BYTES + kb * KB_BYTES + sw128_off(f & 127, ch));
      };
      mbar_wait(BAR(B_UFREE), (it & 1) ^ 1);   // the MMAs of the previous tile have finished reading the three tiles
      if (IMG && tg == 0) {   // the two U tiles are ready-made images: two 32-KB bulk copies
        const uint8_t* ut = reinterpret_cast<const uint8_t*>(d.node_ws) + (size_t)tile * (2 * BUF_BYTES);
        mbar_expect_tx(BAR(B_UTX), 2 * BUF_BYTES);
        bulk_g2s(smem_u32(us), ut, BUF_BYTES, BAR(B_UTX));
        bulk_g2s(smem_u32(ur), ut + BUF_BYTES, BUF_BYTES, BAR(B_UTX));
      }
      // aggh rows of the tile -> swizzled tile: asynchronous 16-byte copies (no registers), zero-filled past the last node
      for (int idx = tg; idx < TILE_M * 16; idx += N_GATHER) {
        const int row = idx >> 4, c16 = idx & 15;
        const bool live = row < nr;
        cp_async16(ag + (c16 >> 3) * KB_BYTES + sw128_off(row, c16 & 7),
                   aggh + (size_t)(live ? row0 + row : 0) * 256 + c16 * 16, live ? 16 : 0);
      }
      // incidence sums, R nodes per group in flight
#pragma unroll 1
      for (int base = grp; !IMG && base < TILE_M; base += N_GROUPS * R) {
        int b_[R], deg[R], c[R][NB];
        uint4 v[R][NB];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int node = base + N_GROUPS * r;
          b_[r] = 0; deg[r] = 0;
          if (node < nr) { b_[r] = sp[node]; deg[r] = sp[node + 1] - b_[r]; }
#pragma unroll
          for (int k = 0; k < NB; ++k) {
            c[r][k] = 0;
            if (k < deg[r]) {
              c[r][k] = entry(b_[r] + k);
              v[r][k] = __ldg(dz_row(c[r][k] >> 1));
            }
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int node = base + N_GROUPS * r;
          float as[8], ar[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { as[j] = 0.f; ar[j] = 0.f; }
          auto add = [&](const uint4& w, int code) {
            const float x[8] = {P::lo(w.x), P::hi(w.x), P::lo(w.y), P::hi(w.y), P::lo(w.z), P::hi(w.z), P::lo(w.w), P::hi(w.w)};
            if (code & 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) ar[j] += x[j];
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) as[j] += x[j];
            }
          };
#pragma unroll
          for (int k = 0; k < NB; ++k)
            if (k < deg[r]) add(v[r][k], c[r][k]);
          for (int t0 = NB; t0 < deg[r]; t0 += NB) {   // nodes with more than NB incident faces
            uint4 w[NB];
            int cc[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
              cc[k] = 0;
              if (t0 + k < deg[r]) {
                cc[k] = entry(b_[r] + t0 + k);
                w[k] = __ldg(dz_row(cc[k] >> 1));
              }
            }
#pragma unroll
            for (int k = 0; k < NB; ++k)
              if (t0 + k < deg[r]) add(w[k], cc[k]);
          }
          if (node < TILE_M) {   // rows past the last node are written as zeros (they are MMA operands)
            const uint32_t off = kb * KB_BYTES + sw128_off(node, ch);
            *reinterpret_cast<uint4*>(us + off) =
                make_uint4(pack16<P>(as[0], as[1]), pack16<P>(as[2], as[3]), pack16<P>(as[4], as[5]), pack16<P>(as[6], as[7]));
            *reinterpret_cast<uint4*>(ur + off) =
                make_uint4(pack16<P>(ar[0], ar[1]), pack16<P>(ar[2], ar[3]), pack16<P>(ar[4], ar[5]), pack16<P>(ar[6], ar[7]));
          }
        }
      }
      cp_async_commit_wait_all();   // the agg tile, the next tile's entries and the pointers of the one after have landed
      fence_proxy_async();
      gather_bar();                 // every gather thread has written (and fenced) its part of the three tiles
      if (tg == 0) mbar_arrive(BAR(B_UFULL));
    }
  } else {
    // =============================================================== epilogue: d(agg) tile -> 16-bit rows
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint8_t* mystg = stg + warp * WSTG_BYTES;
    const int orow = lane >> 3, oseg = lane & 7;
    uint8_t* dh = reinterpret_cast<uint8_t*>(d.d_aggh);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const int64_t wrow0 = tile * TILE_M + warp * 32;
      mbar_wait(BAR(B_ACCFULL + ab), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem + lane_base + ACC + 128 * ab;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {   // 64 columns = 128 B of 16-bit values per row and pass
        uint32_t r[32];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          tmem_ld32(tacc + 64 * half + 32 * q, r);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(wstg_at(mystg, lane, q * 4 + k)) =
                make_uint4(pack16<P>(__uint_as_float(r[8 * k]), __uint_as_float(r[8 * k + 1])),
                           pack16<P>(__uint_as_float(r[8 * k + 2]), __uint_as_float(r[8 * k + 3])),
                           pack16<P>(__uint_as_float(r[8 * k + 4]), __uint_as_float(r[8 * k + 5])),
                           pack16<P>(__uint_as_float(r[8 * k + 6]), __uint_as_float(r[8 * k + 7])));
        }
        if (half == 1) {
          tc_fence_before();
          mbar_arrive(BAR(B_ACCFREE + ab));
        }
        __syncwarp();
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
          const int rr = ps * 4 + orow;
          const int64_t row = wrow0 + rr;
          if (row < n)
            *reinterpret_cast<uint4*>(dh + (size_t)row * 256 + half * 128 + oseg * 16) = *reinterpret_cast<const uint4*>(wstg_at(mystg, rr, oseg));
        }
        __syncwarp();
      }
    }
    // ---------------- flush dW1a | dW1b (TMEM lane = output feature o, column = input column c) -> this CTA's partial
    mbar_wait(BAR(B_W), 1);
    tc_fence_after();
    {
      float* Pn = d.node_partials + (size_t)blockIdx.x * (128 * 256);
      const int o = warp * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + lane_base + DWA + c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) Pn[(size_t)o * 256 + c0 + j] = __uint_as_float(r[j]);
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

// w1 gradient columns 0..255 = unscale * sum over the CTAs' partials (fixed order); w1 is [128, 384] row-major
__global__ void __launch_bounds__(256) node_partial_reduce_kernel(const float* __restrict__ partials, int n_partials,
                                                                  const float* __restrict__ unscale, float* __restrict__ d_w1) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= 128 * 256) return;
  float s = 0.f;
  for (int g = 0; g < n_partials; ++g) s += partials[(size_t)g * (128 * 256) + i];
  d_w1[(i >> 8) * 384 + (i & 255)] = unscale ? s * __ldg(unscale) : s;
}

}  // namespace

int fvgn_mlp_tc_node_partials(int64_t n_nodes) {
  const int64_t ntiles = (n_nodes + TILE_M - 1) / TILE_M;
  const int sms = fvgn_num_sms();
  return (int)(ntiles < 1 ? 1 : (ntiles < sms ? ntiles : sms));
}

int64_t fvgn_mlp_tc_node_ws_bytes(int64_t n_nodes) { return ((n_nodes + TILE_M - 1) / TILE_M) * (int64_t)(2 * BUF_BYTES); }

#ifndef FVGN_DZI_R
#define FVGN_DZI_R 1
#endif
#ifndef FVGN_DZI_MINB
#define FVGN_DZI_MINB 4   // 63 registers, no spills: 1.19 ms per 4 M nodes (5: 48 registers + spills 1.46, R = 2: 1.31; profiles/r2z_dzi_sweep.txt)
#endif

template <class P, int R, int MINB>
static int launch_dz_incidence(const fvgn_mlp_desc& d, void* stream) {
  auto kz = dz_incidence_kernel<P, R, MINB>;
  static unsigned inc_grid[FVGN_MAX_DEV] = {0};
  const int dev = fvgn_cur_device();
  if (inc_grid[dev] == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kz, 256, 0) != cudaSuccess || per_sm < 1) return FVGN_ERR_LAUNCH;
    inc_grid[dev] = (unsigned)(per_sm * fvgn_num_sms());
  }
  const int64_t ntiles = (d.n_nodes + TILE_M - 1) / TILE_M;
  kz<<<(unsigned)(ntiles < inc_grid[dev] ? ntiles : inc_grid[dev]), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint8_t*>(d.workspace), d.inc_ptr, d.inc_code, reinterpret_cast<uint8_t*>(d.node_ws), d.n_nodes);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

template <class P>
int launch_tc_bwd_node(const fvgn_mlp_desc& d, void* stream) {
  auto kn = mlp_tc_bwd_node_kernel<P, false>;
  auto ki = mlp_tc_bwd_node_kernel<P, true>;
  static bool attr_set[FVGN_MAX_DEV] = {false};
  const int dev = fvgn_cur_device();
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(kn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_node()) != cudaSuccess) return FVGN_ERR_LAUNCH;
    if (cudaFuncSetAttribute(ki, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_node()) != cudaSuccess) return FVGN_ERR_LAUNCH;
    attr_set[dev] = true;
  }
  if (d.node_ws) {
#ifdef FVGN_DZI_SWEEP   // tools/dzi_sweep.sh: variant picked per call
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("FVGN_DZI_VARIANT"); variant = e ? atoi(e) : 0; }
    int rc;
    switch (variant) {
      case 1: rc = launch_dz_incidence<P, 1, 4>(d, stream); break;
      case 2: rc = launch_dz_incidence<P, 2, 4>(d, stream); break;
      case 3: rc = launch_dz_incidence<P, 2, 3>(d, stream); break;
      case 4: rc = launch_dz_incidence<P, 1, 6>(d, stream); break;
      case 5: rc = launch_dz_incidence<P, 2, 5>(d, stream); break;
      default: rc = launch_dz_incidence<P, 1, 5>(d, stream); break;
    }
#else
    const int rc = launch_dz_incidence<P, FVGN_DZI_R, FVGN_DZI_MINB>(d, stream);
#endif
    if (rc != FVGN_OK) return rc;
    ki<<<(unsigned)d.n_node_partials, N_THREADS, smem_node(), (cudaStream_t)stream>>>(d);
  } else {
    kn<<<(unsigned)d.n_node_partials, N_THREADS, smem_node(), (cudaStream_t)stream>>>(d);
  }
  FVGN_CHECK_LAUNCH();
  node_partial_reduce_kernel<<<128, 256, 0, (cudaStream_t)stream>>>(d.node_partials, d.n_node_partials, d.grad_unscale, d.d_params);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}
template int launch_tc_bwd_node<PBF16>(const fvgn_mlp_desc&, void*);
template int launch_tc_bwd_node<PF16>(const fvgn_mlp_desc&, void*);
