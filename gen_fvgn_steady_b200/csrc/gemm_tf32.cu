// Dense projections of the Transolver block on tcgen05 tensor cores with kind::tf32 operands -- fp32 data in shared memory
// read by the tensor core as TF32 (10-bit mantissa), fp32 accumulation in TMEM: the arithmetic the reference's GPU scripts
// select for every nn.Linear with torch.backends.cuda.matmul.allow_tf32 (src/pre_train_Adam.py:29).  Replaces the library
// GEMMs of GraphTransolver.py:48-58 (in_project_fx / in_project_x), :92-95 (to_out), :105-127 (mlp.linear_pre / linear_post)
// and of their autograd in the tensor-core precision modes (f16 / bf16); the fp32 parity mode keeps exact fp32 GEMMs.
//
//   fvgn_gemm_tf32(mode, ...):
//     FVGN_GEMM_NT  C[M,N]   = A[M,K] B[N,K]^T (+ bias[N]) (+ addend[M,N])     y = x W^T + b            (forward)
//     FVGN_GEMM_NN  C[M,N]   = A[M,K] B[K,N]   (+ addend[M,N])                 dx = dy W (+ residual gradient)
//     FVGN_GEMM_TN  C[Mo,N]  = A[R,Mo]^T B[R,N]                                dW = dy^T x   (deterministic: one partial per
//                                                                               CTA over a static row split, fixed-order sum)
//   all matrices fp32 row-major, K / N / Mo in {128, 256}.
//
// One persistent CTA per SM.  Operands are streamed in stages of 32 contraction indices: 4 producer warps issue 16-byte
// cp.async copies straight into the canonical SWIZZLE_128B layouts (K-major: [rows][32 k] with 8-row groups 1024 B apart;
// MN-major: [32 k-rows][32-column blocks] in the 32-byte-based 128B swizzle -- row-major global rows ARE MN-major, so the
// transposed products need no transposition anywhere), one thread issues tcgen05.mma (M = 128, N = 128 / 256, K = 8 per instruction), four epilogue warps
// drain the double-buffered TMEM accumulator through a swizzled staging tile into coalesced 128-byte row segments.
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int G_THREADS = 13 * 32;  // warps 0-3 and 9-12: epilogue (two per TMEM lane quarter), 4-7: producers, 8: MMA
constexpr int G_STAGES = 4;
constexpr int KSTEP = 32;           // contraction indices per stage (128 B of fp32)
constexpr uint32_t TF32 = 2;        // InstrDescriptor a_format / b_format

__device__ __forceinline__ void cp_async16z(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

// MN-major operands of 32-bit element types use the SWIZZLE_128B_BASE32B canonical layout (cute::UMMA::Layout_MN_SW128_32B_Atom:
// Swizzle<2,5,2> on byte addresses): atoms of 4 k-rows x 128 B, inside row r the 32-byte chunk j sits at j ^ (r & 3).
__device__ __forceinline__ uint32_t sw32b_off(int kr, int c) {   // c = 16-byte chunk (0..7) of the 128-byte row
  return (uint32_t)((kr >> 2) * 512 + (kr & 3) * 128 + ((((c >> 1) ^ (kr & 3)) << 5) | ((c & 1) << 4)));
}
// descriptor of such an operand: LBO = stride between 32-column blocks, SBO = 512 B between 4-row groups, layout type 1
__device__ __forceinline__ uint64_t make_desc_mn_tf32(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (32ull << 32) | (1ull << 46) |
         (1ull << 61);
}

// copy a [rows x ncols] fp32 block (ncols a multiple of 32) of a row-major matrix into shared memory as ncols / 32 blocks
// of [rows][128 B]; K-major use (MN = false): every block in the SWIZZLE_128B layout (16-byte chunk c of row r at
// sw128_off(r, c)); MN-major use (rows = contraction index): the 32-byte-based layout above.  Rows at or past `row_end` are
// zero-filled.  Issued by the 128 producer threads.
template <bool MN>
__device__ __forceinline__ void load_block(uint32_t dst, const float* __restrict__ g, int64_t ld, int64_t row0, int64_t row_end,
                                           int rows, int col0, int ncols, int pt) {
  const int chunks_per_row = ncols >> 2;             // 16-byte chunks per row
  const int total = rows * chunks_per_row;
  const uint32_t block_bytes = (uint32_t)rows * 128u;
  for (int i = pt; i < total; i += 128) {
    const int r = i / chunks_per_row, cc = i - r * chunks_per_row;
    const int cb = cc >> 3, c = cc & 7;
    const int64_t row = row0 + r;
    const bool live = row < row_end;
    cp_async16z(dst + cb * block_bytes + (MN ? sw32b_off(r, c) : sw128_off(r, c)), g + (live ? row : 0) * ld + col0 + cc * 4,
                live ? 16 : 0);
  }
}

struct GemmArgs {
  const float* A; const float* B; const float* bias; const float* addend; float* C; float* partials;
  int64_t M;       // rows of A (NT / NN) or contraction rows R (TN)
  int32_t N, K;    // NT / NN: C is [M,N], contraction K;  TN: C is [K,N] with K = Mo (columns of A)
};

template <int MODE>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_tf32_kernel(const GemmArgs a) {
  FVGN_DYN_SMEM(smem);
  const int N = a.N, K = a.K;
  // stage = A part | B part
  const uint32_t a_bytes = (MODE == FVGN_GEMM_TN) ? (uint32_t)K * 128u : 128u * 128u;   // TN: [32 r][K cols]; else [128 rows][32 k]
  const uint32_t b_bytes = (uint32_t)N * 128u;                                           // [N rows][32 k] or [32 k][N cols]
  // NT / NN: the whole B operand (the layer's weight, <= 128 KB) stays resident as K / 32 blocks; only A is streamed.
  // TN: both operands are streamed.
  const int nkB = (MODE == FVGN_GEMM_TN) ? 0 : K / KSTEP;
  const uint32_t stage_bytes = (MODE == FVGN_GEMM_TN) ? a_bytes + b_bytes : a_bytes;
  uint8_t* bres = smem;
  uint8_t* stages = bres + (size_t)nkB * b_bytes;
  uint8_t* stg = stages + G_STAGES * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + 8 * WSTG_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_FULL = 0, B_EMPTY = G_STAGES, B_ACCFULL = 2 * G_STAGES, B_ACCFREE = 2 * G_STAGES + 2, B_DONE = 2 * G_STAGES + 4,
                B_WREADY = 2 * G_STAGES + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G_STAGES + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(BAR(B_FULL + s), 128);
      mbar_init(BAR(B_EMPTY + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(B_ACCFULL + s), 1);
      mbar_init(BAR(B_ACCFREE + s), 256);
    }
    mbar_init(BAR(B_DONE), 1);
    mbar_init(BAR(B_WREADY), 128);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // work items: NT / NN: 128-row tiles of C;  TN: chunks of 32 contraction rows, every CTA accumulates its own chunks
  const int64_t nwork = (MODE == FVGN_GEMM_TN) ? (a.M + KSTEP - 1) / KSTEP : (a.M + TILE_M - 1) / TILE_M;
  const int nk = (MODE == FVGN_GEMM_TN) ? 1 : K / KSTEP;   // stages per work item

  if (warp >= 4 && warp < 8) {
    // =============================================================== producers
    const int pt = tid - 128;
    uint32_t it = 0;
    int pending = 0;   // stages whose copies were committed but not yet signalled
    auto signal_oldest = [&](uint32_t its) {
      fence_proxy_async();
      mbar_arrive(BAR(B_FULL + (its % G_STAGES)));
    };
    uint32_t sig = 0;   // next stage iteration to signal
    if (MODE != FVGN_GEMM_TN) {   // resident B: every k-block once
      for (int kb = 0; kb < nkB; ++kb) {
        const uint32_t sb = smem_u32(bres + (size_t)kb * b_bytes);
        if (MODE == FVGN_GEMM_NT) load_block<false>(sb, a.B, K, 0, N, N, kb * KSTEP, KSTEP, pt);          // W rows (N), k-block
        else load_block<true>(sb, a.B, N, (int64_t)kb * KSTEP, K, KSTEP, 0, N, pt);                       // 32 k-rows of B[K,N]
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(BAR(B_WREADY));
    }
    for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
      for (int kb = 0; kb < nk; ++kb, ++it) {
        const int s = it % G_STAGES;
        mbar_wait(BAR(B_EMPTY + s), ((it / G_STAGES) & 1) ^ 1);
        const uint32_t sa = smem_u32(stages + s * stage_bytes), sb = sa + a_bytes;
        if (MODE != FVGN_GEMM_TN) {
          load_block<false>(sa, a.A, K, w * TILE_M, a.M, TILE_M, kb * KSTEP, KSTEP, pt);      // A rows, k-block
        } else {
          load_block<true>(sa, a.A, K, w * KSTEP, a.M, KSTEP, 0, K, pt);                 // 32 rows of A[R,Mo]
          load_block<true>(sb, a.B, N, w * KSTEP, a.M, KSTEP, 0, N, pt);                 // 32 rows of B[R,N]
        }
        cp_async_commit();
        if (++pending == G_STAGES - 1) {   // the oldest outstanding stage has landed once at most G_STAGES-2 are newer
          cp_async_wait<G_STAGES - 2>();
          signal_oldest(sig++);
          --pending;
        }
      }
    }
    while (pending > 0) {   // drain
      cp_async_wait<0>();
      signal_oldest(sig++);
      --pending;
    }
  } else if (warp == 8) {
    // =============================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TF32, N, MODE == FVGN_GEMM_TN ? 1 : 0, MODE == FVGN_GEMM_NT ? 0 : 1);
      uint32_t it = 0, t = 0;
      if (MODE != FVGN_GEMM_TN) {
        mbar_wait(BAR(B_WREADY), 0);
        tc_fence_after();
      }
      for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x, ++t) {
        const int ab = (MODE == FVGN_GEMM_TN) ? 0 : (int)(t & 1);
        if (MODE != FVGN_GEMM_TN) {
          mbar_wait(BAR(B_ACCFREE + ab), ((t >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % G_STAGES;
          mbar_wait(BAR(B_FULL + s), (it / G_STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(stages + s * stage_bytes);
          const uint32_t sb = (MODE == FVGN_GEMM_TN) ? sa + a_bytes : smem_u32(bres + (size_t)kb * b_bytes);
          if (MODE == FVGN_GEMM_TN) {
            // C[Mo,N] += A^T B: both operands MN-major, 32-column blocks 32 x 128 B = 4096 B apart, 8 k-rows per step
            for (int h = 0; h < K / 128; ++h)
              for (int k = 0; k < 4; ++k)
                umma_tf32(tmem + h * N, make_desc_mn_tf32(sa + h * 4 * 4096, 4096) + 64 * k, make_desc_mn_tf32(sb, 4096) + 64 * k,
                          idesc, !(t == 0 && k == 0));
          } else {
            const uint32_t acc = tmem + ab * N;
            for (int k = 0; k < 4; ++k) {
              const uint64_t bd = (MODE == FVGN_GEMM_NT) ? make_desc_k128(sb) + 2 * k : make_desc_mn_tf32(sb, 4096) + 64 * k;
              umma_tf32(acc, make_desc_k128(sa) + 2 * k, bd, idesc, (kb | k) != 0);
            }
          }
          umma_commit(BAR(B_EMPTY + s));
        }
        if (MODE != FVGN_GEMM_TN) umma_commit(BAR(B_ACCFULL + ab));
      }
      umma_commit(BAR(B_DONE));
    }
    __syncwarp();
  } else {
    // =============================================================== epilogue (8 warps: warp pairs (q, q + 9 - ...) share a lane
    // quarter q = warp & 3 and alternate over the 32-column passes)
    const int q = warp & 3, grp = warp < 4 ? 0 : 1;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint8_t* mystg = stg + (grp * 4 + q) * WSTG_BYTES;
    const int orow = lane >> 3, oseg = lane & 7;
    if (MODE != FVGN_GEMM_TN) {
      uint32_t t = 0;
      for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x, ++t) {
        const int ab = t & 1;
        const int64_t wrow0 = w * TILE_M + q * 32;
        mbar_wait(BAR(B_ACCFULL + ab), (t >> 1) & 1);
        tc_fence_after();
        const uint32_t tacc = tmem + lane_base + ab * N;
#pragma unroll 1
        for (int c0 = 32 * grp; c0 < N; c0 += 64) {   // 32 fp32 columns = 128 B per row and pass; the two groups alternate
          uint32_t r[32];
          tmem_ld32(tacc + c0, r);
          tmem_wait_ld();
          if (c0 + 64 >= N) {   // this warp's last pass: the accumulator may be overwritten by the tile after next
            tc_fence_before();
            mbar_arrive(BAR(B_ACCFREE + ab));
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(wstg_at(mystg, lane, k)) =
                make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]),
                            __uint_as_float(r[4 * k + 3]));
          __syncwarp();
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.bias) bv = __ldg(reinterpret_cast<const float4*>(a.bias + c0 + oseg * 4));
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const int rr = ps * 4 + orow;
            const int64_t row = wrow0 + rr;
            if (row < a.M) {
              float4 v = *reinterpret_cast<const float4*>(wstg_at(mystg, rr, oseg));
              v = make_float4(v.x + bv.x, v.y + bv.y, v.z + bv.z, v.w + bv.w);
              const size_t off = (size_t)row * N + c0 + oseg * 4;
              if (a.addend) {
                const float4 ad = __ldg(reinterpret_cast<const float4*>(a.addend + off));
                v = make_float4(v.x + ad.x, v.y + ad.y, v.z + ad.z, v.w + ad.w);
              }
              *reinterpret_cast<float4*>(a.C + off) = v;
            }
          }
          __syncwarp();
        }
      }
    } else {
      // flush this CTA's accumulators [Mo, N] (TMEM lane = row within a 128-row half) into its partial
      mbar_wait(BAR(B_DONE), 0);
      tc_fence_after();
      float* Pc = a.partials + (size_t)blockIdx.x * ((size_t)K * N);
      const bool any = (int64_t)blockIdx.x < nwork;
      for (int h = 0; h < K / 128; ++h) {
        const int o = h * 128 + q * 32 + lane;
#pragma unroll 1
        for (int c0 = 16 * grp; c0 < N; c0 += 32) {
          uint32_t r[16];
          tmem_ld16(tmem + lane_base + h * N + c0, r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) Pc[(size_t)o * N + c0 + j] = any ? __uint_as_float(r[j]) : 0.f;
        }
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

__global__ void __launch_bounds__(256) gemm_partial_reduce_kernel(const float* __restrict__ partials, int n_partials, int64_t count,
                                                                  float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int g = 0; g < n_partials; ++g) s += partials[(size_t)g * count + i];
  out[i] = s;
}

int gemm_smem(int mode, int N, int K) {
  if (mode == FVGN_GEMM_TN) return G_STAGES * (K * 128 + N * 128) + 8 * WSTG_BYTES + 256;
  return (K / KSTEP) * N * 128 + G_STAGES * 128 * 128 + 8 * WSTG_BYTES + 256;   // resident B + streamed A
}

template <int MODE>
int launch_gemm(const GemmArgs& a, int grid, void* stream) {
  auto kern = gemm_tf32_kernel<MODE>;
  static bool attr_set[FVGN_MAX_DEV] = {false};
  const int dev = fvgn_cur_device();
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return FVGN_ERR_LAUNCH;
    attr_set[dev] = true;
  }
  kern<<<(unsigned)grid, G_THREADS, gemm_smem(MODE, a.N, a.K), (cudaStream_t)stream>>>(a);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

}  // namespace

extern "C" int32_t fvgn_gemm_tf32_partials(int64_t rows) {
  const int64_t nwork = (rows + KSTEP - 1) / KSTEP;
  const int sms = fvgn_num_sms();
  return (int32_t)(nwork < 1 ? 1 : (nwork < sms ? nwork : sms));
}

extern "C" int fvgn_gemm_tf32(int32_t mode, const float* A, const float* B, const float* bias, const float* addend, float* C,
                              int64_t rows, int32_t n, int32_t k, float* partials, int32_t n_partials, void* stream) {
  if (rows < 0 || (n != 128 && n != 256) || (k != 128 && k != 256)) return FVGN_ERR_SHAPE;
  if (!A || !B || !C) return FVGN_ERR_NULL;
  if (!fvgn_aligned16(A) || !fvgn_aligned16(B) || !fvgn_aligned16(C) || !fvgn_aligned16(bias) || !fvgn_aligned16(addend))
    return FVGN_ERR_ALIGN;
  GemmArgs a{A, B, bias, addend, C, partials, rows, n, k};
  const int sms = fvgn_num_sms();
  if (mode == FVGN_GEMM_TN) {
    if (bias || addend) return FVGN_ERR_UNSUPPORTED;
    if (!partials || n_partials != fvgn_gemm_tf32_partials(rows)) return FVGN_ERR_SHAPE;
    if ((int64_t)k / 128 * n > 512) return FVGN_ERR_UNSUPPORTED;   // accumulators must fit the 512 TMEM columns
    if (gemm_smem(mode, n, k) > 227 * 1024) return FVGN_ERR_UNSUPPORTED;
    if (rows == 0) {
      if (cudaMemsetAsync(C, 0, (size_t)k * n * sizeof(float), (cudaStream_t)stream) != cudaSuccess) return FVGN_ERR_LAUNCH;
      return FVGN_OK;
    }
    int rc = launch_gemm<FVGN_GEMM_TN>(a, n_partials, stream);
    if (rc) return rc;
    const int64_t count = (int64_t)k * n;
    gemm_partial_reduce_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partials, n_partials, count, C);
    FVGN_CHECK_LAUNCH();
    return FVGN_OK;
  }
  if (rows == 0) return FVGN_OK;
  if (gemm_smem(mode, n, k) > 227 * 1024) return FVGN_ERR_UNSUPPORTED;
  const int64_t ntiles = (rows + TILE_M - 1) / TILE_M;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  if (mode == FVGN_GEMM_NT) return launch_gemm<FVGN_GEMM_NT>(a, grid, stream);
  if (mode == FVGN_GEMM_NN) {
    if (bias) return FVGN_ERR_UNSUPPORTED;
    return launch_gemm<FVGN_GEMM_NN>(a, grid, stream);
  }
  return FVGN_ERR_UNSUPPORTED;
}
