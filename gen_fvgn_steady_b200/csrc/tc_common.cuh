// Shared tcgen05 / TMEM / mbarrier / bulk-copy PTX helpers and the UMMA shared-memory layout used by the
// bf16 tensor-core MLP kernels (mlp_tc.cu forward, mlp_tc_bwd.cu backward).
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace tc {

// ------------------------------------------------------------------------------------------ 16-bit operand formats
// The tensor-core path is templated on the operand format P of tcgen05.mma kind::f16:
//   PBF16 (FVGN_PREC_BF16): bfloat16, 8-bit significand, fp32 exponent range;
//   PF16  (FVGN_PREC_F16) : IEEE half, 11-bit significand -- the significand of TF32, the arithmetic the reference's GPU
//                           path runs its Linear layers in (src/pre_train_Adam.py:29) -- conversions round to nearest and
//                           saturate to +-65504 instead of overflowing to infinity (gradients are pre-scaled by a power of
//                           two chosen per backward pass, ops.GradScaleFn).
// Both accumulate in fp32 (TMEM); everything outside the MMA operands (bias, GELU, LayerNorm, residuals) is fp32.
struct PBF16 {
  static constexpr uint32_t FMT = 1;  // InstrDescriptor a_format / b_format: 1 = BF16
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  static __device__ __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
  static __device__ __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
  static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
  }
};
struct PF16 {
  static constexpr uint32_t FMT = 0;  // 0 = F16
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  static __device__ __forceinline__ float lo(uint32_t w) { return __low2float(*reinterpret_cast<const __half2*>(&w)); }
  static __device__ __forceinline__ float hi(uint32_t w) { return __high2float(*reinterpret_cast<const __half2*>(&w)); }
  static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
  }
};
template <class P> __device__ __forceinline__ uint32_t pack16(float lo, float hi) { return P::pack(lo, hi); }
template <class P> __device__ __forceinline__ uint32_t pack16(float2 v) { return P::pack(v.x, v.y); }
template <class P> __device__ __forceinline__ float2 unpack16(uint32_t w) { return make_float2(P::lo(w), P::hi(w)); }

constexpr int TILE_M = 128;
constexpr int KB_BYTES = 128 * 128;  // one K-block: 128 rows x 64 bf16 (128 B per row), SWIZZLE_128B

__host__ __device__ constexpr int nkb1(int k1p) { return (k1p + 63) / 64; }
// weight image = [W1 K-blocks][W2: 2 K-blocks][W3: 2 K-blocks]
__host__ __device__ constexpr int image_bytes(int k1p) { return (nkb1(k1p) + 4) * KB_BYTES; }

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// one non-blocking probe of a phase: true when the phase with this parity has completed
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();  // watchdog: a protocol bug must not hang the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 (bit 4), A / B format fmt (bits 7-9 / 10-12: 0 = f16,
// 1 = bf16), major-ness of A / B (bits 15 / 16: 0 = K-major), N >> 3 (bits 17-22), M >> 4 (bits 24-28) with M = 128
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// Throughput-mode GELU: tanh form evaluated with the MUFU.TANH approximation (6 instructions).
// |gelu_tanh - gelu_erf| <= 3e-4 absolute, below the bf16 rounding the value receives right afterwards.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(u), hx);
}
// value and derivative of the same tanh-form GELU
__device__ __forceinline__ void gelu_tanh_pair(float x, float& h, float& g) {
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(0.0356774081f, x2, 0.7978845608f));
  const float hx = 0.5f * x;
  h = fmaf(hx, t, hx);
  // d/dx [0.5 x (1 + t)] = 0.5 (1 + t) + 0.5 x (1 - t^2) (a + 3 b x^2)
  g = fmaf(hx * fmaf(-t, t, 1.0f), fmaf(0.1070322243f, x2, 0.7978845608f), fmaf(0.5f, t, 0.5f));
}

// byte offset of (row, 16-byte chunk) inside one K-block in the SWIZZLE_128B K-major layout
__device__ __host__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}


// MN-major SWIZZLE_128B descriptor over the same [64-feature block][row][128 B] storage: the contiguous dimension
// (features) is M/N, rows are K.  LBO = stride between 64-feature blocks, SBO = stride between 8-row groups.
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (64ull << 32) | (1ull << 46) |
         (2ull << 61);
}

// ------------------------------------------------------------------------------------------ producers
// pointer to the 64 fp32 source values of (tile row, K-block kb), or nullptr for an all-zero row
template <int MODE>
__device__ __forceinline__ const float* chunk_src(const fvgn_mlp_desc& d, int64_t row, int kb, int s, int r) {
  if (row >= d.rows) return nullptr;
  if (MODE == FVGN_MLP_EDGE) {
    if (kb < 2) return d.in0 + (size_t)s * 128 + kb * 64;
    if (kb < 4) return d.in0 + (size_t)r * 128 + (kb - 2) * 64;
    return d.in1 + (size_t)row * 128 + (kb - 4) * 64;
  } else if (MODE == FVGN_MLP_NODE) {
    if (kb == 0) return d.in0 + (size_t)row * 64;
    return d.in1 + (size_t)row * 128 + (kb - 1) * 64;
  } else {
    return d.in0 + (size_t)row * 128 + kb * 64;
  }
}

// Endpoint indices of the 8 tile rows a producer lane handles (rows i*16 + pw*4 + lane/8, i = 0..7), fetched once per
// tile so that the gathers of every chunk do not wait on a dependent index load.  The 8 lanes that share lane/8 handle
// the same rows, so lane (lane & 24) | i keeps the pair of row i and the others read it with a shuffle: 2 registers per
// tile instead of 16.
struct TileIdx {
  int s, r;
  __device__ __forceinline__ int sender(int i, int lane) const { return __shfl_sync(0xffffffffu, s, (lane & 24) | i); }
  __device__ __forceinline__ int receiver(int i, int lane) const { return __shfl_sync(0xffffffffu, r, (lane & 24) | i); }
};
template <int MODE>
__device__ __forceinline__ void load_tile_idx(const fvgn_mlp_desc& d, int64_t row0, int pw, int lane, TileIdx& idx) {
  idx.s = 0;
  idx.r = 0;
  if (MODE == FVGN_MLP_EDGE) {
    const int64_t row = row0 + (lane & 7) * 16 + pw * 4 + (lane >> 3);
    if (row < d.rows) {
      idx.s = __ldg(d.idx_s + row);
      idx.r = __ldg(d.idx_r + row);
    }
  }
}

template <int MODE, class P>
__device__ __forceinline__ void produce_chunk(const fvgn_mlp_desc& d, int64_t row0, int kb, uint8_t* stage, int pw, int lane,
                                              const TileIdx& idx) {
  if (MODE == FVGN_MLP_ENC_NODE || MODE == FVGN_MLP_ENC_EDGE) {
    // one thread per row: 16 bf16 (K padded to 16) = chunks 0 and 1 of the row
    const int rloc = pw * 32 + lane;
    const int64_t row = row0 + rloc;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
    if (row < d.rows) {
      if (MODE == FVGN_MLP_ENC_NODE) {
        const float4* p = reinterpret_cast<const float4*>(d.in0 + (size_t)row * 12);
        const float4 a = p[0], b = p[1], c = p[2];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
      } else {  // importer.py:54-78
        const int s = d.idx_s[row], r = d.idx_r[row];
        const float4* ps = reinterpret_cast<const float4*>(d.in0 + (size_t)s * 12);
        const float4* pr = reinterpret_cast<const float4*>(d.in0 + (size_t)r * 12);
        const float4 a = ps[0], b = ps[1], c = ps[2], e = pr[0], f = pr[1], g = pr[2];
        v[0] = a.x - e.x; v[1] = a.y - e.y; v[2] = a.z - e.z; v[3] = a.w - e.w;
        v[4] = b.x - f.x; v[5] = b.y - f.y; v[6] = b.z - f.z; v[7] = b.w - f.w;
        v[8] = c.x - g.x; v[9] = c.y - g.y; v[10] = c.z - g.z; v[11] = c.w - g.w;
        const float2 qs = *reinterpret_cast<const float2*>(d.in1 + (size_t)s * 2);
        const float2 qr = *reinterpret_cast<const float2*>(d.in1 + (size_t)r * 2);
        const float dx = qs.x - qr.x, dy = qs.y - qr.y;
        v[12] = dx; v[13] = dy; v[14] = sqrtf(dx * dx + dy * dy);
      }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c)
      *reinterpret_cast<uint4*>(stage + sw128_off(rloc, c)) =
          make_uint4(pack16<P>(v[c * 8 + 0], v[c * 8 + 1]), pack16<P>(v[c * 8 + 2], v[c * 8 + 3]),
                     pack16<P>(v[c * 8 + 4], v[c * 8 + 5]), pack16<P>(v[c * 8 + 6], v[c * 8 + 7]));
#pragma unroll
    for (int c = 2; c < 8; ++c)  // the rest of the 64-wide block must be zero (it is an MMA operand in the wgrad pass)
      *reinterpret_cast<uint4*>(stage + sw128_off(rloc, c)) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  // generic: 8 lanes per row (32 B of fp32 -> one 16-B bf16 chunk each), 4 rows per warp-instruction, 8 passes
  const int seg = lane & 7;
  float4 lo[8], hi[8];
  const float* src[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rloc = i * 16 + pw * 4 + (lane >> 3);
    const int64_t row = row0 + rloc;
    src[i] = chunk_src<MODE>(d, row, kb, 0, 0);  // only the encoders use this fp32 path (no gathers by index here)
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (src[i]) {
      lo[i] = __ldg(reinterpret_cast<const float4*>(src[i] + seg * 8));
      hi[i] = __ldg(reinterpret_cast<const float4*>(src[i] + seg * 8 + 4));
    } else {
      lo[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      hi[i] = lo[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rloc = i * 16 + pw * 4 + (lane >> 3);
    *reinterpret_cast<uint4*>(stage + sw128_off(rloc, seg)) =
        make_uint4(pack16<P>(lo[i].x, lo[i].y), pack16<P>(lo[i].z, lo[i].w), pack16<P>(hi[i].x, hi[i].y),
                   pack16<P>(hi[i].z, hi[i].w));
  }
}


// ------------------------------------------------------------------------------------------ bf16-shadow producers
// Layer-1 operand chunks read from the bf16 row-major shadows: every 64-feature chunk of a row is 128 contiguous bytes,
// fetched as 8 x 16 B by 8 lanes (no conversion).  A chunk is held in registers (8 x uint4 per thread) between the load
// and the store into the ring, so the global-load latency is decoupled from the depth of the shared-memory ring.
template <int MODE>
__device__ __forceinline__ const uint4* chunk_src_h(const fvgn_mlp_desc& d, int64_t row, int kb, int seg, int s, int r) {
  if (row >= d.rows) return nullptr;
  const uint8_t* p;
  if (MODE == FVGN_MLP_EDGE) {
    if (kb < 2) p = reinterpret_cast<const uint8_t*>(d.in0h) + (size_t)s * 256 + kb * 128;
    else if (kb < 4) p = reinterpret_cast<const uint8_t*>(d.in0h) + (size_t)r * 256 + (kb - 2) * 128;
    else p = reinterpret_cast<const uint8_t*>(d.in1h) + (size_t)row * 256 + (kb - 4) * 128;
  } else if (MODE == FVGN_MLP_NODE) {
    if (kb == 0) p = reinterpret_cast<const uint8_t*>(d.in0h) + (size_t)row * 128;
    else p = reinterpret_cast<const uint8_t*>(d.in1h) + (size_t)row * 256 + (kb - 1) * 128;
  } else {
    p = reinterpret_cast<const uint8_t*>(d.in0h) + (size_t)row * 256 + kb * 128;
  }
  return reinterpret_cast<const uint4*>(p + seg * 16);
}
struct ChunkRegs {
  uint4 v[8];
};
template <int MODE>
__device__ __forceinline__ void load_chunk_h(const fvgn_mlp_desc& d, int64_t row0, int kb, int pw, int lane, const TileIdx& idx,
                                             ChunkRegs& c) {
  const int seg = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = row0 + i * 16 + pw * 4 + (lane >> 3);
    int si = 0, ri = 0;
    if (MODE == FVGN_MLP_EDGE) {  // kb is a compile-time constant at every call site: one shuffle per gathered row
      if (kb < 2) si = idx.sender(i, lane);
      else if (kb < 4) ri = idx.receiver(i, lane);
    }
    const uint4* src = chunk_src_h<MODE>(d, row, kb, seg, si, ri);
    c.v[i] = src ? __ldg(src) : make_uint4(0u, 0u, 0u, 0u);
  }
}
__device__ __forceinline__ void store_chunk_h(uint8_t* stage, int pw, int lane, const ChunkRegs& c) {
  // row i*16 + r0 (r0 = pw*4 + lane/8 < 16): the swizzle only depends on r0, rows 16 apart are 2048 B apart
  uint8_t* base = stage + sw128_off(pw * 4 + (lane >> 3), lane & 7);
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(base + i * 2048) = c.v[i];
}
// modes whose layer-1 operands come from bf16 shadows (the encoders read their few fp32 input columns directly)
__host__ __device__ constexpr bool mode_has_shadow(int mode) {
  return mode == FVGN_MLP_EDGE || mode == FVGN_MLP_NODE || mode == FVGN_MLP_DEC;
}

// Producer loop of one CTA over its tiles (blockIdx.x + i * gridDim.x): chunk kb of a tile is held in register buffer
// kb % PF from the moment its loads are issued until `put(buffer, i, kb)` parks it in the ring, PF chunks later.  kb is a
// compile-time constant everywhere (NKB1 % PF == 0), so the endpoint-index registers are live only where needed.
template <int MODE>
__host__ __device__ constexpr int prefetch_depth() { return MODE == FVGN_MLP_DEC ? 2 : 3; }

// KB0 > 0: only chunks KB0 .. NKB1-1 of every tile are produced (kernel B of the node-level layer-1 path skips the gathered
// agg[s] | agg[r] chunks); `put` still receives the true chunk number.
template <int MODE, int NKB1, int KB0 = 0, class PutFn>
__device__ __forceinline__ void produce_tiles_h(const fvgn_mlp_desc& d, int64_t ntiles, int pw, int lane, PutFn&& put) {
  if constexpr (KB0 > 0) {
    constexpr int NC = NKB1 - KB0;   // chunks per tile, all of them plain row-major rows (no endpoint indices)
    ChunkRegs buf[NC];
    TileIdx none;
    none.s = none.r = 0;
    int64_t tile = blockIdx.x;
    if (tile >= ntiles) return;
#pragma unroll
    for (int c = 0; c < NC; ++c) load_chunk_h<MODE>(d, tile * TILE_M, KB0 + c, pw, lane, none, buf[c]);
    for (uint32_t i = 0; tile < ntiles; tile += gridDim.x, ++i) {
      const int64_t next = tile + gridDim.x;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        put(buf[c], i, KB0 + c);
        if (next < ntiles) load_chunk_h<MODE>(d, next * TILE_M, KB0 + c, pw, lane, none, buf[c]);
      }
    }
    return;
  }
  constexpr int PF = prefetch_depth<MODE>();
  static_assert(NKB1 % PF == 0, "chunks per tile must be a multiple of the prefetch depth");
  ChunkRegs buf[PF];
  TileIdx cur, nxt;
  int64_t tile = blockIdx.x;
  if (tile >= ntiles) return;
  load_tile_idx<MODE>(d, tile * TILE_M, pw, lane, cur);
  load_tile_idx<MODE>(d, (tile + gridDim.x) * TILE_M, pw, lane, nxt);  // rows past the end load nothing
#pragma unroll
  for (int kb = 0; kb < PF; ++kb) load_chunk_h<MODE>(d, tile * TILE_M, kb, pw, lane, cur, buf[kb]);
  for (uint32_t i = 0; tile < ntiles; tile += gridDim.x, ++i) {
    const int64_t next = tile + gridDim.x;
#pragma unroll
    for (int kb = 0; kb < NKB1; ++kb) {
      put(buf[kb % PF], i, kb);
      if (kb + PF < NKB1) {
        load_chunk_h<MODE>(d, tile * TILE_M, kb + PF, pw, lane, cur, buf[kb % PF]);
      } else if (next < ntiles) {
        load_chunk_h<MODE>(d, next * TILE_M, kb + PF - NKB1, pw, lane, nxt, buf[kb % PF]);
      }
    }
    cur = nxt;
    load_tile_idx<MODE>(d, (next + gridDim.x) * TILE_M, pw, lane, nxt);
  }
}

// ------------------------------------------------------------------------------------------ wide epilogue staging
// One 4 KB staging tile per epilogue warp: 32 rows x 128 B, 16-B chunk c of row r at r*128 + ((c ^ (r & 7)) << 4)
// (conflict-free for thread-per-row writes and for 8-lanes-per-row reads).  It turns the thread-per-row TMEM layout
// into 128-B row segments, i.e. 4 rows per fully coalesced warp instruction (4 L1 wavefronts instead of 16-32).
constexpr int WSTG_BYTES = 4096;
__device__ __forceinline__ uint8_t* wstg_at(uint8_t* stg, int r, int c) { return stg + r * 128 + ((c ^ (r & 7)) << 4); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------------------------------ extra helpers (backward)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// linear shared -> global bulk store (TMA engine), bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// named barrier among the 128 epilogue threads (barrier 0 is __syncthreads)
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Visit the eight 16-column chunks of a 128-column fp32 accumulator: f(c0, r[16]).  The tcgen05.ld of the next chunk
// is in flight while the current one is processed (tcgen05.wait::ld only waits once the work is done).
template <int NCOLS = 128, class F>
__device__ __forceinline__ void for_each_chunk16(uint32_t taddr, F&& f) {
  uint32_t ra[16], rb[16];
  tmem_ld16(taddr, ra);
  tmem_wait_ld();
#pragma unroll 1
  for (int c0 = 0; c0 < NCOLS; c0 += 32) {
    tmem_ld16(taddr + c0 + 16, rb);
    f(c0, ra);
    tmem_wait_ld();
    if (c0 + 32 < NCOLS) tmem_ld16(taddr + c0 + 32, ra);
    f(c0 + 16, rb);
    tmem_wait_ld();
  }
}


// ---- packed fp32x2 epilogue arithmetic (FVGN_F32X2 = 1, default): sm_100 executes fma / mul / add on register PAIRS
// (FFMA2 / FMUL2 / FADD2, IEEE per element), which halves the fp32 instruction count of the issue-bound epilogue stages.
// Every packed helper below performs, per element, exactly the operations of its scalar twin (bit-identical results).
#ifndef FVGN_F32X2
#define FVGN_F32X2 1
#endif
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 f2u(uint32_t lo, uint32_t hi) { return make_float2(__uint_as_float(lo), __uint_as_float(hi)); }
__device__ __forceinline__ float2 gelu_tanh2(float2 x) {
  const float2 u = __fmul2_rn(x, __ffma2_rn(splat2(0.0356774081f), __fmul2_rn(x, x), splat2(0.7978845608f)));
  const float2 hx = __fmul2_rn(splat2(0.5f), x);
  return __ffma2_rn(hx, make_float2(tanh_approx(u.x), tanh_approx(u.y)), hx);
}
// gelu_tanh_pair on two elements: (1 - t^2) and the polynomial are carried with flipped signs (exact), so no negations
__device__ __forceinline__ void gelu_tanh_pair2(float2 x, float2& h, float2& g) {
  const float2 x2 = __fmul2_rn(x, x);
  const float2 u = __fmul2_rn(x, __ffma2_rn(splat2(0.0356774081f), x2, splat2(0.7978845608f)));
  const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 hx = __fmul2_rn(splat2(0.5f), x);
  h = __ffma2_rn(hx, t, hx);
  const float2 q = __ffma2_rn(t, t, splat2(-1.0f));                                         // -(1 - t^2)
  const float2 pn = __ffma2_rn(splat2(-0.1070322243f), x2, splat2(-0.7978845608f));         // -(a + 3 b x^2)
  g = __ffma2_rn(__fmul2_rn(hx, q), pn, __ffma2_rn(splat2(0.5f), t, splat2(0.5f)));
}

// Column sums over the 32 lanes (rows) of a warp for 16 columns held per lane: returns, in lane L, the total of
// column (L & 15).  16 + 8 + 4 + 2 + 1 = 31 shuffles.
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
  float a[8];
  {
    const bool up = (lane >> 3) & 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float send = up ? v[j] : v[j + 8];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
      a[j] = (up ? v[j + 8] : v[j]) + recv;
    }
  }
  float b[4];
  {
    const bool up = (lane >> 2) & 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float send = up ? a[j] : a[j + 4];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
      b[j] = (up ? a[j + 4] : a[j]) + recv;
    }
  }
  float c[2];
  {
    const bool up = (lane >> 1) & 1;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float send = up ? b[j] : b[j + 2];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
      c[j] = (up ? b[j + 2] : b[j]) + recv;
    }
  }
  const bool up = lane & 1;
  const float send = up ? c[0] : c[1];
  const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
  return (up ? c[1] : c[0]) + recv;
}

}  // namespace tc
