// Small streaming kernels around the network: deterministic per-graph column statistics, the
// importer prologue (feature normalisation) and the head (tanh clamp, Dirichlet BC, phi assembly).
#include "common.cuh"

// ----------------------------------------------------------------------------------------------
// out_partial[chunk, c] = sum_{rows of chunk} (x[row, c] - center[seg, c])^power,  c < width <= 16.
// One CTA per chunk; per-thread strided partial sums, then a fixed-shape shared-memory tree.
__global__ void __launch_bounds__(256) chunk_colsum_kernel(const float* __restrict__ x, int width, int ld,
                                                           const float* __restrict__ center, int center_ld, int power,
                                                           const int32_t* __restrict__ chunks, float* __restrict__ out) {
  __shared__ float red[256 * 16];
  const int tid = threadIdx.x;
  const int seg = chunks[blockIdx.x * 3 + 0], r0 = chunks[blockIdx.x * 3 + 1], r1 = chunks[blockIdx.x * 3 + 2];
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  for (int r = r0 + tid; r < r1; r += 256) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      if (c < width) {
        float v = x[(size_t)r * ld + c];
        if (center) v -= center[seg * center_ld + c];
        acc[c] += (power == 2) ? v * v : v;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) red[tid * 16 + c] = acc[c];
  __syncthreads();
  for (int s = 128; s >= 1; s >>= 1) {
    if (tid < s) {
#pragma unroll
      for (int c = 0; c < 16; ++c) red[tid * 16 + c] += red[(tid + s) * 16 + c];
    }
    __syncthreads();
  }
  if (tid < width) out[(size_t)blockIdx.x * width + tid] = red[tid];
}

__global__ void chunk_combine_kernel(const float* __restrict__ partial, int width, const int32_t* __restrict__ chunk_ptr,
                                     int nseg, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseg * width) return;
  const int seg = i / width, c = i % width;
  // fixed summation shape (four interleaved chains + tail): deterministic, and the loads of a long chunk list overlap
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int k = chunk_ptr[seg];
  const int end = chunk_ptr[seg + 1];
  for (; k + 4 <= end; k += 4) {
    s0 += partial[(size_t)k * width + c];
    s1 += partial[(size_t)(k + 1) * width + c];
    s2 += partial[(size_t)(k + 2) * width + c];
    s3 += partial[(size_t)(k + 3) * width + c];
  }
  for (; k < end; ++k) s0 += partial[(size_t)k * width + c];
  out[i] = (s0 + s1) + (s2 + s3);
}

extern "C" int fvgn_chunk_colsum(const float* x, int32_t width, int32_t ld, const float* center, int32_t center_ld,
                                 int32_t power, const int32_t* chunks, int32_t nchunks, float* out_partial, void* stream) {
  if (width < 1 || width > 16 || (power != 1 && power != 2)) return FVGN_ERR_UNSUPPORTED;
  if (nchunks <= 0) return FVGN_OK;
  FVGN_LAUNCH(chunk_colsum_kernel, (unsigned)nchunks, 256, 0, stream, x, width, ld, center, center_ld, power, chunks,
              out_partial);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_chunk_combine(const float* partial, int32_t width, const int32_t* chunk_ptr, int32_t nseg, float* out,
                                  void* stream) {
  if (nseg <= 0) return FVGN_OK;
  FVGN_LAUNCH_SEQ(chunk_combine_kernel, (unsigned)((nseg * width + 127) / 128), 128, 0, stream, partial, width, chunk_ptr,
                  nseg, out);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

// ----------------------------------------------------------------------------------------------
// importer.py:168-176 (update_x_attr + uv_old)
__global__ void prologue_kernel(const float* __restrict__ x, const int32_t* __restrict__ batch,
                                const float* __restrict__ uvp_dim, const float* __restrict__ gmean,
                                const float* __restrict__ gstd, const float* __restrict__ nmean,
                                const float* __restrict__ nstd, int norm_uvp, float* __restrict__ xn,
                                float* __restrict__ uv_old, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = batch[i];
  const float* xi = x + (size_t)i * 12;
  float* o = xn + (size_t)i * 12;
  uv_old[i * 2 + 0] = xi[0] / uvp_dim[b * 3 + 0];
  uv_old[i * 2 + 1] = xi[1] / uvp_dim[b * 3 + 1];
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = norm_uvp ? (xi[c] - gmean[b * 3 + c]) / (gstd[b * 3 + c] + 1e-8f) : xi[c];
#pragma unroll
  for (int c = 0; c < 9; ++c) o[3 + c] = nmean ? (xi[3 + c] - nmean[c]) / nstd[c] : xi[3 + c];
}

extern "C" int fvgn_prologue(const float* x, const int32_t* batch, const float* uvp_dim, const float* gmean,
                             const float* gstd, const float* nmean, const float* nstd, int32_t norm_uvp, float* xn,
                             float* uv_old, int64_t n, void* stream) {
  if (n <= 0) return FVGN_OK;
  FVGN_LAUNCH_SEQ(prologue_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, x, batch, uvp_dim, gmean, gstd, nmean,
                  nstd, norm_uvp, xn, uv_old, n);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

// ----------------------------------------------------------------------------------------------
// importer.py:187-201 + FVscheme.py:643-646
__device__ __forceinline__ bool is_dirichlet(int t) {
  return t == NT_WALL || t == NT_INFLOW || t == NT_PRESS_POINT || t == NT_IN_WALL;
}

__global__ void head_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ uv_old, const float* __restrict__ y,
                                const int32_t* __restrict__ node_type, int integrator, float* __restrict__ phi, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = node_type[i];
  float u = tanhf(raw[i * 3 + 0] / 10.f) * 10.f;
  float v = tanhf(raw[i * 3 + 1] / 10.f) * 10.f;
  float p = tanhf(raw[i * 3 + 2] / 10.f) * 10.f;
  if (is_dirichlet(t)) { u = y[i * 2]; v = y[i * 2 + 1]; }
  if (t == NT_PRESS_POINT) p = 0.f;
  const float uo = uv_old[i * 2], vo = uv_old[i * 2 + 1];
  float uh, vh;
  if (integrator == 0) { uh = uo; vh = vo; }
  else if (integrator == 1) { uh = u; vh = v; }
  else { uh = (uo + u) / 2.0f; vh = (vo + v) / 2.0f; }
  float* o = phi + (size_t)i * 7;
  o[0] = u; o[1] = v; o[2] = p; o[3] = uh; o[4] = vh; o[5] = uo; o[6] = vo;
}

__global__ void head_bwd_kernel(const float* __restrict__ raw, const int32_t* __restrict__ node_type, int integrator,
                                const float* __restrict__ d_phi, float* __restrict__ d_raw, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = node_type[i];
  const float* g = d_phi + (size_t)i * 7;
  const float wh = (integrator == 0) ? 0.f : (integrator == 1) ? 1.f : 0.5f;
  float du = g[0] + wh * g[3], dv = g[1] + wh * g[4], dp = g[2];
  if (is_dirichlet(t)) { du = 0.f; dv = 0.f; }
  if (t == NT_PRESS_POINT) dp = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float th = tanhf(raw[i * 3 + c] / 10.f);
    const float gg = (c == 0) ? du : (c == 1) ? dv : dp;
    d_raw[i * 3 + c] = gg * (1.0f - th * th);
  }
}

extern "C" int fvgn_head_forward(const float* raw, const float* uv_old, const float* y, const int32_t* node_type,
                                 int32_t integrator, float* phi, int64_t n, void* stream) {
  if (n <= 0) return FVGN_OK;
  if (integrator < 0 || integrator > 2) return FVGN_ERR_UNSUPPORTED;
  FVGN_LAUNCH_SEQ(head_fwd_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, raw, uv_old, y, node_type, integrator, phi,
                  n);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_head_backward(const float* raw, const int32_t* node_type, int32_t integrator, const float* d_phi,
                                  float* d_raw, int64_t n, void* stream) {
  if (n <= 0) return FVGN_OK;
  if (integrator < 0 || integrator > 2) return FVGN_ERR_UNSUPPORTED;
  FVGN_LAUNCH_SEQ(head_bwd_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, raw, node_type, integrator, d_phi, d_raw,
                  n);
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}
