// Shared definitions for the fvgn_b200 kernels (sm_100a).  See include/fvgn_b200.h for the C-ABI.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef FVGN_EMU
// tests/emu/cuda_emu.h : CPU SIMT emulator, test infrastructure only (never shipped).
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define FVGN_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#define FVGN_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<grid, block, smem, (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define FVGN_LAUNCH_SEQ FVGN_LAUNCH
#endif

#include "../../include/fvgn_b200.h"

#define FVGN_H 128  // hidden width of every latent (EPD.py: hidden_size=128)

// NodeType (reference utils/utilities.py:7-13)
enum { NT_NORMAL = 0, NT_INFLOW = 1, NT_OUTFLOW = 2, NT_WALL = 3, NT_PRESS_POINT = 4, NT_IN_WALL = 5 };

#define FVGN_CHECK_LAUNCH()                                   \
  do {                                                        \
    if (cudaGetLastError() != cudaSuccess) return FVGN_ERR_LAUNCH; \
  } while (0)

static inline int fvgn_aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

#ifndef FVGN_EMU
// per-device caches (one process may drive several GPUs): current device index and its SM count
#define FVGN_MAX_DEV 64
static inline int fvgn_cur_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & (FVGN_MAX_DEV - 1);
}
static inline int fvgn_num_sms() {
  static int n[FVGN_MAX_DEV] = {0};
  const int dev = fvgn_cur_device();
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}
#endif

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// exact GELU (nn.GELU default, EPD.py:24-30) and its derivative
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
