// TEST INFRASTRUCTURE ONLY.  A minimal CPU SIMT emulator so that the SIMT kernels under
// gen_fvgn_steady_b200/csrc can be compiled with g++ (-DFVGN_EMU) and their index math /
// numerics pre-verified against the oracle inside the GPU-less build container.
// The product package never loads the emulated library (it hard-fails without CUDA);
// only tests/ build and load it (tests/emu/build_emu.py).
//
// Execution model: blocks run one after another; the threads of a block are OS threads that
// meet at a std::barrier for __syncthreads(); `__shared__` becomes a function-local static
// (safe because only one block is alive at a time).  Kernels launched through
// FVGN_LAUNCH_SEQ (no barrier / shuffle inside) are run as a plain loop in the caller.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct double2 { double x, y; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }

namespace emu {
struct WarpCtx {
  uint64_t buf[32];
  std::unique_ptr<std::barrier<>> bar;
};
inline thread_local dim3 t_threadIdx, t_blockIdx;
inline thread_local unsigned t_linear = 0;
inline dim3 g_blockDim, g_gridDim;
inline std::barrier<>* g_bar = nullptr;
inline std::vector<WarpCtx> g_warps;
inline std::vector<unsigned char> g_dyn;
inline unsigned char* g_dyn_ptr = nullptr;
inline bool g_in_sync_launch = false;

template <class F>
void launch(bool sync, dim3 grid, dim3 block, size_t smem, F&& body) {
  g_blockDim = block;
  g_gridDim = grid;
  if (g_dyn.size() < smem + 1024) g_dyn.resize(smem + 1024);
  g_dyn_ptr = (unsigned char*)(((uintptr_t)g_dyn.data() + 1023) & ~(uintptr_t)1023);
  const unsigned T = block.x * block.y * block.z;
  const unsigned long long nb = (unsigned long long)grid.x * grid.y * grid.z;
  auto set_tid = [&](unsigned t) {
    t_linear = t;
    t_threadIdx.x = t % block.x;
    t_threadIdx.y = (t / block.x) % block.y;
    t_threadIdx.z = t / (block.x * block.y);
  };
  auto set_bid = [&](unsigned long long b) {
    t_blockIdx.x = (unsigned)(b % grid.x);
    t_blockIdx.y = (unsigned)((b / grid.x) % grid.y);
    t_blockIdx.z = (unsigned)(b / ((unsigned long long)grid.x * grid.y));
  };
  if (!sync) {
    g_in_sync_launch = false;
    for (unsigned long long b = 0; b < nb; ++b) {
      set_bid(b);
      for (unsigned t = 0; t < T; ++t) {
        set_tid(t);
        body();
      }
    }
    return;
  }
  g_in_sync_launch = true;
  std::barrier<> bar(T);
  g_bar = &bar;
  const unsigned nw = (T + 31) / 32;
  g_warps.clear();
  g_warps.resize(nw);
  for (unsigned w = 0; w < nw; ++w) {
    unsigned cnt = std::min(32u, T - w * 32);
    g_warps[w].bar = std::make_unique<std::barrier<>>(cnt);
  }
  std::vector<std::thread> th;
  th.reserve(T);
  for (unsigned t = 0; t < T; ++t) {
    th.emplace_back([&, t] {
      set_tid(t);
      for (unsigned long long b = 0; b < nb; ++b) {
        set_bid(b);
        body();
        bar.arrive_and_wait();
      }
    });
  }
  for (auto& x : th) x.join();
  g_bar = nullptr;
  g_in_sync_launch = false;
}

inline void syncthreads() {
  if (!g_in_sync_launch) {
    fprintf(stderr, "emu: __syncthreads() inside a FVGN_LAUNCH_SEQ kernel\n");
    abort();
  }
  g_bar->arrive_and_wait();
}
template <class T>
T shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  if (!g_in_sync_launch) {
    fprintf(stderr, "emu: warp shuffle inside a FVGN_LAUNCH_SEQ kernel\n");
    abort();
  }
  WarpCtx& w = g_warps[t_linear / 32];
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  w.buf[t_linear % 32] = bits;
  w.bar->arrive_and_wait();
  uint64_t got = w.buf[src_lane & 31];
  w.bar->arrive_and_wait();
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}
inline void syncwarp() {
  if (!g_in_sync_launch) {
    fprintf(stderr, "emu: warp barrier inside a FVGN_LAUNCH_SEQ kernel\n");
    abort();
  }
  g_warps[t_linear / 32].bar->arrive_and_wait();
}
}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim
#define __syncthreads() emu::syncthreads()
#define __syncwarp(...) ((void)0)
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl(v, (int)(emu::t_linear % 32) ^ m); }
template <class T> static inline T __shfl_sync(unsigned, T v, int l) { return emu::shfl(v, l); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  int l = (int)(emu::t_linear % 32) + d;
  return emu::shfl(v, l > 31 ? (int)(emu::t_linear % 32) : l);
}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
using std::max;
using std::min;

#define FVGN_DYN_SMEM(name) unsigned char* name = emu::g_dyn_ptr
#define FVGN_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch(true, dim3(grid), dim3(block), (size_t)(smem), [&] { kern(__VA_ARGS__); })
#define FVGN_LAUNCH_SEQ(kern, grid, block, smem, stream, ...) \
  emu::launch(false, dim3(grid), dim3(block), (size_t)(smem), [&] { kern(__VA_ARGS__); })
