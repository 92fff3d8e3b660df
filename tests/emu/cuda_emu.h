// TEST INFRASTRUCTURE ONLY.  A minimal CPU SIMT emulator so that the SIMT kernels under
// gen_fvgn_steady_b200/csrc can be compiled with g++ (-DFVGN_EMU) and their index math /
// numerics pre-verified against the oracle inside the GPU-less build container.
// The product package never loads the emulated library (it hard-fails without CUDA);
// only tests/ build and load it (tests/emu/build_emu.py).
//
// Execution model: blocks run one after another; the threads of a block are user-level fibers
// (ucontext) of the calling OS thread, scheduled round-robin: __syncthreads() / warp shuffles
// yield until every fiber of the block / warp has arrived (a generation-counting barrier), so a
// barrier costs a few context swaps instead of the futex storm of 256 OS threads on 8 cores;
// `__shared__` becomes a function-local static (safe because only one block is alive at a
// time).  Kernels launched through FVGN_LAUNCH_SEQ (no barrier / shuffle inside) are run as a
// plain loop in the caller.
#pragma once
#include <algorithm>
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <functional>
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
struct Fiber {
  ucontext_t ctx;
  std::vector<unsigned char> stack;
  dim3 tid;
  unsigned long long bid = 0;
  bool done = false;
};
struct Barrier {
  unsigned count = 0, gen = 0;
};
inline thread_local dim3 t_threadIdx, t_blockIdx;
inline thread_local unsigned t_linear = 0;
inline dim3 g_blockDim, g_gridDim;
inline std::vector<Fiber> g_fibers;
inline ucontext_t g_main_ctx;
inline Barrier g_block_bar;
inline std::vector<Barrier> g_warp_bar;
inline std::vector<uint64_t> g_warp_buf;   // [warp][32] shuffle exchange
inline std::vector<unsigned> g_warp_size;
inline std::vector<unsigned char> g_dyn;
inline unsigned char* g_dyn_ptr = nullptr;
inline bool g_in_sync_launch = false;
inline std::function<void()>* g_body = nullptr;
constexpr size_t FIBER_STACK = 512 * 1024;

inline void set_bid(unsigned long long b) {
  t_blockIdx.x = (unsigned)(b % g_gridDim.x);
  t_blockIdx.y = (unsigned)((b / g_gridDim.x) % g_gridDim.y);
  t_blockIdx.z = (unsigned)(b / ((unsigned long long)g_gridDim.x * g_gridDim.y));
}
// switch to the next unfinished fiber (round robin); returns when this fiber is resumed
inline void yield() {
  const unsigned me = t_linear, T = (unsigned)g_fibers.size();
  unsigned nxt = me;
  do { nxt = (nxt + 1) % T; } while (g_fibers[nxt].done && nxt != me);
  if (nxt == me) return;
  swapcontext(&g_fibers[me].ctx, &g_fibers[nxt].ctx);
  t_linear = me;   // the thread_local "registers" are shared by the fibers: restore ours
  t_threadIdx = g_fibers[me].tid;
  set_bid(g_fibers[me].bid);
}
inline void barrier_wait(Barrier& b, unsigned parties) {
  const unsigned my_gen = b.gen;
  if (++b.count == parties) {
    b.count = 0;
    ++b.gen;
    return;
  }
  while (b.gen == my_gen) yield();
}
inline void fiber_main(unsigned t) {
  Fiber& f = g_fibers[t];
  t_linear = t;
  t_threadIdx = f.tid;
  const unsigned long long nb = (unsigned long long)g_gridDim.x * g_gridDim.y * g_gridDim.z;
  const unsigned T = (unsigned)g_fibers.size();
  for (unsigned long long b = 0; b < nb; ++b) {
    f.bid = b;
    set_bid(b);
    (*g_body)();
    barrier_wait(g_block_bar, T);   // one block alive at a time (function-local static "shared memory")
  }
  f.done = true;
  // hand over to any unfinished fiber, else back to the launcher
  for (unsigned i = 1; i <= T; ++i) {
    const unsigned nxt = (t + i) % T;
    if (!g_fibers[nxt].done) {
      setcontext(&g_fibers[nxt].ctx);
    }
  }
  setcontext(&g_main_ctx);
}
inline void fiber_entry(int t) { fiber_main((unsigned)t); }

template <class F>
void launch(bool sync, dim3 grid, dim3 block, size_t smem, F&& body) {
  g_blockDim = block;
  g_gridDim = grid;
  if (g_dyn.size() < smem + 1024) g_dyn.resize(smem + 1024);
  g_dyn_ptr = (unsigned char*)(((uintptr_t)g_dyn.data() + 1023) & ~(uintptr_t)1023);
  const unsigned T = block.x * block.y * block.z;
  const unsigned long long nb = (unsigned long long)grid.x * grid.y * grid.z;
  auto tid_of = [&](unsigned t) { return dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y)); };
  if (!sync) {
    g_in_sync_launch = false;
    for (unsigned long long b = 0; b < nb; ++b) {
      set_bid(b);
      for (unsigned t = 0; t < T; ++t) {
        t_linear = t;
        t_threadIdx = tid_of(t);
        body();
      }
    }
    return;
  }
  if (nb == 0 || T == 0) return;
  g_in_sync_launch = true;
  std::function<void()> fn = [&] { body(); };
  g_body = &fn;
  g_block_bar = Barrier();
  const unsigned nw = (T + 31) / 32;
  g_warp_bar.assign(nw, Barrier());
  g_warp_buf.assign((size_t)nw * 32, 0);
  g_warp_size.resize(nw);
  for (unsigned w = 0; w < nw; ++w) g_warp_size[w] = std::min(32u, T - w * 32);
  if (g_fibers.size() != T) g_fibers.resize(T);
  for (unsigned t = 0; t < T; ++t) {
    Fiber& f = g_fibers[t];
    if (f.stack.size() != FIBER_STACK) f.stack.resize(FIBER_STACK);
    f.tid = tid_of(t);
    f.bid = 0;
    f.done = false;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.data();
    f.ctx.uc_stack.ss_size = f.stack.size();
    f.ctx.uc_link = &g_main_ctx;
    makecontext(&f.ctx, (void (*)())fiber_entry, 1, (int)t);
  }
  swapcontext(&g_main_ctx, &g_fibers[0].ctx);   // returns when the last fiber has finished
  g_body = nullptr;
  g_in_sync_launch = false;
}

inline void syncthreads() {
  if (!g_in_sync_launch) {
    fprintf(stderr, "emu: __syncthreads() inside a FVGN_LAUNCH_SEQ kernel\n");
    abort();
  }
  barrier_wait(g_block_bar, (unsigned)g_fibers.size());
}
template <class T>
T shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  if (!g_in_sync_launch) {
    fprintf(stderr, "emu: warp shuffle inside a FVGN_LAUNCH_SEQ kernel\n");
    abort();
  }
  const unsigned w = t_linear / 32;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  g_warp_buf[(size_t)w * 32 + t_linear % 32] = bits;
  barrier_wait(g_warp_bar[w], g_warp_size[w]);
  uint64_t got = g_warp_buf[(size_t)w * 32 + (src_lane & 31)];
  barrier_wait(g_warp_bar[w], g_warp_size[w]);
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}
inline void syncwarp() {
  if (!g_in_sync_launch) {
    fprintf(stderr, "emu: warp barrier inside a FVGN_LAUNCH_SEQ kernel\n");
    abort();
  }
  const unsigned w = t_linear / 32;
  barrier_wait(g_warp_bar[w], g_warp_size[w]);
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
