// TEST INFRASTRUCTURE ONLY.  A minimal CPU emulation of the CUDA launch model, just enough to
// execute the kernels and the host orchestration of csrc/backward.cu without a GPU: every
// thread of a block is a real host thread, __syncthreads() is a barrier over the block,
// __shfl_xor_sync exchanges through a per-warp slot array, blocks run one after another.  The
// tensor-core GEMM (gemm_launch) is replaced by a plain loop with the same argument contract.
// Only tests/ compiles this (tests/emu/build_emu.py -> tests/emu/_build/libstat_bw_emu.so);
// the product library never sees it, and nothing in the package can load the emulated library.
#pragma once

#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/stat_b200.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static   // blocks run sequentially: one instance per kernel = block-shared

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void *cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) {
  memcpy(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height,
                                     cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; ++r)
    memcpy(static_cast<char *>(d) + r * dpitch, static_cast<const char *>(s) + r * spitch, width);
  return cudaSuccess;
}

namespace emu {
struct Warp {
  float slot[32];
  std::unique_ptr<std::barrier<>> bar;
};
struct Block {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<Warp> warps;
};
inline thread_local Block *cur_block = nullptr;
inline thread_local int cur_tid = 0;
}  // namespace emu

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

inline void __syncthreads() { emu::cur_block->bar->arrive_and_wait(); }

inline float __shfl_xor_sync(unsigned, float v, int o) {
  emu::Warp &w = emu::cur_block->warps[emu::cur_tid >> 5];
  const int lane = emu::cur_tid & 31;
  w.slot[lane] = v;
  w.bar->arrive_and_wait();
  const float r = w.slot[lane ^ o];
  w.bar->arrive_and_wait();
  return r;
}

using std::max;
using std::min;

namespace emu {
// kernel<<<grid, block>>>(args...): block sizes must be multiples of 32 (full warps)
template <typename... KArgs, typename... Args>
void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, Args... args) {
  const int nt = static_cast<int>(block.x * block.y * block.z);
  if (nt % 32 != 0) {
    fprintf(stderr, "emu: block size %d is not a multiple of 32\n", nt);
    abort();
  }
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block blk;
        blk.bar.reset(new std::barrier<>(nt));
        blk.warps.resize(nt / 32);
        for (auto &w : blk.warps) w.bar.reset(new std::barrier<>(32));
        std::vector<std::thread> th;
        th.reserve(nt);
        for (int t = 0; t < nt; ++t) {
          th.emplace_back([&, t]() {
            cur_block = &blk;
            cur_tid = t;
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            blockIdx = dim3(bx, by, bz);
            blockDim = block;
            gridDim = grid;
            kernel(KArgs(args)...);
          });
        }
        for (auto &t : th) t.join();
      }
}
}  // namespace emu

// ---- the pieces of stat_common.cuh the backward pass uses (contracts must stay identical) ----
namespace stat {

inline thread_local char g_emu_err[512];
inline void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_emu_err, sizeof(g_emu_err), fmt, ap);
  va_end(ap);
}
inline void note_launch() {}

#define STAT_CUDA_CHECK(expr)                                     \
  do {                                                            \
    cudaError_t e__ = (expr);                                     \
    if (e__ != cudaSuccess) {                                     \
      stat::set_error("%s:%d %s failed", __FILE__, __LINE__, #expr); \
      return STAT_ECUDA;                                          \
    }                                                             \
  } while (0)
#define STAT_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      stat::set_error(__VA_ARGS__);   \
      return (code);                  \
    }                                 \
  } while (0)
#define STAT_TRY(expr)                 \
  do {                                 \
    int r__ = (expr);                  \
    if (r__ != STAT_OK) return r__;    \
  } while (0)

struct GemmSeg {
  float *C;
  int ldc;
  const float *bias;
  const float *addend;
  int ld_add;
  float alpha;
  float post;
  int act;
  int f0, f1;
};
struct GemmArgs {
  const float *P;
  int ldp;
  int NP;
  const float *Q;
  int ldq;
  int NQ;
  int K;
  int feat_on_p;
  int nseg;
  GemmSeg seg[2];
  int ksplit;
  size_t plane;
};

// out = post * act(alpha * P.Q^T + bias + addend), routing exactly as documented in stat_common.cuh; k-split:
// slice z of the 32-wide k-blocks (the kernels' partition) goes to plane z, bias / addend ride on plane 0
inline int gemm_launch(const GemmArgs &a, cudaStream_t) {
  STAT_REQUIRE(a.NP > 0 && a.NQ > 0 && a.K > 0, STAT_EINVAL, "gemm: empty problem");
  const int nk_all = (a.K + 31) / 32;
  const int ks = a.ksplit < 1 ? 1 : (a.ksplit > nk_all ? nk_all : a.ksplit);
  if (ks > 1) STAT_REQUIRE(a.nseg == 1 && a.seg[0].act == 0, STAT_EINVAL, "gemm: k-split needs a single linear segment");
  for (int kz = 0; kz < ks; ++kz) {
    const int kbeg = ((kz * nk_all) / ks) * 32;
    const int kend = std::min(a.K, (((kz + 1) * nk_all) / ks) * 32);
    for (int i = 0; i < a.NP; ++i)
      for (int j = 0; j < a.NQ; ++j) {
        float acc = 0.f;
        for (int k = kbeg; k < kend; ++k)
          acc = fmaf(a.P[static_cast<size_t>(i) * a.ldp + k], a.Q[static_cast<size_t>(j) * a.ldq + k], acc);
        const int feat = a.feat_on_p ? i : j, row = a.feat_on_p ? j : i;
        const GemmSeg &s = a.seg[(a.nseg > 1 && feat >= a.seg[1].f0) ? 1 : 0];
        if (feat < s.f0 || feat >= s.f1) continue;
        const int f = feat - s.f0;
        float x = s.alpha * acc + ((s.bias && kz == 0) ? s.bias[f] : 0.f);
        if (s.addend && kz == 0) x += s.addend[static_cast<size_t>(row) * s.ld_add + f];
        if (s.act) x = tanhf(x);
        s.C[static_cast<size_t>(kz) * a.plane + static_cast<size_t>(row) * s.ldc + f] = x * s.post;
      }
  }
  return STAT_OK;
}

}  // namespace stat
