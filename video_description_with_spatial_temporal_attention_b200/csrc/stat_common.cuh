// Shared device helpers for the STAT decoder kernels (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stat_b200.h"

namespace stat {

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
// every kernel launch of the library passes through here (stat_launch_count)
void note_launch();
void set_launch_label(const char *label);      // STAT_SYNC_DEBUG: names the next launch in the fault report

// Programmatic dependent launch (STAT_PDL=0 switches it off): fills attr[0] and returns the number
// of launch attributes (0 or 1).  Kernels launched this way call pdl_wait() before the first read of
// anything the previous kernel wrote, and pdl_trigger() after it (never before: a kernel whose
// own dependency is still pending must not let its successor start).
int pdl_attr(cudaLaunchAttribute *attr);

#define STAT_CUDA_CHECK(expr)                                                      \
  do {                                                                             \
    cudaError_t e__ = (expr);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      stat::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,                 \
                      cudaGetErrorString(e__));                                    \
      return STAT_ECUDA;                                                           \
    }                                                                              \
  } while (0)

#define STAT_REQUIRE(cond, code, ...)                                              \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      stat::set_error(__VA_ARGS__);                                                \
      return (code);                                                               \
    }                                                                              \
  } while (0)

#define STAT_TRY(expr)                                                             \
  do {                                                                             \
    int r__ = (expr);                                                              \
    if (r__ != STAT_OK) return r__;                                                \
  } while (0)

// ---------------------------------------------------------------------------
// math.  The attention loop evaluates ~9.4 M tanh per decode step, so the hot
// variant is two MUFU ops (ex2, rcp) plus three FMA-pipe ops; its absolute
// error is ~1.5e-7, far inside the 1e-5 tolerance on attention weights.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// tanh(x) = 1 - 2/(1+e^{2x});  e^{2x} = 2^{x * 2*log2(e)}
__device__ __forceinline__ float tanh_fast(float x) {
  float t = ex2_approx(x * 2.885390081777927f);
  return fmaf(-2.0f, rcp_approx(t + 1.0f), 1.0f);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: raise it once per (kernel
// instantiation, device) and only when a launch needs more than was configured before.  `table` is a
// zero-initialised static array of MAX_DEV entries owned by the caller (one per instantiation).
constexpr int STAT_MAX_DEV = 64;
template <typename K>
int ensure_dyn_smem(K kernel, size_t bytes, size_t *table) {
  int dev = 0;
  STAT_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= STAT_MAX_DEV) dev = 0;
  if (bytes > table[dev]) {
    STAT_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    table[dev] = bytes;
  }
  return STAT_OK;
}

// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-serialization attribute
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr(attr);
  STAT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
  note_launch();
  return STAT_OK;
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// accurate variants for the few elementwise sites outside the attention loop
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------
// dense primitive (gemm_tf32x3.cu)
// ---------------------------------------------------------------------------
// out[i][j] = post * act(alpha * dot(P[i,:], Q[j,:]) + bias + addend)
// P (NP,K) and Q (NQ,K) are row-major (K-major operands).  P rides the 128-lane
// axis of the tensor core tile, Q the column axis.
//   feat_on_p == 0 ("normal"):  C[i*ldc + j], bias[j], features = Q index
//   feat_on_p == 1 ("swap")  :  C[j*ldc + i], bias[i], features = P index
// Up to two feature segments (boundaries multiples of 128) may route to
// different outputs with different epilogues.
struct GemmSeg {
  float *C;
  int ldc;
  const float *bias;     // per feature, indexed from the segment start; or null
  const float *addend;   // same indexing as C; or null
  int ld_add;
  float alpha;
  float post;
  int act;               // 0 none, 1 tanh
  int f0, f1;            // feature range [f0,f1) of this segment
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
  // k-split: slice z of the K range accumulates into output plane z (C + z*plane floats);
  // plane 0 carries bias/addend, the consumer sums the planes in order.  0/1 = off.
  int ksplit;
  size_t plane;
};

int gemm_launch(const GemmArgs &a, cudaStream_t stream);
// 2-D tensor map of a K-major fp32 operand (rows, K), row pitch ld floats (multiple of 4, base 16-byte aligned):
// box = box_rows x 32 floats (one 128-byte swizzle row), SWIZZLE_128B, out-of-range elements read as 0
int make_tensor_map(CUtensorMap *tm, const float *base, int rows, int K, int ld, int box_rows);
void gemm_set_trace(long long *p);
long long *gemm_get_trace();
void gemm_set_impl(int impl);
int gemm_get_impl();

}  // namespace stat
