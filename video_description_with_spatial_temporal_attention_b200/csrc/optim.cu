// Parameter update of the training step (second half): global-norm gradient clipping
// (model_attention.py:1194-1203) and the reference's optimizers (common.py:178-230) over ONE flat fp32
// buffer that holds all parameters in init_params order -- the layout the data-parallel gradient
// all-reduce uses too.  HBM-bound elementwise kernels: float4 accesses, grids of a multiple of the SM
// count, every byte touched once per step (adam: 16 B read + 12 B written per parameter).
#include <math.h>

#include "kernels.cuh"
#include "stat_common.cuh"

namespace stat {
namespace {

constexpr int OPT_THREADS = 256;
constexpr int OPT_BLOCKS = 148 * 4;        // fixed grid: the partial sums below are reproducible

// ---- sum of squares, two fixed-shape stages (deterministic) -------------------------------------
__global__ void __launch_bounds__(OPT_THREADS) sumsq_partial_kernel(const float *g, size_t n, double *partial) {
  double acc = 0.0;
  const size_t n4 = n >> 2;
  const float4 *g4 = reinterpret_cast<const float4 *>(g);
  for (size_t i = blockIdx.x * static_cast<size_t>(OPT_THREADS) + threadIdx.x; i < n4;
       i += static_cast<size_t>(OPT_BLOCKS) * OPT_THREADS) {
    const float4 x = g4[i];
    acc += static_cast<double>(x.x) * x.x + static_cast<double>(x.y) * x.y + static_cast<double>(x.z) * x.z +
           static_cast<double>(x.w) * x.w;
  }
  if (blockIdx.x == 0)
    for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) acc += static_cast<double>(g[i]) * g[i];
  __shared__ double s[OPT_THREADS];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

// g2 = sum of the partials; factor = clip_c / sqrt(g2) when g2 > clip_c^2, else 1  (:1199-1201)
__global__ void __launch_bounds__(OPT_THREADS) clip_factor_kernel(const double *partial, float clip_c, float *out) {
  __shared__ double s[OPT_THREADS];
  double acc = 0.0;
  for (int i = threadIdx.x; i < OPT_BLOCKS; i += OPT_THREADS) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float g2 = static_cast<float>(s[0]);
    out[0] = g2;
    out[1] = (clip_c > 0.f && g2 > clip_c * clip_c) ? clip_c / sqrtf(g2) : 1.0f;
  }
}

__global__ void __launch_bounds__(OPT_THREADS) scale_by_kernel(float *g, size_t n, const float *factor) {
  const float f = factor[1];
  if (f == 1.0f) return;
  const size_t n4 = n >> 2;
  float4 *g4 = reinterpret_cast<float4 *>(g);
  for (size_t i = blockIdx.x * static_cast<size_t>(OPT_THREADS) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * OPT_THREADS) {
    float4 x = g4[i];
    x.x *= f; x.y *= f; x.z *= f; x.w *= f;
    g4[i] = x;
  }
  if (blockIdx.x == 0)
    for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) g[i] *= f;
}

// ---- adam, exactly as common.py:197-230: lr0 = 2e-4 (the lr argument is ignored there), b1 = 0.1 and
// b2 = 0.001 are the weights of the NEW gradient, bias correction lr_t = lr0 sqrt(1 - b2^t) / (1 - b1^t)
__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, float lr_t) {
  const float b1 = 0.1f, b2 = 0.001f, e = 1e-8f;
  m = (b1 * g) + ((1.0f - b1) * m);
  v = (b2 * (g * g)) + ((1.0f - b2) * v);
  p = p - (lr_t * (m / (sqrtf(v) + e)));
}
__global__ void __launch_bounds__(OPT_THREADS) adam_kernel(float *p, const float *g, float *m, float *v, size_t n,
                                                          float lr_t) {
  const size_t n4 = n >> 2;
  float4 *p4 = reinterpret_cast<float4 *>(p), *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
  const float4 *g4 = reinterpret_cast<const float4 *>(g);
  for (size_t i = blockIdx.x * static_cast<size_t>(OPT_THREADS) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * OPT_THREADS) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, lr_t);
    adam_one(pp.y, gg.y, mm.y, vv.y, lr_t);
    adam_one(pp.z, gg.z, mm.z, vv.z, lr_t);
    adam_one(pp.w, gg.w, mm.w, vv.w, lr_t);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  if (blockIdx.x == 0)
    for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) adam_one(p[i], g[i], m[i], v[i], lr_t);
}

// ---- adadelta (common.py:178-195): rg2 <- 0.95 rg2 + 0.05 g^2 belongs to f_grad_shared (phase 0),
// the step itself to f_update (phase 1): ud = -sqrt(ru2 + 1e-6) / sqrt(rg2 + 1e-6) g, ru2 <- 0.95 ru2 +
// 0.05 ud^2, p <- p + ud
__global__ void __launch_bounds__(OPT_THREADS) adadelta_kernel(float *p, const float *g, float *rg2, float *ru2,
                                                              size_t n, int phase) {
  for (size_t i = blockIdx.x * static_cast<size_t>(OPT_THREADS) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * OPT_THREADS) {
    const float gg = g[i];
    if (phase == 0) {
      rg2[i] = 0.95f * rg2[i] + 0.05f * (gg * gg);
    } else {
      const float ud = -sqrtf(ru2[i] + 1e-6f) / sqrtf(rg2[i] + 1e-6f) * gg;
      ru2[i] = 0.95f * ru2[i] + 0.05f * (ud * ud);
      p[i] = p[i] + ud;
    }
  }
}

// ---- attention-coverage regulariser (model_attention.py:1138-1147): for alphas (L, rows, n),
// ((1 - alphas.sum(0))**2).sum(0).mean() = (1/n) sum_{b,j} (1 - sum_l alphas[l,b,j])^2; same two
// fixed-shape stages as the sum of squares
__global__ void __launch_bounds__(OPT_THREADS) coverage_partial_kernel(const float *alphas, int L, size_t rows_n,
                                                                      double *partial) {
  double acc = 0.0;
  for (size_t i = blockIdx.x * static_cast<size_t>(OPT_THREADS) + threadIdx.x; i < rows_n;
       i += static_cast<size_t>(OPT_BLOCKS) * OPT_THREADS) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += alphas[static_cast<size_t>(l) * rows_n + i];
    const double dlt = 1.0 - static_cast<double>(s);
    acc += dlt * dlt;
  }
  __shared__ double sm[OPT_THREADS];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(OPT_THREADS) coverage_final_kernel(const double *partial, double inv_n, float *out) {
  __shared__ double sm[OPT_THREADS];
  double acc = 0.0;
  for (int i = threadIdx.x; i < OPT_BLOCKS; i += OPT_THREADS) acc += partial[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = static_cast<float>(sm[0] * inv_n);
}

}  // namespace

int coverage_launch(const float *alphas, int L, int rows, int n, void *scratch, float *out, cudaStream_t stream) {
  double *partial = static_cast<double *>(scratch);
  coverage_partial_kernel<<<OPT_BLOCKS, OPT_THREADS, 0, stream>>>(alphas, L, static_cast<size_t>(rows) * n, partial);
  note_launch();
  coverage_final_kernel<<<1, OPT_THREADS, 0, stream>>>(partial, 1.0 / n, out);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

size_t clip_scratch_bytes() { return OPT_BLOCKS * sizeof(double) + 16; }

int grad_clip_launch(float *grads, size_t n, float clip_c, void *scratch, cudaStream_t stream) {
  double *partial = static_cast<double *>(scratch);
  float *out = reinterpret_cast<float *>(partial + OPT_BLOCKS);
  sumsq_partial_kernel<<<OPT_BLOCKS, OPT_THREADS, 0, stream>>>(grads, n, partial);
  note_launch();
  clip_factor_kernel<<<1, OPT_THREADS, 0, stream>>>(partial, clip_c, out);
  note_launch();
  scale_by_kernel<<<OPT_BLOCKS, OPT_THREADS, 0, stream>>>(grads, n, out);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

int adam_launch(float *p, const float *g, float *m, float *v, size_t n, int step, cudaStream_t stream) {
  // the scalar schedule in fp32, as the reference's float32 graph computes it
  const float b1 = 0.1f, b2 = 0.001f, lr0 = 0.0002f;
  const float it = static_cast<float>(step);
  const float fix1 = 1.0f - powf(b1, it), fix2 = 1.0f - powf(b2, it);
  const float lr_t = lr0 * (sqrtf(fix2) / fix1);
  adam_kernel<<<OPT_BLOCKS, OPT_THREADS, 0, stream>>>(p, g, m, v, n, lr_t);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

int adadelta_launch(float *p, const float *g, float *rg2, float *ru2, size_t n, int phase, cudaStream_t stream) {
  adadelta_kernel<<<OPT_BLOCKS, OPT_THREADS, 0, stream>>>(p, g, rg2, ru2, n, phase);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

}  // namespace stat
