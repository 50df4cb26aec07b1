// Streaming probe: how fast can 128/148 CTAs pull the attention context blocks (95 MB at B=64)
// through cp.async.bulk into shared memory, with no compute?  Compares the seven-array layout
// with a per-frame record layout, slot counts and chunk sizes.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_wait(uint32_t b, uint32_t ph) {
  uint32_t ok, spins = 0;
  do {
    if (++spins > (1u << 24)) __trap();
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk(uint32_t dst, const void *src, uint32_t n, uint32_t b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(n), "r"(b) : "memory");
}

// Each CTA streams `nchunk` chunks of `chunk` bytes; NS slots in flight.  mode 0: chunk c of CTA i is
// contiguous at base + (i*nchunk + c)*chunk.  mode 1: a chunk is `pieces` pieces of chunk/pieces bytes,
// piece p taken from array p (arrays `astride` bytes apart), like the seven-array layout.
__global__ void __launch_bounds__(128, 1) probe(const uint8_t *base, int nchunk, uint32_t chunk, int ns, int pieces,
                                               size_t astride, float *sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm);
  uint8_t *slots = sm + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ns; ++s) mb_init(s32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t cta = blockIdx.x;
  auto issue = [&](int c, int s) {
    const uint32_t b = s32(bars + s), dst = s32(slots + (size_t)s * chunk);
    mb_expect(b, chunk);
    if (pieces <= 1) {
      bulk(dst, base + (cta * nchunk + c) * (size_t)chunk, chunk, b);
    } else {
      const uint32_t pb = chunk / pieces;
      for (int p = 0; p < pieces; ++p) bulk(dst + p * pb, base + p * astride + (cta * nchunk + c) * (size_t)pb, pb, b);
    }
  };
  float acc = 0.f;
  if (threadIdx.x == 0)
    for (int s = 0; s < ns && s < nchunk; ++s) issue(s, s);
  for (int c = 0; c < nchunk; ++c) {
    const int s = c % ns;
    mb_wait(s32(bars + s), (c / ns) & 1);
    acc += reinterpret_cast<const float *>(slots + (size_t)s * chunk)[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0 && c + ns < nchunk) issue(c + ns, s);
  }
  if (acc == 12345.678f) sink[0] = acc;
}

__global__ void flush_read(const float4 *p, size_t n, float *sink) {
  float a = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a += p[i].x;
  if (a == 1.2345f) sink[0] = a;
}

int main() {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  const size_t total = 96ull << 20;                  // ~ the 95 MB of context blocks
  uint8_t *buf, *fl;
  float *sink;
  CK(cudaMalloc(&buf, total + (8 << 20)));
  CK(cudaMalloc(&fl, 512ull << 20));
  CK(cudaMalloc(&sink, 64));
  CK(cudaMemset(buf, 1, total));
  CK(cudaMemset(fl, 1, 512ull << 20));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  struct Cfg { int ctas; uint32_t chunk; int ns; int pieces; const char *name; };
  const Cfg cfgs[] = {
      {128, 57344, 4, 1, "128 CTAs, 56 KB contiguous records, 4 slots"},
      {148, 57344, 4, 1, "148 CTAs, 56 KB contiguous records, 4 slots"},
      {148, 28672, 8, 1, "148 CTAs, 28 KB contiguous, 8 slots"},
      {148, 16384, 14, 1, "148 CTAs, 16 KB contiguous, 14 slots"},
      {148, 57344, 4, 7, "148 CTAs, 7 pieces of 8 KB from 7 arrays, 4 slots"},
      {128, 57344, 4, 7, "128 CTAs, 7 pieces of 8 KB from 7 arrays, 4 slots"},
      {148, 57344, 2, 1, "148 CTAs, 56 KB contiguous, 2 slots"},
      {148, 57344, 3, 1, "148 CTAs, 56 KB contiguous, 3 slots"},
      {148, 8192, 28, 1, "148 CTAs, 8 KB contiguous, 28 slots"},
      {148, 2048, 96, 1, "148 CTAs, 2 KB contiguous, 96 slots"},
  };
  for (const Cfg &c : cfgs) {
    const int nchunk = (int)(total / ((size_t)c.ctas * c.chunk));
    const size_t moved = (size_t)nchunk * c.ctas * c.chunk;
    const size_t astride = (total / 7) & ~(size_t)4095;
    const size_t smem = 1024 + (size_t)c.ns * c.chunk;
    for (int cold = 1; cold >= 0; --cold) {
      float best = 1e9f, sum = 0.f;
      for (int it = 0; it < 5; ++it) {
        if (cold) flush_read<<<592, 256>>>(reinterpret_cast<const float4 *>(fl), (512ull << 20) / 16, sink);
        CK(cudaEventRecord(e0));
        probe<<<c.ctas, 128, smem>>>(buf, nchunk, c.chunk, c.ns, c.pieces, astride, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it) { sum += ms; if (ms < best) best = ms; }
      }
      printf("%-52s %s: %6.1f us avg %6.1f us best -> %.2f TB/s (%.1f MB)\n", c.name, cold ? "cold" : "warm", sum / 4 * 1e3,
             best * 1e3, moved / (best * 1e-3) / 1e12, moved / 1048576.0);
    }
  }
  CK(cudaGetLastError());
  return 0;
}
