// Does cp.async.bulk with an L2 evict_last policy keep a cyclically re-read buffer resident across an
// interleaved 48 MB stream?  X (S MB) is re-read every iteration by 128 CTAs, W (48 MB) in between.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_persist_probe l2_persist_probe.cu
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
// mode 0: no hint; 1: constant evict_last; 2: createpolicy evict_last 1.0; 3: constant evict_first; 4: LDG loop (no TMA)
__global__ void __launch_bounds__(128, 1) probe(const uint8_t *base, int nchunk, uint32_t chunk, int ns, int mode, float *sink, int rot) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm);
  uint8_t *slots = sm + 1024;
  const size_t cta = (blockIdx.x + rot) % gridDim.x;   // rot != 0: another SM reads this CTA's data
  float acc = 0.f;
  if (mode == 4) {
    const float4 *p = reinterpret_cast<const float4 *>(base + cta * nchunk * (size_t)chunk);
    const size_t n = (size_t)nchunk * chunk / 16;
    for (size_t i = threadIdx.x; i < n; i += 128 * 4) {
      float4 a = p[i], b = i + 128 < n ? p[i + 128] : a, c = i + 256 < n ? p[i + 256] : a, d = i + 384 < n ? p[i + 384] : a;
      acc += a.x + b.x + c.x + d.x;
    }
    if (acc == 12345.678f) sink[0] = acc;
    return;
  }
  uint64_t pol = 0;
  if (mode == 1) pol = 0x14F0000000000000ull;
  if (mode == 3) pol = 0x12F0000000000000ull;
  if (mode == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  if (threadIdx.x == 0) {
    for (int s = 0; s < ns; ++s) mb_init(s32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int c, int s) {
    const uint32_t b = s32(bars + s), dst = s32(slots + (size_t)s * chunk);
    const void *src = base + (cta * nchunk + c) * (size_t)chunk;
    mb_expect(b, chunk);
    if (mode == 0)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(chunk), "r"(b) : "memory");
    else
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(chunk), "r"(b), "l"(pol) : "memory");
  };
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
int main() {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  int mx = 0;
  CK(cudaDeviceGetAttribute(&mx, cudaDevAttrMaxPersistingL2CacheSize, 0));
  uint8_t *x, *w;
  float *sink;
  CK(cudaMalloc(&x, 128ull << 20));
  CK(cudaMalloc(&w, 64ull << 20));
  CK(cudaMalloc(&sink, 64));
  CK(cudaMemset(x, 1, 128ull << 20));
  CK(cudaMemset(w, 1, 64ull << 20));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const uint32_t chunk = 32768;
  const int ns = 6, ctas = 128;
  const size_t smem = 1024 + (size_t)ns * chunk;
  const char *names[] = {"no hint", "evict_last const", "evict_last createpolicy", "evict_first const", "LDG"};
  for (int carve = 0; carve < 2; ++carve) {
    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve ? (size_t)mx : 0));
    printf("persisting carve-out %d bytes\n", carve ? mx : 0);
    for (int rotate = 0; rotate < 2; ++rotate)
    for (int S : {64, 80}) {
      const int nchunk = (int)(((size_t)S << 20) / ((size_t)ctas * chunk));
      const int nchunk_w = (int)((48ull << 20) / ((size_t)ctas * chunk));
      for (int mode : {0, 1}) {
        for (int wmode : {0, 3}) {
          float tx = 0.f;
          for (int it = 0; it < 12; ++it) {
            probe<<<ctas, 128, smem>>>(w, nchunk_w, chunk, ns, wmode, sink, 0);
            CK(cudaEventRecord(e0));
            probe<<<ctas, 128, smem>>>(x, nchunk, chunk, ns, mode, sink, rotate ? it * 37 : 0);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it >= 4) tx += ms;
          }
          tx /= 8;
          printf("  rotate=%d X=%3d MB %-24s | W=48 MB %-17s : X pass %6.1f us -> %.2f TB/s\n", rotate, S, names[mode], names[wmode], tx * 1e3,
                 (double)nchunk * ctas * chunk / (tx * 1e-3) / 1e12);
        }
      }
    }
  }
  CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
  return 0;
}
