// FFMA vs FFMA2 issue rate on one SM-full of warps: cycles per warp-instruction.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float *out, int n) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float b = out[0], c = out[1];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
    a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
  }
  long long t1 = clock64();
  out[2 + threadIdx.x + blockIdx.x * blockDim.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = 0.f * (float)(t1 - t0), printf("FFMA : %.2f cycles per warp-instruction per SMSP (4 warps/SMSP)\n", (double)(t1 - t0) / (8.0 * n) / 4.0 * 1.0);
}
__global__ void k_ffma2(float *out, int n) {
  float2 a0 = make_float2(threadIdx.x, 1), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
  a1.x += 1; a2.x += 2; a3.x += 3; a4.x += 4; a5.x += 5; a6.x += 6; a7.x += 7;
  float2 b = make_float2(out[0], out[1]), c = make_float2(out[1], out[0]);
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    a0 = __ffma2_rn(a0, b, c); a1 = __ffma2_rn(a1, b, c); a2 = __ffma2_rn(a2, b, c); a3 = __ffma2_rn(a3, b, c);
    a4 = __ffma2_rn(a4, b, c); a5 = __ffma2_rn(a5, b, c); a6 = __ffma2_rn(a6, b, c); a7 = __ffma2_rn(a7, b, c);
  }
  long long t1 = clock64();
  out[2 + threadIdx.x + blockIdx.x * blockDim.x] = a0.x + a1.x + a2.x + a3.x + a4.y + a5.y + a6.y + a7.y;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("FFMA2: %.2f cycles per warp-instruction per SMSP (4 warps/SMSP)\n", (double)(t1 - t0) / (8.0 * n) / 4.0);
}
int main() {
  float *d;
  cudaMalloc(&d, 1 << 20);
  cudaMemset(d, 0, 1 << 20);
  for (int r = 0; r < 2; ++r) {
    k_ffma<<<1, 512>>>(d, 4096);
    cudaDeviceSynchronize();
    k_ffma2<<<1, 512>>>(d, 4096);
    cudaDeviceSynchronize();
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
