// Issue rate of the legacy tensor path on sm_100a: mma.sync.m16n8k8 tf32 (and m16n8k16 bf16) per warp scheduler.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_tf32(float *out, int n) {
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 7, b1 = 9;
  float c[8][4] = {};
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("mma.sync m16n8k8 tf32: %.2f cycles per MMA per scheduler (%d warps/SM) -> %.0f MAC/clk/SM\n",
           (double)(t1 - t0) / (8.0 * n) / (blockDim.x / 128), blockDim.x / 32, 1024.0 * 4 / ((double)(t1 - t0) / (8.0 * n) / (blockDim.x / 128)));
}
__global__ void k_bf16(float *out, int n) {
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 7, b1 = 9;
  float c[8][4] = {};
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("mma.sync m16n8k16 bf16: %.2f cycles per MMA per scheduler (%d warps/SM) -> %.0f MAC/clk/SM\n",
           (double)(t1 - t0) / (8.0 * n) / (blockDim.x / 128), blockDim.x / 32, 2048.0 * 4 / ((double)(t1 - t0) / (8.0 * n) / (blockDim.x / 128)));
}
int main() {
  float *d;
  cudaMalloc(&d, 1 << 20);
  for (int threads : {128, 256, 512}) {
    k_tf32<<<1, threads>>>(d, 2048);
    cudaDeviceSynchronize();
    k_bf16<<<1, threads>>>(d, 2048);
    cudaDeviceSynchronize();
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
