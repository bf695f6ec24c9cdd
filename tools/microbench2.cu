// Micro-benchmarks for the slice-lane forward design (run on the B200 via gpurun):
//   1. red.global.add.f32 scalar vs red.global.add.v4.f32 (16 B per lane), coalesced 128 B / 512 B
//      per warp instruction, footprint 16 MiB .. 1 GiB
//   2. shared-memory float atomicAdd in the pattern the design uses: all 32 lanes of a warp add to
//      32 consecutive floats (no intra-warp conflict), 16 warps of a CTA hitting a 32-row window
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench2 microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void red_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void k_red1(float* buf, size_t n, int iters) {
  size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  size_t rows = n / 32;
  for (int it = 0; it < iters; ++it) {
    size_t row = (warp * 257 + (size_t)it * 7919) % rows;
    atomicAdd(buf + row * 32 + lane, 1.0f);
  }
}
__global__ void k_red4(float* buf, size_t n, int iters) {
  size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  size_t rows = n / 128;
  for (int it = 0; it < iters; ++it) {
    size_t row = (warp * 257 + (size_t)it * 7919) % rows;
    red_v4(buf + row * 128 + lane * 4, 1.0f, 2.0f, 3.0f, 4.0f);
  }
}

// 16 warps; each iteration every warp adds 4 floats per lane (128 consecutive floats x ... ) to row r
__global__ void k_smem_atomic_rows(float* out, int iters, int use_atomic) {
  __shared__ __align__(16) float win[32][128];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 128; i += blockDim.x) (&win[0][0])[i] = 0.f;
  __syncthreads();
  float v = 1.0f + lane;
  for (int it = 0; it < iters; ++it) {
    int r = (warp * 2 + it) & 31;
    float* p = &win[r][lane * 4];
    if (use_atomic) {
      atomicAdd(p, v); atomicAdd(p + 1, v); atomicAdd(p + 2, v); atomicAdd(p + 3, v);
    } else {
      float4 a = *reinterpret_cast<float4*>(p);
      a.x += v; a.y += v; a.z += v; a.w += v;
      *reinterpret_cast<float4*>(p) = a;
    }
  }
  __syncthreads();
  if (threadIdx.x < 128) out[blockIdx.x * 128 + threadIdx.x] = win[0][threadIdx.x] + win[31][threadIdx.x];
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  for (size_t mib : {16, 64, 256, 1024}) {
    size_t n = mib * 1024 * 1024 / 4;
    float* buf; CK(cudaMalloc(&buf, n * 4)); CK(cudaMemset(buf, 0, n * 4));
    int blocks = prop.multiProcessorCount * 8, iters = 512;
    k_red1<<<blocks, 256>>>(buf, n, 16); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k_red1<<<blocks, 256>>>(buf, n, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double ops = (double)blocks * 256 * iters;
    printf("RED.f32    span %4zu MiB: %7.3f ms  %.3e floats/s  %7.1f GB/s\n", mib, ms, ops / (ms * 1e-3), ops * 4 / (ms * 1e-3) / 1e9);
    k_red4<<<blocks, 256>>>(buf, n, 16); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k_red4<<<blocks, 256>>>(buf, n, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("RED.v4.f32 span %4zu MiB: %7.3f ms  %.3e floats/s  %7.1f GB/s\n", mib, ms, ops * 4 / (ms * 1e-3), ops * 16 / (ms * 1e-3) / 1e9);
    CK(cudaFree(buf));
  }
  float* out; CK(cudaMalloc(&out, 148 * 8 * 128 * 4));
  for (int use_atomic = 0; use_atomic < 2; ++use_atomic) {
    int iters = 8192, blocks = prop.multiProcessorCount;
    k_smem_atomic_rows<<<blocks, 512>>>(out, 64, use_atomic); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k_smem_atomic_rows<<<blocks, 512>>>(out, iters, use_atomic); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double ops = (double)blocks * 512 * iters * 4;
    printf("smem %s rows (16 warps x 128 floats): %7.3f ms  %.3e floats/s  %.2f floats/clk/SM @1.9GHz\n",
           use_atomic ? "atomicAdd" : "plain RMW", ms, ops / (ms * 1e-3), ops / (ms * 1e-3) / blocks / 1.9e9);
  }
  return 0;
}
