// Micro-benchmarks backing the design choices in DESIGN.md (run on the B200 via gpurun):
//   1. shared-memory fp32 atomicAdd (conflict-free addresses) vs plain LDS/FADD/STS RMW
//   2. coalesced RED.ADD.F32 to global memory (the forward kernel's flush)
//   3. F2I.FLOOR / FRND.CEIL conversion throughput vs FADD
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_smem_atomic(float* out, int stride) {
  __shared__ float acc[8][128];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) (&acc[0][0])[i] = 0.f;
  __syncthreads();
  float v = 1.0f + lane;
  for (int it = 0; it < ITERS; ++it) {
    int idx = (lane * stride + it) & 127;
    atomicAdd(&acc[warp][idx], v);
  }
  __syncthreads();
  if (threadIdx.x < 128) out[blockIdx.x * 128 + threadIdx.x] = acc[0][threadIdx.x] + acc[7][threadIdx.x];
}

__global__ void k_smem_rmw(float* out, int stride) {
  __shared__ float2 acc[8][128];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) (&acc[0][0])[i] = make_float2(0.f, 0.f);
  __syncthreads();
  float v = 1.0f + lane;
  for (int it = 0; it < ITERS; ++it) {
    int idx = (lane * stride + it) & 127;
    float2 a = acc[warp][idx];
    a.x = fmaf(v, 0.25f, a.x);
    a.y = fmaf(v, 0.75f, a.y);
    acc[warp][idx] = a;
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < 128) out[blockIdx.x * 128 + threadIdx.x] = acc[0][threadIdx.x].x + acc[7][threadIdx.x].y;
}

__global__ void k_red_global(float* buf, size_t n) {
  // every warp adds to 32 consecutive floats; rows rotate so that different warps rarely collide
  size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  size_t rows = n / 32;
  for (int it = 0; it < 256; ++it) {
    size_t row = (warp * 257 + (size_t)it * 7919) % rows;
    atomicAdd(buf + row * 32 + lane, 1.0f);
  }
}

__global__ void k_conv(float* out, float seed) {
  float a = seed + threadIdx.x * 0.37f, s = 0.f;
  int si = 0;
  for (int it = 0; it < ITERS; ++it) {
    si += __float2int_rd(a);
    s += ceilf(a) - a;
    a += 0.731f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + si;
}
__global__ void k_fadd(float* out, float seed) {
  float a = seed + threadIdx.x * 0.37f, s = 0.f, s2 = 0.f;
  for (int it = 0; it < ITERS; ++it) {
    s2 += a * 1.0001f;
    s += (a + 0.5f) - a;
    a += 0.731f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + s2;
}

template <class F> float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();  // warm-up
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  float* out; CK(cudaMalloc(&out, sizeof(float) * (1 << 24)));
  const int blocks = sms * 8, threads = 256;
  for (int stride : {1, 3}) {
    float ms = time_ms([&] { k_smem_atomic<<<blocks, threads>>>(out, stride); });
    double ops = (double)blocks * threads * ITERS;
    printf("smem atomicAdd f32 stride %d : %.3f ms  %.3e lane-ops/s  (%.2f lane-ops/clk/SM @1.9GHz)\n", stride, ms, ops / ms * 1e3, ops / ms * 1e3 / sms / 1.9e9);
    ms = time_ms([&] { k_smem_rmw<<<blocks, threads>>>(out, stride); });
    printf("smem float2 RMW    stride %d : %.3f ms  %.3e lane-ops/s  (%.2f lane-ops/clk/SM @1.9GHz)\n", stride, ms, ops / ms * 1e3, ops / ms * 1e3 / sms / 1.9e9);
  }
  {
    size_t n = (size_t)1 << 28;  // 1 GiB of floats: larger than L2
    float* buf; CK(cudaMalloc(&buf, n * sizeof(float))); CK(cudaMemset(buf, 0, n * sizeof(float)));
    for (size_t span : {(size_t)1 << 22, (size_t)1 << 28}) {
      float ms = time_ms([&] { k_red_global<<<sms * 16, 256>>>(buf, span); });
      double ops = (double)sms * 16 * 256 * 256;
      printf("RED.ADD.F32 coalesced, span %zu MiB: %.3f ms  %.3e lane-ops/s  (%.1f GB/s of 4B ops)\n", span * 4 >> 20, ms, ops / ms * 1e3, ops * 4 / ms * 1e3 / 1e9);
    }
    cudaFree(buf);
  }
  {
    float ms = time_ms([&] { k_conv<<<blocks, threads>>>(out, 1.5f); });
    double ops = (double)blocks * threads * ITERS;
    printf("F2I.FLOOR+FRND.CEIL+2FADD loop: %.3f ms  %.3e iters/s\n", ms, ops / ms * 1e3);
    ms = time_ms([&] { k_fadd<<<blocks, threads>>>(out, 1.5f); });
    printf("FMUL/FADD-only loop           : %.3f ms  %.3e iters/s\n", ms, ops / ms * 1e3);
  }
  return 0;
}
