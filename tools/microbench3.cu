// Issue-rate micro-benchmark for the packed fp32 instructions of sm_100 (FFMA2 / FADD2) against
// scalar FFMA, and of LDS.32 / LDS.64 / LDS.128 with a conflict-free pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench3 tools/microbench3.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) fma_kernel(float* out, int iters, float w) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  const float2 x = make_float2(1.0001f, 0.9999f), ww = make_float2(w, w);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // 2 scalar FFMA
        a[i].x = fmaf(a[i].x, w, x.x);
        a[i].y = fmaf(a[i].y, w, x.y);
      } else if (MODE == 1) {  // 1 FFMA2
        a[i] = __ffma2_rn(a[i], ww, x);
      } else {  // 1 FADD2
        a[i] = __fadd2_rn(a[i], x);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int BYTES>
__global__ void __launch_bounds__(256) lds_kernel(float* out, int iters) {
  __shared__ __align__(16) float sm[256 * 4 + 64];
  for (int i = threadIdx.x; i < 256 * 4 + 64; i += 256) sm[i] = i;
  __syncthreads();
  float s = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int off = warp * 128 + lane * (BYTES / 4);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float* p = sm + ((off + k * 4) & 1023 & ~(BYTES / 4 - 1));
      if (BYTES == 4) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
        s += v;
      } else if (BYTES == 8) {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
        s += v.x + v.y;
      } else {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p)));
        s += v.x + v.y + v.z + v.w;
      }
    }
    off += (int)s & 4;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int blocks = 148 * 8, iters = 20000;
  const char* names[3] = {"2x FFMA (scalar)", "1x FFMA2", "1x FADD2"};
  for (int m = 0; m < 3; ++m) {
    float ms = m == 0 ? time_ms([&] { fma_kernel<0><<<blocks, 256>>>(out, iters, 0.5f); })
             : m == 1 ? time_ms([&] { fma_kernel<1><<<blocks, 256>>>(out, iters, 0.5f); })
                      : time_ms([&] { fma_kernel<2><<<blocks, 256>>>(out, iters, 0.5f); });
    double pairs = (double)blocks * 256 * iters * 8;  // float2 results
    printf("%-18s %8.3f ms  %.2f float2-results/clk/SM (%.1f flop-lanes/clk/SM)\n", names[m], ms,
           pairs / (ms * 1e-3) / 148 / (clk_khz * 1e3), 2 * pairs / (ms * 1e-3) / 148 / (clk_khz * 1e3));
  }
  {
    float ms4 = time_ms([&] { lds_kernel<4><<<148 * 8, 256>>>(out, 4000); });
    float ms8 = time_ms([&] { lds_kernel<8><<<148 * 8, 256>>>(out, 4000); });
    float ms16 = time_ms([&] { lds_kernel<16><<<148 * 8, 256>>>(out, 4000); });
    double n = (double)148 * 8 * 8 * 4000 * 8;  // warp-level LDS instructions
    printf("LDS.32  %8.3f ms  %.3f warp-instr/clk/SM\n", ms4, n / (ms4 * 1e-3) / 148 / (clk_khz * 1e3));
    printf("LDS.64  %8.3f ms  %.3f warp-instr/clk/SM\n", ms8, n / (ms8 * 1e-3) / 148 / (clk_khz * 1e3));
    printf("LDS.128 %8.3f ms  %.3f warp-instr/clk/SM\n", ms16, n / (ms16 * 1e-3) / 148 / (clk_khz * 1e3));
  }
  return 0;
}
