// Probe: which (box, L2 promotion, smem alignment) combinations of cp.async.bulk.tensor.3d work on this part.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu
//   tools/tma_probe <boxW> <boxH> <promo 0..3> <smem_offset_bytes> <D1> <D0> <x> <y>
// One configuration per process (a faulting TMA poisons the context).  Prints OK / MISMATCH / the CUDA error.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int box_elems, int smem_off, int x, int y, int z) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* dst = reinterpret_cast<float*>(smem + smem_off);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + 32768);
  const unsigned bar_sa = (unsigned)__cvta_generic_to_shared(bar), dst_sa = (unsigned)__cvta_generic_to_shared(dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_sa) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_sa), "r"(box_elems * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst_sa),
        "l"(&tmap), "r"(bar_sa), "r"(x), "r"(y), "r"(z)
        : "memory");
  }
  unsigned done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar_sa), "r"(0u)
        : "memory");
  } while (!done);
  for (int i = threadIdx.x; i < box_elems; i += 32) out[i] = dst[i];
}

int main(int argc, char** argv) {
  if (argc < 9) return 2;
  const int bw = atoi(argv[1]), bh = atoi(argv[2]), promo = atoi(argv[3]), off = atoi(argv[4]);
  const int D1 = atoi(argv[5]), D0 = atoi(argv[6]), x = atoi(argv[7]), y = atoi(argv[8]);
  const int V = 3, z = 1;
  std::vector<float> h((size_t)V * D0 * D1);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, bw * bh * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)D1, (cuuint64_t)D0, (cuuint64_t)V};
  const cuuint64_t strides[2] = {(cuuint64_t)D1 * 4, (cuuint64_t)D0 * D1 * 4};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u}, es[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<Fn>(f)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %dx%d promo %d smem_off %d D %dx%d at (%d,%d): encode=%d ", bw, bh, promo, off, D0, D1, x, y, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 0; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  probe<<<1, 32, 40960>>>(tm, o, bw * bh, off, x, y, z);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 0; }
  std::vector<float> got(bw * bh);
  cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r_ = 0; r_ < bh; ++r_)
    for (int c = 0; c < bw; ++c) {
      const int row = y + r_, col = x + c;
      const float want = (row >= 0 && row < D0 && col >= 0 && col < D1) ? h[((size_t)z * D0 + row) * D1 + col] : 0.f;
      bad += got[r_ * bw + c] != want;
    }
  printf("%s (%d mismatches)\n", bad ? "MISMATCH" : "OK", bad);
  return 0;
}
