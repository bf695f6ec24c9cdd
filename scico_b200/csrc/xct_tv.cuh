// Fused elementwise / stencil kernels of the TV-regularised PDHG iteration around the projector
// pair.  They restate, fused and on device memory,
//   FiniteDifference(append=0) forward/adjoint   scico/linop/_diff.py:25-96,236-272
//   L21Norm.prox (l2_axis=0)                     scico/functional/_norm.py:254-263
//   SquaredL2Loss(y).prox                        scico/loss.py:220-226
//   Functional.conj_prox                         scico/functional/_functional.py:102-128
//   PDHG.step                                    scico/optimize/_primaldual.py:219-231
// All are HBM-bound streaming kernels: one thread per voxel / sinogram element, coalesced along
// the fastest axis; neighbours along axes 1 and 0 are whole rows / planes away and hit L1/L2.
// z-slab sharding: the volume block may be a slab of a larger volume; `lo_halo` / `hi_halo` are the
// neighbouring ranks' boundary planes (nullptr at the global boundary).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace xct {

struct TvDims {
  int n0, n1, n2;    // local block
  int first, last;   // 1 if this block holds the global first / last slice along axis 0
};

// Sum of one double per thread over the block, added to a device scalar with one atomicAdd(double) per
// block.  Every thread of the block must call it; two calls in one kernel need a __syncthreads() between.
__device__ __forceinline__ void block_reduce_add(double v, double* slot) {
  __shared__ double sh[32];
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = l < nw ? sh[l] : 0.0;
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (l == 0) atomicAdd(slot, v);
  }
}

// Which elements of a (V, rows, inner) sinogram block count in global sums: detector rows [lo, hi) of the
// block (z-slab sharding: rows shared with another slab are counted by their owner only).
struct SinoRows {
  long long inner;
  int rows, lo, hi;
  __device__ __forceinline__ bool counted(size_t idx) const {
    if (lo <= 0 && hi >= rows) return true;
    const int r = (int)((idx / (size_t)inner) % (size_t)rows);
    return r >= lo && r < hi;
  }
};

// x_new = prox_{tau f}(x - tau (ATz0 + D^T z1)),  xbar = (1 + alpha) x_new - alpha x
//   z1: (3, n0, n1, n2) dual of the gradient;  lo_halo: plane z1[0][-1] of the previous slab
// STAT: *stat += ||x_new - x_old||^2 (the primal residual of the iteration statistics, _primaldual.py:175-189)
template <bool STAT>
__global__ void __launch_bounds__(256)
tv_primal_kernel(TvDims d, float* __restrict__ x, float* __restrict__ xbar, const float* __restrict__ atz,
                 const float* __restrict__ z1, const float* __restrict__ lo_halo, float tau, float alpha,
                 int nonneg, double* __restrict__ stat) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
  double acc = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % d.n2);
    const size_t ij = idx / d.n2;
    const int j = (int)(ij % d.n1), i = (int)(ij / d.n1);
    const float* za = z1;
    const float* zb = z1 + n;
    const float* zc = z1 + 2 * n;
    // (D^T z)_a[i] = z_a[i-1] - z_a[i], with z_a[n-1] := 0 (zero row of D) and z_a[-1] := 0
    float t0 = 0.f;
    if (!(d.last && i == d.n0 - 1)) t0 = -za[idx];
    if (i > 0) t0 += za[idx - plane];
    else if (!d.first && lo_halo) t0 += lo_halo[(size_t)j * d.n2 + k];
    float t1 = (j < d.n1 - 1) ? -zb[idx] : 0.f;
    if (j > 0) t1 += zb[idx - d.n2];
    float t2 = (k < d.n2 - 1) ? -zc[idx] : 0.f;
    if (k > 0) t2 += zc[idx - 1];
    const float xo = x[idx];
    const float ctz = atz[idx] + ((t0 + t1) + t2);
    float xn = xo - tau * ctz;
    if (nonneg) xn = fmaxf(xn, 0.f);
    x[idx] = xn;
    xbar[idx] = (1.f + alpha) * xn - alpha * xo;
    if (STAT) {
      const double dx = (double)(xn - xo);
      acc += dx * dx;
    }
  }
  if (STAT) block_reduce_add(acc, stat);
}

// z1 = conj_prox_{sigma, lam ||.||_{2,1}}(z1 + sigma D xbar)
//   hi_halo: plane xbar[n0] of the next slab
// STAT: *stat += ||z1_new - z1_old||^2
template <bool STAT>
__global__ void __launch_bounds__(256)
tv_dual_kernel(TvDims d, float* __restrict__ z1, const float* __restrict__ xbar, const float* __restrict__ hi_halo,
               float sigma, float lam, double* __restrict__ stat) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
  const float inv_sigma = 1.0f / sigma;
  double acc = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % d.n2);
    const size_t ij = idx / d.n2;
    const int j = (int)(ij % d.n1), i = (int)(ij / d.n1);
    const float xc = xbar[idx];
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    if (i < d.n0 - 1) d0 = xbar[idx + plane] - xc;
    else if (!d.last && hi_halo) d0 = hi_halo[(size_t)j * d.n2 + k] - xc;
    if (j < d.n1 - 1) d1 = xbar[idx + d.n2] - xc;
    if (k < d.n2 - 1) d2 = xbar[idx + 1] - xc;
    const float o0 = z1[idx], o1 = z1[idx + n], o2 = z1[idx + 2 * n];
    const float p0 = o0 + sigma * d0, p1 = o1 + sigma * d1, p2 = o2 + sigma * d2;
    // conj_prox: p - sigma * prox_{(lam/sigma) ||.||}(p / sigma)
    // (v = p / sigma is evaluated as p * (1/sigma): 1 ulp from the reference's division, far inside
    // the 1e-5 tolerance, and keeps this kernel HBM-bound instead of issue-bound)
    const float v0 = p0 * inv_sigma, v1 = p1 * inv_sigma, v2 = p2 * inv_sigma;
    const float len = sqrtf((v0 * v0 + v1 * v1) + v2 * v2);
    float nl = len - lam * inv_sigma;
    nl = 0.5f * (nl + fabsf(nl));
    const float sc = len != 0.f ? __fdividef(nl, len) : 0.f;
    const float q0 = p0 - sigma * (v0 * sc), q1 = p1 - sigma * (v1 * sc), q2 = p2 - sigma * (v2 * sc);
    z1[idx] = q0;
    z1[idx + n] = q1;
    z1[idx + 2 * n] = q2;
    if (STAT) {
      const double e0 = (double)(q0 - o0), e1 = (double)(q1 - o1), e2 = (double)(q2 - o2);
      acc += (e0 * e0 + e1 * e1) + e2 * e2;
    }
  }
  if (STAT) block_reduce_add(acc, stat);
}

// z0 = conj_prox_{sigma, 1/2 ||. - y||^2}(z0 + sigma A xbar) = p - sigma * ((y / sigma + p / sigma) / (1 / sigma + 1))
// STAT: the iteration statistics that involve the sinogram, in the same pass.  `ax` is A xbar with
// xbar = (1 + alpha) x_new - alpha x_old, so A x_new = (A xbar + alpha A x_old) / (1 + alpha): the kernel keeps
// ax_x = A x up to date by that recurrence (its rounding error is halved every iteration for alpha = 1: it stays
// at the fp32 rounding level, no drift) instead of a second forward projection per iteration, and adds
//   stat[0] += ||z0_new - z0_old||^2,  stat[1] += ||A x_new - y||^2   over the rows `rows` counts.
template <bool STAT>
__global__ void __launch_bounds__(256)
l2_dual_kernel(size_t n, float* __restrict__ z0, const float* __restrict__ ax, const float* __restrict__ y, float sigma,
               float* __restrict__ ax_x, float alpha, SinoRows rows, double* __restrict__ stat) {
  const float c = 1.0f / sigma;  // 2 * scale * lam with scale = 1/2, lam = 1/sigma
  const float rc1 = 1.0f / (c + 1.0f);
  const float r1a = 1.0f / (1.0f + alpha);
  double acc_z = 0.0, acc_r = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const float zo = z0[idx], a = ax[idx], yy = y[idx];
    const float p = zo + sigma * a;
    const float v = p * c;
    const float zn = p - sigma * ((c * yy + v) * rc1);
    z0[idx] = zn;
    if (STAT) {
      const float axn = (a + alpha * ax_x[idx]) * r1a;
      ax_x[idx] = axn;
      if (rows.counted(idx)) {
        const double dz = (double)(zn - zo), r = (double)(axn - yy);
        acc_z += dz * dz;
        acc_r += r * r;
      }
    }
  }
  if (STAT) {
    block_reduce_add(acc_z, stat);
    __syncthreads();
    block_reduce_add(acc_r, stat + 1);
  }
}

// *stat += ||D x||_{2,1} = sum over voxels of the Euclidean length of the finite-difference 3-vector
// (L21Norm with l2_axis = 0, scico/functional/_norm.py:225-252; D = FiniteDifference(append=0)).
__global__ void __launch_bounds__(256)
tv_norm_kernel(TvDims d, const float* __restrict__ x, const float* __restrict__ hi_halo, double* __restrict__ stat) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
  double acc = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % d.n2);
    const size_t ij = idx / d.n2;
    const int j = (int)(ij % d.n1), i = (int)(ij / d.n1);
    const float xc = x[idx];
    float d0 = 0.f;
    if (i < d.n0 - 1) d0 = x[idx + plane] - xc;
    else if (!d.last && hi_halo) d0 = hi_halo[(size_t)j * d.n2 + k] - xc;
    const float d1 = (j < d.n1 - 1) ? x[idx + d.n2] - xc : 0.f;
    const float d2 = (k < d.n2 - 1) ? x[idx + 1] - xc : 0.f;
    acc += sqrt(((double)d0 * d0 + (double)d1 * d1) + (double)d2 * d2);
  }
  block_reduce_add(acc, stat);
}

// FiniteDifference(append=0) forward: out (3, n0, n1, n2)
__global__ void __launch_bounds__(256)
fd_forward_kernel(TvDims d, const float* __restrict__ x, const float* __restrict__ hi_halo, float* __restrict__ out) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % d.n2);
    const size_t ij = idx / d.n2;
    const int j = (int)(ij % d.n1), i = (int)(ij / d.n1);
    const float xc = x[idx];
    float d0 = 0.f;
    if (i < d.n0 - 1) d0 = x[idx + plane] - xc;
    else if (!d.last && hi_halo) d0 = hi_halo[(size_t)j * d.n2 + k] - xc;
    out[idx] = d0;
    out[idx + n] = (j < d.n1 - 1) ? x[idx + d.n2] - xc : 0.f;
    out[idx + 2 * n] = (k < d.n2 - 1) ? x[idx + 1] - xc : 0.f;
  }
}

// FiniteDifference(append=0) adjoint: z (3, n0, n1, n2) -> out (n0, n1, n2)
__global__ void __launch_bounds__(256)
fd_adjoint_kernel(TvDims d, const float* __restrict__ z1, const float* __restrict__ lo_halo, float* __restrict__ out) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % d.n2);
    const size_t ij = idx / d.n2;
    const int j = (int)(ij % d.n1), i = (int)(ij / d.n1);
    float t0 = 0.f;
    if (!(d.last && i == d.n0 - 1)) t0 = -z1[idx];
    if (i > 0) t0 += z1[idx - plane];
    else if (!d.first && lo_halo) t0 += lo_halo[(size_t)j * d.n2 + k];
    float t1 = (j < d.n1 - 1) ? -z1[idx + n] : 0.f;
    if (j > 0) t1 += z1[idx + n - d.n2];
    float t2 = (k < d.n2 - 1) ? -z1[idx + 2 * n] : 0.f;
    if (k > 0) t2 += z1[idx + 2 * n - 1];
    out[idx] = (t0 + t1) + t2;
  }
}

}  // namespace xct
