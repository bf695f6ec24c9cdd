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

// x_new = prox_{tau f}(x - tau (ATz0 + D^T z1)),  xbar = (1 + alpha) x_new - alpha x
//   z1: (3, n0, n1, n2) dual of the gradient;  lo_halo: plane z1[0][-1] of the previous slab
__global__ void __launch_bounds__(256)
tv_primal_kernel(TvDims d, float* __restrict__ x, float* __restrict__ xbar, const float* __restrict__ atz,
                 const float* __restrict__ z1, const float* __restrict__ lo_halo, float tau, float alpha,
                 int nonneg) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
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
  }
}

// z1 = conj_prox_{sigma, lam ||.||_{2,1}}(z1 + sigma D xbar)
//   hi_halo: plane xbar[n0] of the next slab
__global__ void __launch_bounds__(256)
tv_dual_kernel(TvDims d, float* __restrict__ z1, const float* __restrict__ xbar, const float* __restrict__ hi_halo,
               float sigma, float lam) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
  const float inv_sigma = 1.0f / sigma;
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
    const float p0 = z1[idx] + sigma * d0, p1 = z1[idx + n] + sigma * d1, p2 = z1[idx + 2 * n] + sigma * d2;
    // conj_prox: p - sigma * prox_{(lam/sigma) ||.||}(p / sigma)
    // (v = p / sigma is evaluated as p * (1/sigma): 1 ulp from the reference's division, far inside
    // the 1e-5 tolerance, and keeps this kernel HBM-bound instead of issue-bound)
    const float v0 = p0 * inv_sigma, v1 = p1 * inv_sigma, v2 = p2 * inv_sigma;
    const float len = sqrtf((v0 * v0 + v1 * v1) + v2 * v2);
    float nl = len - lam * inv_sigma;
    nl = 0.5f * (nl + fabsf(nl));
    const float sc = len != 0.f ? __fdividef(nl, len) : 0.f;
    z1[idx] = p0 - sigma * (v0 * sc);
    z1[idx + n] = p1 - sigma * (v1 * sc);
    z1[idx + 2 * n] = p2 - sigma * (v2 * sc);
  }
}

// z0 = conj_prox_{sigma, 1/2 ||. - y||^2}(z0 + sigma A xbar) = p - sigma * ((y / sigma + p / sigma) / (1 / sigma + 1))
__global__ void __launch_bounds__(256)
l2_dual_kernel(size_t n, float* __restrict__ z0, const float* __restrict__ ax, const float* __restrict__ y, float sigma) {
  const float c = 1.0f / sigma;  // 2 * scale * lam with scale = 1/2, lam = 1/sigma
  const float rc1 = 1.0f / (c + 1.0f);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const float p = z0[idx] + sigma * ax[idx];
    const float v = p * c;
    z0[idx] = p - sigma * ((c * y[idx] + v) * rc1);
  }
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
