// Fused kernels of the ADMM family around the projector pair:
//   ADMM + CG x-step           scico/optimize/_admm.py:334-378, _admmaux.py:231-269, scico/solver.py:367-405
//   LinearizedADMM.step        scico/optimize/_ladmm.py:253-277
//   ProximalADMM.step          scico/optimize/_padmm.py:349-363
// for the TV-regularised CT problems of the reference's examples (ct_tv_admm.py, ct_3d_tv_padmm.py):
//   C or A = FiniteDifference(append=0) (possibly scaled and stacked under the projector),
//   g = lam ||.||_{2,1}  (+ 1/2 ||. - y||^2 on the sinogram block).
// Everything is a single-pass HBM-bound stream: one thread per voxel / sinogram element, neighbours
// along axes 1 and 0 come from L1/L2.  Reductions (CG inner products) are accumulated in double per
// thread, reduced per block and added to a device scalar with one atomicAdd(double) per block; the
// scalars never leave the device inside an iteration (alpha and beta are recomputed per thread from
// them), so a CG iteration has no host synchronisation of its own.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "xct_tv.cuh"

namespace xct {

enum SplitMode { kSplitAdmm = 0, kSplitLadmm = 1, kSplitPadmm = 2 };

struct Vox {
  int i, j, k;
};
__device__ __forceinline__ Vox unravel(const TvDims& d, size_t idx) {
  Vox v;
  v.k = (int)(idx % d.n2);
  const size_t ij = idx / d.n2;
  v.j = (int)(ij % d.n1);
  v.i = (int)(ij / d.n1);
  return v;
}

// (D x)[idx] for the three axes (append=0: zero at the global upper boundary of each axis).
__device__ __forceinline__ void fd_fwd_at(const TvDims& d, const float* __restrict__ x, const float* __restrict__ hi_halo,
                                          size_t idx, const Vox& v, size_t plane, float& d0, float& d1, float& d2) {
  const float xc = x[idx];
  d0 = 0.f;
  if (v.i < d.n0 - 1) d0 = x[idx + plane] - xc;
  else if (!d.last && hi_halo) d0 = hi_halo[(size_t)v.j * d.n2 + v.k] - xc;
  d1 = (v.j < d.n1 - 1) ? x[idx + d.n2] - xc : 0.f;
  d2 = (v.k < d.n2 - 1) ? x[idx + 1] - xc : 0.f;
}

// prox of thr*||.||_2 on a 3-vector, scico/functional/_norm.py:254-263: returns the scale s with
// prox(v) = s * v  (s = max(|v| - thr, 0) / |v|, 0 when |v| = 0).
__device__ __forceinline__ float l21_scale(float v0, float v1, float v2, float thr) {
  const float len = sqrtf((v0 * v0 + v1 * v1) + v2 * v2);
  float nl = len - thr;
  nl = 0.5f * (nl + fabsf(nl));
  return len != 0.f ? __fdividef(nl, len) : 0.f;
}

// Gradient block of the split: with Cx = dscale * D x,
//   ADMM / LADMM: z = prox_{thr ||.||_{2,1}}(Cx + u);           u = u + Cx - z     (_admm.py:366-378, _ladmm.py:274-277)
//   PADMM:        z = prox_{thr ||.||_{2,1}}(z + inv_nu ((Cx - z) + u)); u = (u + Cx) - z   (_padmm.py:354-363)
// and the array the NEXT x-step applies D^T to:
//   LADMM: w = (Cx - z) + u (new z, u; _ladmm.py:270)     PADMM: w = 2 u_new - u_old (_padmm.py:351)
// STAT (iteration statistics in the same pass, _padmm.py:148-177,294-345, _ladmm.py:160-200):
//   stat[0] += ||Cx - z_new||^2 (primal residual),  stat[1] += ||z_new - z_old||^2 (fast dual residual),
//   stat[2] += ||z_new||_{2,1} (the gradient block's share of g(z))
template <int MODE, bool STAT>
__global__ void __launch_bounds__(256)
grad_prox_kernel(TvDims d, const float* __restrict__ x, const float* __restrict__ hi_halo, float* __restrict__ z,
                 float* __restrict__ u, float* __restrict__ w, float dscale, float thr, float inv_nu,
                 double* __restrict__ stat) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
  double acc_p = 0.0, acc_d = 0.0, acc_g = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const Vox v = unravel(d, idx);
    float c0, c1, c2;
    fd_fwd_at(d, x, hi_halo, idx, v, plane, c0, c1, c2);
    c0 *= dscale; c1 *= dscale; c2 *= dscale;  // exact when dscale == 1
    const float u0 = u[idx], u1 = u[idx + n], u2 = u[idx + 2 * n];
    float a0, a1, a2;
    float z0 = 0.f, z1 = 0.f, z2 = 0.f;
    if (MODE == kSplitPadmm || STAT) {
      z0 = z[idx]; z1 = z[idx + n]; z2 = z[idx + 2 * n];
    }
    if (MODE == kSplitPadmm) {
      a0 = z0 + inv_nu * ((c0 - z0) + u0);
      a1 = z1 + inv_nu * ((c1 - z1) + u1);
      a2 = z2 + inv_nu * ((c2 - z2) + u2);
    } else {
      a0 = c0 + u0; a1 = c1 + u1; a2 = c2 + u2;
    }
    const float s = l21_scale(a0, a1, a2, thr);
    const float zn0 = s * a0, zn1 = s * a1, zn2 = s * a2;
    const float un0 = (u0 + c0) - zn0, un1 = (u1 + c1) - zn1, un2 = (u2 + c2) - zn2;
    z[idx] = zn0; z[idx + n] = zn1; z[idx + 2 * n] = zn2;
    u[idx] = un0; u[idx + n] = un1; u[idx + 2 * n] = un2;
    if (MODE == kSplitLadmm) {
      w[idx] = (c0 - zn0) + un0; w[idx + n] = (c1 - zn1) + un1; w[idx + 2 * n] = (c2 - zn2) + un2;
    } else if (MODE == kSplitPadmm) {
      w[idx] = 2.f * un0 - u0; w[idx + n] = 2.f * un1 - u1; w[idx + 2 * n] = 2.f * un2 - u2;
    }
    if (STAT) {
      const double p0 = (double)(c0 - zn0), p1 = (double)(c1 - zn1), p2 = (double)(c2 - zn2);
      const double e0 = (double)(zn0 - z0), e1 = (double)(zn1 - z1), e2 = (double)(zn2 - z2);
      acc_p += (p0 * p0 + p1 * p1) + p2 * p2;
      acc_d += (e0 * e0 + e1 * e1) + e2 * e2;
      acc_g += sqrt(((double)zn0 * zn0 + (double)zn1 * zn1) + (double)zn2 * zn2);
    }
  }
  if (STAT) {
    block_reduce_add(acc_p, stat);
    __syncthreads();
    block_reduce_add(acc_d, stat + 1);
    __syncthreads();
    block_reduce_add(acc_g, stat + 2);
  }
}

// Sinogram block of the split, g0 = 1/2 ||. - y||^2 with SquaredL2Loss.prox (scico/loss.py:220-226):
//   prox_{c g0}(v) = (c y + v) / (c + 1).
//   LADMM: z = prox(ax + u), c = nu;  PADMM: z = prox(z + inv_nu ((ax - z) + u)), c = 1/(rho nu).
//   u, w as in grad_prox_kernel.
//   STAT: stat[0] += ||ax - z_new||^2,  stat[1] += ||z_new - z_old||^2,  stat[2] += ||z_new - y||^2 (= 2 g0(z))
//   over the detector rows `rows` counts.
template <int MODE, bool STAT>
__global__ void __launch_bounds__(256)
sino_prox_kernel(size_t n, const float* __restrict__ ax, const float* __restrict__ y, float* __restrict__ z,
                 float* __restrict__ u, float* __restrict__ w, float c, float inv_nu, SinoRows rows,
                 double* __restrict__ stat) {
  const float rc1 = 1.0f / (c + 1.0f);
  double acc_p = 0.0, acc_d = 0.0, acc_g = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const float a = ax[idx], uo = u[idx], yy = y[idx];
    float v, zo = 0.f;
    if (MODE == kSplitPadmm || STAT) zo = z[idx];
    if (MODE == kSplitPadmm) {
      v = zo + inv_nu * ((a - zo) + uo);
    } else {
      v = a + uo;
    }
    const float zn = (c * yy + v) * rc1;
    const float un = (uo + a) - zn;
    z[idx] = zn;
    u[idx] = un;
    w[idx] = (MODE == kSplitPadmm) ? 2.f * un - uo : (a - zn) + un;
    if (STAT && rows.counted(idx)) {
      const double pr = (double)(a - zn), dz = (double)(zn - zo), r = (double)(zn - yy);
      acc_p += pr * pr;
      acc_d += dz * dz;
      acc_g += r * r;
    }
  }
  if (STAT) {
    block_reduce_add(acc_p, stat);
    __syncthreads();
    block_reduce_add(acc_d, stat + 1);
    __syncthreads();
    block_reduce_add(acc_g, stat + 2);
  }
}

// x = prox_f(x - step (atq + dscale D^T w)), f = 0 or the non-negativity indicator
// (LADMM: step = mu/nu, dscale = 1, _ladmm.py:270-271;  PADMM: step = 1/mu, dscale = alpha, _padmm.py:351-352).
__global__ void __launch_bounds__(256)
grad_primal_kernel(TvDims d, float* __restrict__ x, const float* __restrict__ atq, const float* __restrict__ w,
                   const float* __restrict__ lo_halo, float step, float dscale, int nonneg) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
  const float* wa = w;
  const float* wb = w + n;
  const float* wc = w + 2 * n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const Vox v = unravel(d, idx);
    float t0 = 0.f;
    if (!(d.last && v.i == d.n0 - 1)) t0 = -wa[idx];
    if (v.i > 0) t0 += wa[idx - plane];
    else if (!d.first && lo_halo) t0 += lo_halo[(size_t)v.j * d.n2 + v.k];
    float t1 = (v.j < d.n1 - 1) ? -wb[idx] : 0.f;
    if (v.j > 0) t1 += wb[idx - d.n2];
    float t2 = (v.k < d.n2 - 1) ? -wc[idx] : 0.f;
    if (v.k > 0) t2 += wc[idx - 1];
    const float ctw = atq[idx] + dscale * ((t0 + t1) + t2);
    float xn = x[idx] - step * ctw;
    if (nonneg) xn = fmaxf(xn, 0.f);
    x[idx] = xn;
  }
}

// ADMM right-hand side (_admmaux.py:231-255): rhs = A^T y + rho D^T (z - u);  *sumsq += ||rhs||^2.
//   lo_halo: plane (z - u)[0][-1] of the previous slab.
__global__ void __launch_bounds__(256)
admm_rhs_kernel(TvDims d, const float* __restrict__ aty, const float* __restrict__ z, const float* __restrict__ u,
                const float* __restrict__ lo_halo, float rho, float* __restrict__ rhs, double* sumsq) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
  double acc = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const Vox v = unravel(d, idx);
    float t0 = 0.f;
    if (!(d.last && v.i == d.n0 - 1)) t0 = -(z[idx] - u[idx]);
    if (v.i > 0) t0 += z[idx - plane] - u[idx - plane];
    else if (!d.first && lo_halo) t0 += lo_halo[(size_t)v.j * d.n2 + v.k];
    float t1 = (v.j < d.n1 - 1) ? -(z[idx + n] - u[idx + n]) : 0.f;
    if (v.j > 0) t1 += z[idx + n - d.n2] - u[idx + n - d.n2];
    float t2 = (v.k < d.n2 - 1) ? -(z[idx + 2 * n] - u[idx + 2 * n]) : 0.f;
    if (v.k > 0) t2 += z[idx + 2 * n - 1] - u[idx + 2 * n - 1];
    const float r = aty[idx] + rho * ((t0 + t1) + t2);
    rhs[idx] = r;
    acc += (double)r * (double)r;
  }
  block_reduce_add(acc, sumsq);
}

// (D^T D p)[idx], evaluated as D^T applied to the fp32-rounded differences (the reference applies the
// two operators one after the other, scico/linop/_linop.py:403-420).
//   lo_halo: plane p[-1] of the previous slab;  hi_halo: plane p[n0] of the next slab.
__device__ __forceinline__ float dtd_at(const TvDims& d, const float* __restrict__ p, const float* __restrict__ lo_halo,
                                        const float* __restrict__ hi_halo, size_t idx, const Vox& v, size_t plane) {
  const float pc = p[idx];
  float t0 = 0.f;
  if (v.i < d.n0 - 1) t0 = -(p[idx + plane] - pc);
  else if (!d.last && hi_halo) t0 = -(hi_halo[(size_t)v.j * d.n2 + v.k] - pc);
  if (v.i > 0) t0 += pc - p[idx - plane];
  else if (!d.first && lo_halo) t0 += pc - lo_halo[(size_t)v.j * d.n2 + v.k];
  float t1 = (v.j < d.n1 - 1) ? -(p[idx + d.n2] - pc) : 0.f;
  if (v.j > 0) t1 += pc - p[idx - d.n2];
  float t2 = (v.k < d.n2 - 1) ? -(p[idx + 1] - pc) : 0.f;
  if (v.k > 0) t2 += pc - p[idx - 1];
  return (t0 + t1) + t2;
}

// CG start (scico/solver.py:367-377): r = b - (rho D^T D x + atax);  p = r;  *num += r.r
__global__ void __launch_bounds__(256)
cg_init_kernel(TvDims d, const float* __restrict__ x, const float* __restrict__ lo_halo, const float* __restrict__ hi_halo,
               const float* __restrict__ atax, const float* __restrict__ b, float rho, float* __restrict__ r,
               float* __restrict__ p, double* num) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
  double acc = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const Vox v = unravel(d, idx);
    const float ax = rho * dtd_at(d, x, lo_halo, hi_halo, idx, v, plane) + atax[idx];
    const float rr = b[idx] - ax;
    r[idx] = rr;
    p[idx] = rr;
    acc += (double)rr * (double)rr;
  }
  block_reduce_add(acc, num);
}

// q = rho D^T D p + atap;  *pq += p.q;  *zero_me = 0 (ring slot of a later reduction, see xct_cg_lhs)
__global__ void __launch_bounds__(256)
cg_lhs_kernel(TvDims d, const float* __restrict__ p, const float* __restrict__ lo_halo, const float* __restrict__ hi_halo,
              const float* __restrict__ atap, float rho, float* __restrict__ q, double* pq, double* zero_me) {
  const size_t plane = (size_t)d.n1 * d.n2, n = plane * d.n0;
  if (zero_me && blockIdx.x == 0 && threadIdx.x == 0) *zero_me = 0.0;
  double acc = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const Vox v = unravel(d, idx);
    const float qq = rho * dtd_at(d, p, lo_halo, hi_halo, idx, v, plane) + atap[idx];
    q[idx] = qq;
    acc += (double)p[idx] * (double)qq;
  }
  block_reduce_add(acc, pq);
}

// alpha = num / pq;  x += alpha p;  r -= alpha q;  *num_new += r.r   (scico/solver.py:392-398)
__global__ void __launch_bounds__(256)
cg_xr_kernel(size_t n, float* __restrict__ x, float* __restrict__ r, const float* __restrict__ p,
             const float* __restrict__ q, const double* num, const double* pq, double* num_new, double* zero_me) {
  const float alpha = (float)(*num / *pq);
  if (zero_me && blockIdx.x == 0 && threadIdx.x == 0) *zero_me = 0.0;
  double acc = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    x[idx] = x[idx] + alpha * p[idx];
    const float rr = r[idx] - alpha * q[idx];
    r[idx] = rr;
    acc += (double)rr * (double)rr;
  }
  block_reduce_add(acc, num_new);
}

// beta = num_new / num;  p = r + beta p   (scico/solver.py:399-402)
__global__ void __launch_bounds__(256)
cg_p_kernel(size_t n, float* __restrict__ p, const float* __restrict__ r, const double* num, const double* num_new) {
  const float beta = (float)(*num_new / *num);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x)
    p[idx] = r[idx] + beta * p[idx];
}

}  // namespace xct
