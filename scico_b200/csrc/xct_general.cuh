// General-geometry kernels: correct for ANY set of 2x4 matrices (3D) / any view table (2D).
// They are the fallback when a plan is outside the plane kernels' envelope (non-separable 3D
// matrices, projected pixel wider than the staged window, ...).  Thread-per-voxel: the adjoint
// gathers 4 (2) taps per view straight from the L1/L2-cached sinogram, the forward scatters with
// RED.ADD.F32.  Same fp32 expression trees as the oracle (see xct_geom.cuh).
#pragma once
#include "xct_geom.cuh"

namespace xct {

struct Taps3 {
  int r0, c0;
  float w[4];  // ul, ur, ll, lr (ur = row + 1, ll = col + 1: _xray3d.py:155-158), masked
};

// _xray3d.py:211-264 for one voxel.  M: 8 floats (2x4 row-major).  Rows are tested against the
// GLOBAL detector (rows_total) and against the local slab [row_off, row_off + D0).
__device__ __forceinline__ void taps3d(const float* __restrict__ M, float xi, float xj, float xk, int D0,
                                       int D1, int row_off, int rows_total, Taps3& o) {
  const float4 m0 = __ldg(reinterpret_cast<const float4*>(M));
  const float4 m1 = __ldg(reinterpret_cast<const float4*>(M) + 1);
  const float P0 = __fadd_rn(
      __fadd_rn(__fadd_rn(__fmul_rn(m0.x, xi), __fmul_rn(m0.y, xj)), __fmul_rn(m0.z, xk)), m0.w);
  const float P1 = __fadd_rn(
      __fadd_rn(__fadd_rn(__fmul_rn(m1.x, xi), __fmul_rn(m1.y, xj)), __fmul_rn(m1.z, xk)), m1.w);
  const float l0 = __fadd_rn(P0, -0.25f), l1 = __fadd_rn(P1, -0.25f);
  const float t0 = fminf(__fadd_rn(ceilf(l0), -l0), 0.5f);
  const float t1 = fminf(__fadd_rn(ceilf(l1), -l1), 0.5f);
  const int rg = __float2int_rd(l0);
  o.c0 = __float2int_rd(l1);
  o.r0 = rg - row_off;
  const float u0 = __fadd_rn(0.5f, -t0), u1 = __fadd_rn(0.5f, -t1);
  o.w[0] = __fmul_rn(__fmul_rn(t0, t1), 4.0f);
  o.w[1] = __fmul_rn(__fmul_rn(u0, t1), 4.0f);
  o.w[2] = __fmul_rn(__fmul_rn(t0, u1), 4.0f);
  o.w[3] = __fmul_rn(__fmul_rn(u0, u1), 4.0f);
  const bool rin0 = rg >= 0 && rg < rows_total && o.r0 >= 0 && o.r0 < D0;
  const bool rin1 = rg + 1 >= 0 && rg + 1 < rows_total && o.r0 + 1 >= 0 && o.r0 + 1 < D0;
  const bool cin0 = o.c0 >= 0 && o.c0 < D1, cin1 = o.c0 + 1 >= 0 && o.c0 + 1 < D1;
  if (!(rin0 && cin0)) o.w[0] = 0.f;
  if (!(rin1 && cin0)) o.w[1] = 0.f;
  if (!(rin0 && cin1)) o.w[2] = 0.f;
  if (!(rin1 && cin1)) o.w[3] = 0.f;
}

struct Gen3Params {
  const float* mats;  // device (V,2,4)
  int V, N0, N1, N2, D0, D1;
  int slice_offset, row_off, rows_total;
};

__device__ __forceinline__ float voxel_coord(int idx, int offset) {
  return __fadd_rn((float)idx + 0.5f, (float)offset);  // mgrid + 0.5 (+ slice_offset), _xray3d.py:211-212
}

// thread = one voxel (k fastest); views in ascending order, taps in the reference's order.
// ROUTE: add into the slab owners' memory (OutRoute, xct_geom.cuh) instead of storing to `vol`.
template <bool ROUTE = false>
__global__ void __launch_bounds__(256)
gen3d_adjoint_kernel(Gen3Params p, const float* __restrict__ sino, float* __restrict__ vol,
                     const __grid_constant__ OutRoute route) {
  const size_t n = (size_t)p.N0 * p.N1 * p.N2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % p.N2);
    const size_t ij = idx / p.N2;
    const int j = (int)(ij % p.N1), i = (int)(ij / p.N1);
    const float xi = voxel_coord(i, p.slice_offset), xj = voxel_coord(j, 0), xk = voxel_coord(k, 0);
    float acc = 0.f;
    for (int v = 0; v < p.V; ++v) {
      Taps3 t;
      taps3d(p.mats + 8 * (size_t)v, xi, xj, xk, p.D0, p.D1, p.row_off, p.rows_total, t);
      const float* y = sino + (size_t)v * p.D0 * p.D1;
      const size_t base = (size_t)((long long)t.r0 * p.D1 + t.c0);
      if (t.w[0] != 0.f) acc = fmaf(__ldg(y + base), t.w[0], acc);
      if (t.w[1] != 0.f) acc = fmaf(__ldg(y + base + p.D1), t.w[1], acc);
      if (t.w[2] != 0.f) acc = fmaf(__ldg(y + base + 1), t.w[2], acc);
      if (t.w[3] != 0.f) acc = fmaf(__ldg(y + base + p.D1 + 1), t.w[3], acc);
    }
    if constexpr (ROUTE) route_add(route, i, (long long)j * p.N2 + k, acc);
    else vol[idx] = acc;
  }
}

// thread = one voxel; scatter with RED.  `sino` must be zero on entry.
__global__ void __launch_bounds__(256)
gen3d_forward_kernel(Gen3Params p, const float* __restrict__ vol, float* __restrict__ sino) {
  const size_t n = (size_t)p.N0 * p.N1 * p.N2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % p.N2);
    const size_t ij = idx / p.N2;
    const int j = (int)(ij % p.N1), i = (int)(ij / p.N1);
    const float xi = voxel_coord(i, p.slice_offset), xj = voxel_coord(j, 0), xk = voxel_coord(k, 0);
    const float val = vol[idx];
    if (val == 0.f) continue;
    for (int v = 0; v < p.V; ++v) {
      Taps3 t;
      taps3d(p.mats + 8 * (size_t)v, xi, xj, xk, p.D0, p.D1, p.row_off, p.rows_total, t);
      float* y = sino + (size_t)v * p.D0 * p.D1;
      const size_t base = (size_t)((long long)t.r0 * p.D1 + t.c0);
      if (t.w[0] != 0.f) atomicAdd(y + base, t.w[0] * val);
      if (t.w[1] != 0.f) atomicAdd(y + base + p.D1, t.w[1] * val);
      if (t.w[2] != 0.f) atomicAdd(y + base + 1, t.w[2] * val);
      if (t.w[3] != 0.f) atomicAdd(y + base + p.D1 + 1, t.w[3] * val);
    }
  }
}

__global__ void __launch_bounds__(256)
gen3d_weights_kernel(Gen3Params p, int view, int32_t* __restrict__ ul, float* __restrict__ w) {
  const size_t n = (size_t)p.N0 * p.N1 * p.N2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % p.N2);
    const size_t ij = idx / p.N2;
    const int j = (int)(ij % p.N1), i = (int)(ij / p.N1);
    Taps3 t;
    taps3d(p.mats + 8 * (size_t)view, voxel_coord(i, p.slice_offset), voxel_coord(j, 0),
           voxel_coord(k, 0), p.D0, p.D1, p.row_off, p.rows_total, t);
    ul[idx] = t.r0 + p.row_off;
    ul[n + idx] = t.c0;
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q * n + idx] = t.w[q];
  }
}

// ---------------------------------------------------------------------------------- 2D
struct Gen2Params {
  const ViewRec* views;
  int V, N0, N1, ny, batch;
};

// thread = one pixel of one batch item; gathers both taps for every view.
template <bool ROUTE = false>  // ROUTE: one image, rows added into the row-block owners' memory
__global__ void __launch_bounds__(256)
gen2d_adjoint_kernel(Gen2Params p, const float* __restrict__ sino, float* __restrict__ im,
                     const __grid_constant__ OutRoute route) {
  const size_t npix = (size_t)p.N0 * p.N1, n = npix * p.batch;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int bt = (int)(idx / npix);
    const size_t pix = idx % npix;
    const int j = (int)(pix % p.N1), i = (int)(pix / p.N1);
    float s0 = 0.f, s1 = 0.f;  // two separate sums, added at the end (_xray2d.py:298-304)
    for (int v = 0; v < p.V; ++v) {
      const ViewRec vr = p.views[v];
      const float u = Geom2::combine(vr, Geom2::hoistA(vr, i), Geom2::hoistB(vr, j));
      int c;
      float w0, w1;
      Geom2::bins(vr, u, c, w0, w1);
      const float* y = sino + ((size_t)bt * p.V + v) * p.ny;
      if (c >= 0 && c < p.ny) s0 = fmaf(__ldg(y + c), w0, s0);
      if (c + 1 >= 0 && c + 1 < p.ny) s1 = fmaf(__ldg(y + c + 1), w1, s1);
    }
    if constexpr (ROUTE) route_add(route, i, j, s0 + s1);
    else im[idx] = s0 + s1;
  }
}

__global__ void __launch_bounds__(256)
gen2d_forward_kernel(Gen2Params p, const float* __restrict__ im, float* __restrict__ sino) {
  const size_t npix = (size_t)p.N0 * p.N1, n = npix * p.batch;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int bt = (int)(idx / npix);
    const size_t pix = idx % npix;
    const int j = (int)(pix % p.N1), i = (int)(pix / p.N1);
    const float val = im[idx];
    if (val == 0.f) continue;
    for (int v = 0; v < p.V; ++v) {
      const ViewRec vr = p.views[v];
      const float u = Geom2::combine(vr, Geom2::hoistA(vr, i), Geom2::hoistB(vr, j));
      int c;
      float w0, w1;
      Geom2::bins(vr, u, c, w0, w1);
      float* y = sino + ((size_t)bt * p.V + v) * p.ny;
      if (c >= 0 && c < p.ny) atomicAdd(y + c, val * w0);
      if (c + 1 >= 0 && c + 1 < p.ny) atomicAdd(y + c + 1, val * w1);
    }
  }
}

__global__ void __launch_bounds__(256)
gen2d_weights_kernel(Gen2Params p, int view, int32_t* __restrict__ inds, float* __restrict__ w) {
  const size_t n = (size_t)p.N0 * p.N1;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % p.N1), i = (int)(idx / p.N1);
    const ViewRec vr = p.views[view];
    const float u = Geom2::combine(vr, Geom2::hoistA(vr, i), Geom2::hoistB(vr, j));
    int c;
    float w0, w1;
    Geom2::bins(vr, u, c, w0, w1);
    inds[idx] = c;
    w[idx] = w0;
  }
}

// weights as the separable 3D path computes them (row record x column bins), for the test hook
__global__ void __launch_bounds__(256)
sep3d_weights_kernel(const ViewRec* __restrict__ views, const RowRec* __restrict__ rows, int view, int N0,
                     int N1, int N2, int D1, int row_off, int32_t* __restrict__ ul, float* __restrict__ w) {
  const size_t n = (size_t)N0 * N1 * N2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % N2);
    const size_t ij = idx / N2;
    const int j = (int)(ij % N1), i = (int)(ij / N1);
    const ViewRec vr = views[view];
    const RowRec rr = rows[(size_t)view * N0 + i];
    const float u = Geom3::combine(vr, Geom3::hoistA(vr, j), Geom3::hoistB(vr, k));
    int c;
    float w0, w1;
    Geom3::bins(vr, u, c, w0, w1);
    const bool cin0 = c >= 0 && c < D1, cin1 = c + 1 >= 0 && c + 1 < D1;
    ul[idx] = rr.r0 + row_off;
    ul[n + idx] = c;
    w[0 * n + idx] = cin0 ? rr.w0 * w0 : 0.f;
    w[1 * n + idx] = cin0 ? rr.w1 * w0 : 0.f;
    w[2 * n + idx] = cin1 ? rr.w0 * w1 : 0.f;
    w[3 * n + idx] = cin1 ? rr.w1 * w1 : 0.f;
  }
}

}  // namespace xct
