// Brick kernels: XRayTransform3D with GENERAL 2x4 matrices (both detector coordinates depend on all three
// voxel indices: tilted rotation axes, the 74-degree XY tilt of examples/scripts/ct_projector_comparison_3d.py).
// Reference semantics: scico/linop/xray/_xray3d.py:141-204 (_project / _back_project), :206-266 (_calc_weights).
//
// Nothing factorises over slices here, so the kernels tile the VOLUME into small bricks, one brick per warp,
// and exploit what is left: a brick of B^3 voxels projects, in any view, into a detector window of about
// (sqrt(3) B + 3)^2 bins -- a few hundred floats that fit in a warp-private shared-memory window.
//
//   adjoint : the warp's 8 x 8 x 8 voxels are register accumulators (16 per lane) for ALL views, in order.
//             Per view ONE TMA box (cp.async.bulk.tensor.3d, WR x WC bins of the (V, D0, D1) sinogram, issued
//             by one elected lane, completing on a per-(warp, stage) mbarrier, hardware zero fill outside the
//             detector = the reference's zero weights for out-of-range taps; first column a multiple of 4: the
//             innermost TMA coordinate must be 16-byte aligned) lands in a STAGES-deep ring; the
//             four taps of a voxel are four LDS off one address register.  No global gather, no bounds test.
//   forward : the warp's 512 voxels are register-stationary for every view of the launch; lanes sit on a
//             2-voxel lattice across the two volume axes that are most perpendicular to the rays, so that the
//             bins the 32 lanes of one instruction touch are distinct (checked per view on the host, exactly,
//             for all lane pairs) and plain shared-memory read-modify-writes are race free; views that fail
//             the check take the same kernel with shared-memory atomics.  The window is flushed once per
//             (view, brick) with vector RED (red.global.add.v4.f32) when detector rows are 16-byte aligned.
//
// Coordinates: the reference's expression tree (xct_general.cuh::taps3d) term by term, with explicit
// round-to-nearest products and sums; the two detector coordinates of a voxel share packed FADD2 sums (the
// products stay scalar: ptxas contracts a PACKED product into a packed sum, which would change bin indices).
// The window origin is the exact minimum of the brick's coordinates: every rounding step is monotone, so the
// minimum is attained at the corner picked by the signs of the three coefficients.
#pragma once
#include <cuda.h>

#include "xct_general.cuh"
#include "xct_plane2.cuh"  // mbarrier / TMA / cp.async helpers

namespace xct {

struct BrickParams {
  const float* mats;     // device (V, 4, 2): the 2x4 matrices TRANSPOSED, so that the two rows' coefficients of one
                         // axis (and the two offsets) are adjacent = one aligned register pair for the packed sums
  const int* view_list;  // forward: views of this launch's class; nullptr = 0..n_list-1
  int n_list;            // views to process
  int V;                 // views held by the plan
  int N0, N1, N2, D0, D1;
  int slice_offset, row_off;
  int nb0, nb1, nb2;     // brick grid
  int views_per_chunk;   // small problems: views per blockIdx.y
};

struct Mat24 {
  float4 r0, r1;  // rows of the 2x4 matrix
  float2 off;     // (m03, m13) as loaded: an aligned pair
};
__device__ __forceinline__ Mat24 load_mat(const float* mats, int v) {
  const float4* q = reinterpret_cast<const float4*>(mats) + 2 * (size_t)v;
  const float4 lo = __ldg(q), hi = __ldg(q + 1);  // (m00, m10, m01, m11), (m02, m12, m03, m13)
  Mat24 m;
  m.r0 = make_float4(lo.x, lo.z, hi.x, hi.z);
  m.r1 = make_float4(lo.y, lo.w, hi.y, hi.w);
  m.off = make_float2(hi.z, hi.w);
  return m;
}

// left edge of the footprint, one coordinate: ((m_0 x_i + m_1 x_j) + m_2 x_k) + m_3 - 0.25  (_xray3d.py:216,223)
__device__ __forceinline__ float left_edge(const float4 m, float xi, float xj, float xk) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m.x, xi), __fmul_rn(m.y, xj)), __fmul_rn(m.z, xk)), m.w), -0.25f);
}
// floor of the smallest left edge over a brick [lo, hi] per axis (voxel-centre coordinates): exact, see header
__device__ __forceinline__ int brick_origin(const float4 m, float xi_lo, float xi_hi, float xj_lo, float xj_hi, float xk_lo,
                                            float xk_hi) {
  return __float2int_rd(left_edge(m, m.x >= 0.f ? xi_lo : xi_hi, m.y >= 0.f ? xj_lo : xj_hi, m.z >= 0.f ? xk_lo : xk_hi));
}

// bins and the two 1D weights of both coordinates of one voxel.  l = (l0, l1) left edges.
//   r = floor(l0), c = floor(l1); t = min(ceil(l) - l, 0.5) (0 when l is an integer, _xray3d.py:224); u = 0.5 - t
// The conversion pipe (XU: 16 lanes per clock and SM) takes three operations per voxel here -- two floors and the
// row coordinate's ceil -- which leaves it at ~3/4 of the issue time; the column coordinate's ceil is rebuilt from
// its floor on the ALU / FMA pipes instead (I2FP + compare + add): fl + 1 is exact for |l| < 2^24, so
// (fl + 1) - l == ceil(l) - l bit for bit when l is not an integer, and the select restores the 0 of the integer case.
__device__ __forceinline__ void bins3(float2 l, int& r, int& c, float2& t, float2& u) {
  r = __float2int_rd(l.x);
  c = __float2int_rd(l.y);
  const float fy = __int2float_rn(c);
  const float2 d = __fadd2_rn(make_float2(ceilf(l.x), __fadd_rn(fy, 1.0f)), make_float2(-l.x, -l.y));
  t.x = fminf(d.x, 0.5f);
  t.y = fy == l.y ? 0.f : fminf(d.y, 0.5f);
  u = __fadd2_rn(make_float2(0.5f, 0.5f), make_float2(-t.x, -t.y));
}

// shared-memory load at a 32-bit shared address plus an immediate byte offset (not volatile: the window is
// read-only between the stage's arrival and the __syncwarp that precedes its refill, so ptxas may schedule it)
template <int OFF>
__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
  return v;
}

__device__ __forceinline__ void cp_async4_zfill(float* smem_dst, const float* gmem_src, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}

// ------------------------------------------------------------------------------------------ adjoint
// Brick 8 (axis 0) x 8 (axis 1) x 8 (axis 2) per warp.  Lane (li, lj) = (lane >> 3, lane & 7) owns the voxels
// (i0 + li + 4 p, j0 + lj, k0 + n), p = 0..1, n = 0..7: 16 accumulators.  Window WR x WC floats per stage.
// TMA: one cp.async.bulk.tensor.3d box per (view, brick); otherwise (detector rows not 16-byte aligned) the
// lanes stage the window with 4-byte cp.async (LDGSTS, zero fill by src-size 0), same ring, same consumer.
// ROUTE: result slices go to the row-block owners (view-block sharding, xct_adjoint_scatter).
constexpr int kBrick = 8;

template <int WR, int WC, int STAGES, int WARPS, bool TMA, bool ROUTE>
__global__ void __launch_bounds__(WARPS * 32)
brick_adjoint_kernel(BrickParams p, const float* __restrict__ sino, float* __restrict__ vol,
                     const __grid_constant__ CUtensorMap tmap, const __grid_constant__ OutRoute route) {
  constexpr int B = kBrick;
  constexpr int STAGE_FLOATS = ((WR * WC + 31) / 32) * 32;  // 128-byte multiple: TMA destination alignment
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ntasks = (long long)p.nb0 * p.nb1 * p.nb2;
  long long task = (long long)blockIdx.x * WARPS + warp;
  if (task >= ntasks) return;  // warp-uniform; no block barrier below
  const int bk = (int)(task % p.nb2);
  task /= p.nb2;
  const int bj = (int)(task % p.nb1);
  const int bi = (int)(task / p.nb1);
  const int i0 = bi * B, j0 = bj * B, k0 = bk * B;
  const int li = lane >> 3, lj = lane & 7;

  float* ring = smem + (size_t)warp * (STAGES * STAGE_FLOATS);
  const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring);
  const unsigned bars_sa = (unsigned)__cvta_generic_to_shared(smem + (size_t)WARPS * (STAGES * STAGE_FLOATS)) + warp * STAGES * 8u;
  if (TMA) {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) mbar_init(bars_sa + 8u * s, 1);
      mbar_fence_init();
    }
    __syncwarp();
  }

  // voxel-centre coordinates (indices clamped into the volume: out-of-range voxels of an edge brick compute a
  // valid neighbour's taps and are not stored)
  float xi[2], xk[B];
#pragma unroll
  for (int q = 0; q < 2; ++q) xi[q] = voxel_coord(min(i0 + li + 4 * q, p.N0 - 1), p.slice_offset);
  const float xj = voxel_coord(min(j0 + lj, p.N1 - 1), 0);
#pragma unroll
  for (int n = 0; n < B; ++n) xk[n] = voxel_coord(min(k0 + n, p.N2 - 1), 0);
  // brick extent for the window origin
  const float xi_lo = voxel_coord(i0, p.slice_offset), xi_hi = voxel_coord(min(i0 + B, p.N0) - 1, p.slice_offset);
  const float xj_lo = voxel_coord(j0, 0), xj_hi = voxel_coord(min(j0 + B, p.N1) - 1, 0);
  const float xk_lo = voxel_coord(k0, 0), xk_hi = voxel_coord(min(k0 + B, p.N2) - 1, 0);

  float acc[2][B];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int n = 0; n < B; ++n) acc[q][n] = 0.f;

  const int v_begin = blockIdx.y * p.views_per_chunk;
  const int v_end = min(p.n_list, v_begin + p.views_per_chunk);

  // producer: window of view v into `stage`; returns the packed origin (global row rb, column cb)
  auto fetch = [&](int v, int stage, int& rb, int& cb) {
    const Mat24 m = load_mat(p.mats, v);
    rb = brick_origin(m.r0, xi_lo, xi_hi, xj_lo, xj_hi, xk_lo, xk_hi);
    // the innermost TMA coordinate must address a 16-byte aligned element (measured: tools/tma_probe.cu, a box
    // starting at column 5 raises "illegal instruction"): the window starts at a multiple of 4 columns
    cb = brick_origin(m.r1, xi_lo, xi_hi, xj_lo, xj_hi, xk_lo, xk_hi) & ~3;
    if (TMA) {
      if (elect_one()) {
        const unsigned bar = bars_sa + 8u * stage;
        mbar_expect_tx(bar, WR * WC * (unsigned)sizeof(float));
        tma_load_3d(ring_sa + stage * (STAGE_FLOATS * (unsigned)sizeof(float)), &tmap, cb, rb - p.row_off, v, bar);
      }
    } else {
      float* zb = ring + stage * STAGE_FLOATS;
      const float* y = sino + (size_t)v * p.D0 * (size_t)p.D1;
      for (int e = lane; e < WR * WC; e += 32) {
        const int wr = e / WC, wc = e - wr * WC;
        const int row = rb - p.row_off + wr, col = cb + wc;
        const bool ok = (unsigned)row < (unsigned)p.D0 && (unsigned)col < (unsigned)p.D1;
        cp_async4_zfill(zb + e, y + (ok ? (size_t)row * p.D1 + col : 0), ok);
      }
    }
  };

  int rbq[STAGES], cbq[STAGES];
#pragma unroll
  for (int s = 0; s < STAGES; ++s) rbq[s] = cbq[s] = 0;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (v_begin + s < v_end) fetch(v_begin + s, s, rbq[s], cbq[s]);
    if (!TMA) cp_async_commit();
  }

  int st_c = 0, st_p = STAGES - 1;
  unsigned par = 0;
  for (int v = v_begin; v < v_end; ++v) {
    __syncwarp();  // every lane has finished reading the stage that is refilled now
    if (v + STAGES - 1 < v_end) fetch(v + STAGES - 1, st_p, rbq[STAGES - 1], cbq[STAGES - 1]);
    if (TMA) {
      mbar_wait(bars_sa + 8u * st_c, par);
    } else {
      cp_async_commit();
      cp_async_wait<STAGES - 1>();
      __syncwarp();
    }
    const float* zb = ring + st_c * STAGE_FLOATS;
    st_p = st_c;
    if (++st_c == STAGES) { st_c = 0; par ^= 1u; }
    const int rb = rbq[0], cb = cbq[0];
#pragma unroll
    for (int s = 0; s + 1 < STAGES; ++s) { rbq[s] = rbq[s + 1]; cbq[s] = cbq[s + 1]; }

    const Mat24 m = load_mat(p.mats, v);
    // bin (r, c) of this view sits at byte zbase + 4 * (r * WC + c) of the shared window (32-bit shared
    // address arithmetic: one IMAD + one LEA per voxel, the four taps are immediate offsets)
    const unsigned zbase = (unsigned)__cvta_generic_to_shared(zb) - 4u * (unsigned)(rb * WC + cb);
    const float2 off = m.off;
    const float b0 = __fmul_rn(m.r0.y, xj), b1 = __fmul_rn(m.r1.y, xj);
    float2 ab[2];
#pragma unroll
    for (int q = 0; q < 2; ++q)
      ab[q] = make_float2(__fadd_rn(__fmul_rn(m.r0.x, xi[q]), b0), __fadd_rn(__fmul_rn(m.r1.x, xi[q]), b1));
#pragma unroll
    for (int n = 0; n < B; ++n) {
      const float2 ck = make_float2(__fmul_rn(m.r0.z, xk[n]), __fmul_rn(m.r1.z, xk[n]));
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float2 l = __fadd2_rn(__fadd2_rn(__fadd2_rn(ab[q], ck), off), make_float2(-0.25f, -0.25f));
        int r, c;
        float2 t, u;
        bins3(l, r, c, t, u);
        const unsigned za = zbase + 4u * (unsigned)(r * WC + c);
        // taps: ul (r, c), ur (r + 1, c), ll (r, c + 1), lr (r + 1, c + 1)  (_xray3d.py:155-158, 200-203);
        // weights (t0 t1, u0 t1, t0 u1, u0 u1) * 4 applied as t1 (t0 y_ul + u0 y_ur) + u1 (t0 y_ll + u0 y_lr),
        // the common factor 4 = 1 / w^2 at the end (a power of two: exact)
        const float s0 = fmaf(u.x, lds_f32<4 * WC>(za), __fmul_rn(t.x, lds_f32<0>(za)));
        const float s1 = fmaf(u.x, lds_f32<4 * WC + 4>(za), __fmul_rn(t.x, lds_f32<4>(za)));
        acc[q][n] = fmaf(u.y, s1, fmaf(t.y, s0, acc[q][n]));
      }
    }
  }
  if (!TMA) cp_async_wait<0>();

  const int j = j0 + lj;
  if (j >= p.N1) return;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = i0 + li + 4 * q;
    if (i >= p.N0) continue;
    if constexpr (ROUTE) {
      RouteCursor cur;
      cur.seek(route, i, (long long)j * p.N2 + k0);
      if (route.store && k0 + B <= p.N2 && (p.N2 & 3) == 0 && (reinterpret_cast<uintptr_t>(cur.q) & 15) == 0) {
        // 32 contiguous bytes per lane as two 16-byte stores: scalar stores would cross NVLink as 4 useful
        // bytes per 32-byte sector (measured: the fused exchange of the tilted case was 0.4 ms slower than NCCL)
        reinterpret_cast<float4*>(cur.q)[0] = make_float4(4.0f * acc[q][0], 4.0f * acc[q][1], 4.0f * acc[q][2], 4.0f * acc[q][3]);
        reinterpret_cast<float4*>(cur.q)[1] = make_float4(4.0f * acc[q][4], 4.0f * acc[q][5], 4.0f * acc[q][6], 4.0f * acc[q][7]);
      } else {
#pragma unroll
        for (int n = 0; n < B; ++n) {
          if (k0 + n >= p.N2) break;
          if (route.store) cur.q[n] = 4.0f * acc[q][n];
          else atomicAdd_system(cur.q + n, 4.0f * acc[q][n]);
        }
      }
    } else {
      float* o = vol + ((size_t)i * p.N1 + j) * (size_t)p.N2 + k0;
      if (gridDim.y > 1) {
#pragma unroll
        for (int n = 0; n < B; ++n)
          if (k0 + n < p.N2) atomicAdd(o + n, 4.0f * acc[q][n]);
      } else if (k0 + B <= p.N2 && (p.N2 & 3) == 0 && (reinterpret_cast<uintptr_t>(vol) & 15) == 0) {
        reinterpret_cast<float4*>(o)[0] = make_float4(4.0f * acc[q][0], 4.0f * acc[q][1], 4.0f * acc[q][2], 4.0f * acc[q][3]);
        reinterpret_cast<float4*>(o)[1] = make_float4(4.0f * acc[q][4], 4.0f * acc[q][5], 4.0f * acc[q][6], 4.0f * acc[q][7]);
      } else {
#pragma unroll
        for (int n = 0; n < B; ++n)
          if (k0 + n < p.N2) o[n] = 4.0f * acc[q][n];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ forward
// DEPTH: the volume axis most parallel to the rays in this launch's views (smallest projected length).
// Lanes sit on a 2-voxel lattice over the other two axes (a: 8 lanes, b: 4 lanes -- 17 % fewer shared-memory
// bank-conflict wavefronts than 4 x 8 for the tilted test geometry); a lane owns the voxels
// (a0 + 2 la + da, b0 + 2 lb + db, c0 + n), da, db = 0..1, n = 0..3: brick 16 (a) x 8 (b) x 4 (depth).
//   DEPTH 2: (a, b, c) = (axis 0, axis 1, axis 2);  DEPTH 1: (axis 0, axis 2, axis 1);  DEPTH 0: (axis 1, axis 2, axis 0)
// ATOMIC: shared-memory atomics instead of plain read-modify-write (views whose lane lattice can put two
// lanes of one instruction on the same bin; decided per view on the host).
// VEC4: detector rows are 16-byte aligned (D1 % 4 == 0, aligned pointer): window columns start at a multiple of
// 4 and are flushed with red.global.add.v4.f32.
struct BrickFwdGeom {
  static constexpr int LA = 8, LB = 4, NC = 4;
  static constexpr int EA = 2 * LA, EB = 2 * LB, EC = NC;  // brick extent along (a, b, c)
};

template <int DEPTH, int WR, int WC, int WARPS, bool ATOMIC, bool VEC4>
__global__ void __launch_bounds__(WARPS * 32)
brick_forward_kernel(BrickParams p, const float* __restrict__ vol, float* __restrict__ sino) {
  using BG = BrickFwdGeom;
  static_assert(WC % 4 == 0, "window rows are flushed as float4 groups");
  constexpr int AX_A = DEPTH == 0 ? 1 : 0, AX_B = DEPTH == 2 ? 1 : 2, AX_C = DEPTH;
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ntasks = (long long)p.nb0 * p.nb1 * p.nb2;  // brick grid over (a, b, c)
  long long task = (long long)blockIdx.x * WARPS + warp;
  if (task >= ntasks) return;
  const int tc = (int)(task % p.nb2);
  task /= p.nb2;
  const int tb = (int)(task % p.nb1);
  const int ta = (int)(task / p.nb1);
  const int a0 = ta * BG::EA, b0 = tb * BG::EB, c0 = tc * BG::EC;
  const int la = lane / BG::LB, lb = lane % BG::LB;
  const int dims[3] = {p.N0, p.N1, p.N2};
  const int NA = dims[AX_A], NB = dims[AX_B], NCd = dims[AX_C];
  float* win = smem + (size_t)warp * (WR * WC);

  auto coord = [&](int axis, int idx) { return voxel_coord(idx, axis == 0 ? p.slice_offset : 0); };
  auto vox_ptr = [&](int a, int b, int c) {
    int ijk[3];
    ijk[AX_A] = a; ijk[AX_B] = b; ijk[AX_C] = c;
    return vol + ((size_t)ijk[0] * p.N1 + ijk[1]) * (size_t)p.N2 + ijk[2];
  };

  // register-stationary voxels (pre-multiplied by 1 / w^2 = 4, a power of two: bit-identical to scaling the
  // weights) and their coordinates (clamped indices: out-of-range voxels hold 0 and land on a neighbour's bins)
  float x[2][2][BG::NC];
  float xa[2], xb[2], xc[BG::NC];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    xa[d] = coord(AX_A, min(a0 + 2 * la + d, NA - 1));
    xb[d] = coord(AX_B, min(b0 + 2 * lb + d, NB - 1));
  }
#pragma unroll
  for (int n = 0; n < BG::NC; ++n) xc[n] = coord(AX_C, min(c0 + n, NCd - 1));
#pragma unroll
  for (int da = 0; da < 2; ++da)
#pragma unroll
    for (int db = 0; db < 2; ++db)
#pragma unroll
      for (int n = 0; n < BG::NC; ++n) {
        const int a = a0 + 2 * la + da, b = b0 + 2 * lb + db, c = c0 + n;
        x[da][db][n] = (a < NA && b < NB && c < NCd) ? 4.0f * __ldg(vox_ptr(a, b, c)) : 0.f;
      }
  const float xa_lo = coord(AX_A, a0), xa_hi = coord(AX_A, min(a0 + BG::EA, NA) - 1);
  const float xb_lo = coord(AX_B, b0), xb_hi = coord(AX_B, min(b0 + BG::EB, NB) - 1);
  const float xc_lo = coord(AX_C, c0), xc_hi = coord(AX_C, min(c0 + BG::EC, NCd) - 1);

  auto comp = [](const float4 m, int axis) { return axis == 0 ? m.x : (axis == 1 ? m.y : m.z); };

  const int v_begin = blockIdx.y * p.views_per_chunk;
  const int v_end = min(p.n_list, v_begin + p.views_per_chunk);
  for (int vi = v_begin; vi < v_end; ++vi) {
    const int v = p.view_list ? __ldg(p.view_list + vi) : vi;
    const Mat24 m = load_mat(p.mats, v);
    // coefficient of axes (a, b, c) in both rows; the reference's sum order is axis 0, 1, 2
    const float ma0 = comp(m.r0, AX_A), mb0 = comp(m.r0, AX_B), mc0 = comp(m.r0, AX_C);
    const float ma1 = comp(m.r1, AX_A), mb1 = comp(m.r1, AX_B), mc1 = comp(m.r1, AX_C);
    auto origin = [&](float ma, float mb, float mc, float4 row) {
      float xyz[3];
      xyz[AX_A] = ma >= 0.f ? xa_lo : xa_hi;
      xyz[AX_B] = mb >= 0.f ? xb_lo : xb_hi;
      xyz[AX_C] = mc >= 0.f ? xc_lo : xc_hi;
      return __float2int_rd(left_edge(row, xyz[0], xyz[1], xyz[2]));
    };
    const int rb = origin(ma0, mb0, mc0, m.r0);
    int cb = origin(ma1, mb1, mc1, m.r1);
    if (VEC4) cb &= ~3;

    for (int e = lane; e < WR * WC / 4; e += 32) reinterpret_cast<float4*>(win)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    float* w0 = win - (rb * WC + cb);  // w0[r * WC + c] is bin (r, c)
    const float2 off = m.off;
    float2 pa[2], pb[2], pc[BG::NC];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      pa[d] = make_float2(__fmul_rn(ma0, xa[d]), __fmul_rn(ma1, xa[d]));
      pb[d] = make_float2(__fmul_rn(mb0, xb[d]), __fmul_rn(mb1, xb[d]));
    }
#pragma unroll
    for (int n = 0; n < BG::NC; ++n) pc[n] = make_float2(__fmul_rn(mc0, xc[n]), __fmul_rn(mc1, xc[n]));

#pragma unroll
    for (int da = 0; da < 2; ++da)
#pragma unroll
      for (int db = 0; db < 2; ++db)
#pragma unroll
        for (int n = 0; n < BG::NC; ++n) {
          // ((m_0 x_0 + m_1 x_1) + m_2 x_2) + m_3 in the reference's axis order
          float2 s;
          if (DEPTH == 2) s = __fadd2_rn(__fadd2_rn(pa[da], pb[db]), pc[n]);       // (a, b, c) = (0, 1, 2)
          else if (DEPTH == 1) s = __fadd2_rn(__fadd2_rn(pa[da], pc[n]), pb[db]);  // (a, c, b) = (0, 1, 2)
          else s = __fadd2_rn(__fadd2_rn(pc[n], pa[da]), pb[db]);                   // (c, a, b) = (0, 1, 2)
          const float2 l = __fadd2_rn(__fadd2_rn(s, off), make_float2(-0.25f, -0.25f));
          int r, c;
          float2 t, u;
          bins3(l, r, c, t, u);
          float* z = w0 + (r * WC + c);
          const float val = x[da][db][n];
          const float vt = __fmul_rn(val, t.y), vu = __fmul_rn(val, u.y);  // column weights t1, u1
          if (ATOMIC) {
            atomicAdd(z, vt * t.x);
            atomicAdd(z + WC, vt * u.x);
            atomicAdd(z + 1, vu * t.x);
            atomicAdd(z + WC + 1, vu * u.x);
          } else {
            // lanes of one instruction touch distinct bins (host-checked lattice): plain RMW per tap,
            // __syncwarp orders a tap's stores before the next tap's loads of other lanes
            z[0] = fmaf(vt, t.x, z[0]);
            __syncwarp();
            z[WC] = fmaf(vt, u.x, z[WC]);
            __syncwarp();
            z[1] = fmaf(vu, t.x, z[1]);
            __syncwarp();
            z[WC + 1] = fmaf(vu, u.x, z[WC + 1]);
            __syncwarp();
          }
        }
    __syncwarp();

    // flush the window into view v of the (local) sinogram
    float* y = sino + (size_t)v * p.D0 * (size_t)p.D1;
    if (VEC4) {
      for (int e = lane; e < WR * WC / 4; e += 32) {
        const int wr = e / (WC / 4), g = e - wr * (WC / 4);
        const int row = rb - p.row_off + wr, col = cb + 4 * g;
        const float4 q = reinterpret_cast<const float4*>(win)[e];
        const unsigned any = __float_as_uint(q.x) | __float_as_uint(q.y) | __float_as_uint(q.z) | __float_as_uint(q.w);
        if ((unsigned)row < (unsigned)p.D0 && (unsigned)col < (unsigned)p.D1 && (any << 1) != 0u)
          red_add_v4(y + (size_t)row * p.D1 + col, q.x, q.y, q.z, q.w);
      }
    } else {
      for (int e = lane; e < WR * WC; e += 32) {
        const int wr = e / WC, wc = e - wr * WC;
        const int row = rb - p.row_off + wr, col = cb + wc;
        const float q = win[e];
        if ((unsigned)row < (unsigned)p.D0 && (unsigned)col < (unsigned)p.D1 && q != 0.f)
          atomicAdd(y + (size_t)row * p.D1 + col, q);
      }
    }
    __syncwarp();  // all window reads done before the next view zeroes it
  }
}

}  // namespace xct
