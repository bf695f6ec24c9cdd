// Warp-autonomous "plane" kernels: the fast path of both projectors.
//
// Problem shape shared by the 2D projector and the separable 3D projector:
//   * a stack of NS independent slices (3D: voxel axis 0, 2D: the image batch),
//   * each slice is an NA x NB plane whose points (a, b) project, per view, to ONE detector
//     coordinate u(a, b) that does not depend on the slice, with a 2-bin footprint (c, c+1),
//   * the slice only selects the detector ROW(s) the two bins live in.
// So (c, w0, w1) is computed once per (view, a, b) and reused for every slice a warp owns.
//
// Both directions are register-stationary in the operand that is re-used across views and stream
// the other one through a small warp-private window in shared memory:
//   adjoint : thread owns TA x S voxels (accumulators in registers), loops over ALL views in
//             order; per view the warp stages the S sinogram row segments its tile projects onto
//             (rows pre-combined with the axis-0 weights) and gathers 2 taps per voxel.
//   forward : thread owns TN x GS x S voxels (values in registers), loops over the views of its
//             class; per view it accumulates into a warp-private window of (A, B) bin pairs with
//             plain read-modify-write -- lanes sit GS voxels apart along the view's major axis,
//             so no two lanes of one instruction ever touch the same bin and no shared-memory
//             atomic (a CAS loop for fp32) is needed -- then flushes the window to the sinogram
//             with coalesced RED.ADD.F32, once per (view, tile, slice).
// No block-level barrier is used anywhere: warps never share data, only __syncwarp().
#pragma once
#include "xct_geom.cuh"

namespace xct {

struct PlaneParams {
  const ViewRec* views;  // [V]
  const RowRec* rows;    // 3D: [V][NS]   2D: nullptr
  const int* view_list;  // forward: views of this launch's class; nullptr = 0..n_list-1
  int n_list;            // views to process
  int V;                 // views held by the plan (leading dim of the sinogram)
  int NA, NB, NS;        // plane dims, number of slices
  int D0, D1;            // local detector rows (3D only), bins per row
  int tilesA, tilesB;    // tile grid of the launch
  int views_per_chunk;   // forward: views per blockIdx.y
};

// uniform (warp-wide identical address) loads of the small per-view records
__device__ __forceinline__ ViewRec load_view(const ViewRec* p) {
  const float4* q = reinterpret_cast<const float4*>(p);
  float4 lo = __ldg(q), hi = __ldg(q + 1);
  ViewRec v;
  v.ca = lo.x; v.cb = lo.y; v.off = lo.z; v.width = lo.w;
  v.rwidth = hi.x; v.jump = hi.y; v.fjump = hi.z; v.krow = __float_as_int(hi.w);
  return v;
}
__device__ __forceinline__ RowRec load_row(const RowRec* p) {
  int4 q = __ldg(reinterpret_cast<const int4*>(p));
  RowRec r;
  r.r0 = q.x; r.w0 = __int_as_float(q.y); r.w1 = __int_as_float(q.z); r.pad = 0;
  return r;
}

// First bin of the window a tile projects onto.  u is monotone in a and in b (every rounding
// step is monotone), so the minimum over the tile is attained at the corner picked by the signs
// of the two slopes, evaluated with the same arithmetic as every voxel: the bound is exact.
template <class G>
__device__ __forceinline__ int window_start(const ViewRec& vr, int a_lo, int a_hi, int b_lo, int b_hi) {
  int a = vr.ca >= 0.f ? a_lo : a_hi;
  int b = vr.cb >= 0.f ? b_lo : b_hi;
  return __float2int_rd(G::combine(vr, G::hoistA(vr, a), G::hoistB(vr, b)));
}

// ------------------------------------------------------------------------------------ adjoint
// Tile: TA rows (axis A) x 32 columns (axis B, lane = column) x S slices.  WIN = staged bins.
// ROUTE: the result is ADDED into the row blocks of `route` (the owners' memory, possibly across NVLink)
// instead of being stored to `out`; 2D: one image, row = image axis 0; 3D: row = slice.
template <class G, bool IS3D, int S, int TA, int WIN, int WARPS, bool ROUTE = false>
__global__ void __launch_bounds__(WARPS * 32, ROUTE ? (IS3D ? 2 : 4) : 0)  // routed: the plain kernel's occupancy (0 = unspecified)
plane_adjoint_kernel(PlaneParams p, const float* __restrict__ sino, float* __restrict__ out,
                     const __grid_constant__ OutRoute route) {
  static_assert(WIN % 32 == 0, "window is staged 32 bins at a time");
  constexpr int Q = WIN / 32;
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sgroups = (p.NS + S - 1) / S;
  const long long ntasks = (long long)sgroups * p.tilesA * p.tilesB;
  long long task = (long long)blockIdx.x * WARPS + warp;
  if (task >= ntasks) return;  // warp-uniform; no block barrier below
  const int tb = (int)(task % p.tilesB);
  task /= p.tilesB;
  const int ta = (int)(task % p.tilesA);
  const int sg = (int)(task / p.tilesA);
  const int a0 = ta * TA, b0 = tb * 32, s0 = sg * S;
  const int b = b0 + lane;
  float* zs = smem + warp * (2 * S * WIN);

  float acc[TA][S];
#pragma unroll
  for (int n = 0; n < TA; ++n)
#pragma unroll
    for (int s = 0; s < S; ++s) acc[n][s] = 0.f;

  float pre[S][Q];  // next view's window, in flight while the current view is consumed
  int c0_next = 0;

  auto fetch = [&](int v, int& c0) {
    const ViewRec vr = load_view(p.views + v);
    c0 = window_start<G>(vr, a0, a0 + TA - 1, b0, b0 + 31);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int sl = s0 + s;
      if (IS3D) {
        RowRec rr;
        rr.r0 = 0; rr.w0 = 0.f; rr.w1 = 0.f;
        if (sl < p.NS) rr = load_row(p.rows + (size_t)v * p.NS + sl);
        const float* y0 = sino + ((size_t)v * p.D0 + rr.r0) * (size_t)p.D1;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int col = c0 + lane + 32 * q;
          float val = 0.f;
          if (col >= 0 && col < p.D1) {
            if (rr.w0 != 0.f) val = rr.w0 * __ldg(y0 + col);
            if (rr.w1 != 0.f) val = fmaf(rr.w1, __ldg(y0 + p.D1 + col), val);
          }
          pre[s][q] = val;
        }
      } else {
        const float* y0 = sino + ((size_t)sl * p.V + v) * (size_t)p.D1;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int col = c0 + lane + 32 * q;
          pre[s][q] = (sl < p.NS && col >= 0 && col < p.D1) ? __ldg(y0 + col) : 0.f;
        }
      }
    }
  };

  // small problems: the views are split over blockIdx.y so that the grid fills the SMs; the partial
  // sums of the chunks then meet in the (pre-zeroed) output through RED.ADD
  const int v_begin = blockIdx.y * p.views_per_chunk;
  const int v_end = min(p.n_list, v_begin + p.views_per_chunk);
  if (v_begin < v_end) fetch(v_begin, c0_next);
  for (int v = v_begin; v < v_end; ++v) {
    float* zb = zs + ((v - v_begin) & 1) * (S * WIN);
    const int c0 = c0_next;
    if constexpr (S == 4) {
      // four slices: the window is [bin][slice], one float4 per bin, so that a tap of all four slices is ONE LDS.128
      // (a quarter-warp's 8 lanes sit on <= 8 consecutive bins: conflict-free) instead of four LDS.32
#pragma unroll
      for (int q = 0; q < Q; ++q)
        reinterpret_cast<float4*>(zb)[lane + 32 * q] = make_float4(pre[0][q], pre[1][q], pre[2][q], pre[3][q]);
    } else {
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int q = 0; q < Q; ++q) zb[s * WIN + lane + 32 * q] = pre[s][q];
    }
    __syncwarp();
    if (v + 1 < v_end) fetch(v + 1, c0_next);

    const ViewRec vr = load_view(p.views + v);
    const float hB = G::hoistB(vr, b);
    if constexpr (S == 1 && !IS3D && TA % 2 == 0) {
      // single 2D image: nothing amortises the coordinate arithmetic over slices, so two ROWS share
      // every packed instruction (coordinates, weights and the two taps: ~13 instead of 20 issue slots
      // per update; the kernel is issue-bound at 88 %, profiles/README.md)
      const float2 hB2 = make_float2(hB, hB);
      const float xa0 = G::coordA(a0);
#pragma unroll
      for (int n = 0; n < TA; n += 2) {
        const float2 u = __fadd2_rn(G::hoistA2(vr, make_float2(xa0 + (float)n, xa0 + (float)(n + 1))), hB2);
        int ca_, cb_;
        float2 w0, w1;
        G::bins2(vr, u, ca_, cb_, w0, w1);
        const float* za = zb + min((unsigned)(ca_ - c0), (unsigned)(WIN - 2));  // in range for validated plans
        const float* zc = zb + min((unsigned)(cb_ - c0), (unsigned)(WIN - 2));
        const float2 lo = make_float2(za[0], zc[0]), hi = make_float2(za[1], zc[1]);
        const float2 r = __ffma2_rn(hi, w1, __ffma2_rn(lo, w0, make_float2(acc[n][0], acc[n + 1][0])));
        acc[n][0] = r.x;
        acc[n + 1][0] = r.y;
      }
    } else {
#pragma unroll
      for (int n = 0; n < TA; ++n) {
        const float u = G::combine(vr, G::hoistA(vr, a0 + n), hB);
        int c;
        float w0, w1;
        G::bins(vr, u, c, w0, w1);
        int t = c - c0;
        t = min(max(t, 0), WIN - 2);  // never taken for validated plans; keeps smem access in range
        if constexpr (S == 4) {
          const float4 lo = reinterpret_cast<const float4*>(zb)[t], hi = reinterpret_cast<const float4*>(zb)[t + 1];
          acc[n][0] = fmaf(hi.x, w1, fmaf(lo.x, w0, acc[n][0]));
          acc[n][1] = fmaf(hi.y, w1, fmaf(lo.y, w0, acc[n][1]));
          acc[n][2] = fmaf(hi.z, w1, fmaf(lo.z, w0, acc[n][2]));
          acc[n][3] = fmaf(hi.w, w1, fmaf(lo.w, w0, acc[n][3]));
        } else {
          const float* z = zb + t;
#pragma unroll
          for (int s = 0; s < S; ++s)
            acc[n][s] = fmaf(z[s * WIN + 1], w1, fmaf(z[s * WIN], w0, acc[n][s]));
        }
      }
    }
  }

  if constexpr (ROUTE) {
    // Routed epilogue.  It is kept SMALL on purpose: an owner search inlined into every one of the TA (x S)
    // unrolled writes made this kernel 9 % slower than the plain one although the epilogue runs once per warp
    // (its ~100 KB of straight-line code went through the instruction cache of every SM 16 000 times and
    // evicted the view loop of the other warps).  So: one owner lookup per column (2D) or per slice (3D),
    // then plain pointer steps; a tile that straddles two row blocks takes the per-row lookup (rare).
    if (b >= p.NB) return;
    if constexpr (!IS3D) {
      RouteCursor cur;
      cur.seek(route, a0, b);
      if (a0 + TA <= cur.next) {  // the whole tile belongs to one row block (warp-uniform)
        float* q = cur.q;
        const long long step = route.inner;
        if (route.store) {
#pragma unroll
          for (int n = 0; n < TA; ++n)
            if (a0 + n < p.NA) q[n * step] = acc[n][0];
        } else {
#pragma unroll
          for (int n = 0; n < TA; ++n)
            if (a0 + n < p.NA) atomicAdd_system(q + n * step, acc[n][0]);
        }
      } else {
#pragma unroll
        for (int n = 0; n < TA; ++n)
          if (a0 + n < p.NA) route_add(route, a0 + n, b, acc[n][0]);
      }
    } else {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (s0 + s >= p.NS) break;
        RouteCursor cur;
        cur.seek(route, s0 + s, (long long)a0 * p.NB + b);  // the slice picks the owner
        float* q = cur.q;
        if (route.store) {
#pragma unroll
          for (int n = 0; n < TA; ++n)
            if (a0 + n < p.NA) q[(size_t)n * p.NB] = acc[n][s];
        } else {
#pragma unroll
          for (int n = 0; n < TA; ++n)
            if (a0 + n < p.NA) atomicAdd_system(q + (size_t)n * p.NB, acc[n][s]);
        }
      }
    }
    return;
  }
  if (b < p.NB) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (s0 + s >= p.NS) break;
#pragma unroll
      for (int n = 0; n < TA; ++n) {
        if (a0 + n < p.NA) {
          float* o = out + ((size_t)(s0 + s) * p.NA + (a0 + n)) * (size_t)p.NB + b;
          if (gridDim.y > 1) atomicAdd(o, acc[n][s]);
          else *o = acc[n][s];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------ forward
// Window slot of bin t.  Lanes of one RMW instruction sit >= 1.41 bins apart (GS voxels along the
// major axis), so a half-warp's 16 bins span up to 32 bins and a linear layout wraps the 16 float2
// bank slots twice (2-way conflicts, half of all wavefronts in the first ncu capture).  Splitting
// the window by bin parity halves the span inside each parity class; the odd half is offset by
// 8 slots (mod 16) so that the two classes of one half-warp rarely meet.
template <int WIN>
struct FwdSlots {
  static constexpr int HALF = WIN / 2;
  static constexpr int ODD_BASE = ((HALF + 15) / 16) * 16 + 8;
  static constexpr int SIZE = ODD_BASE + HALF;  // float2 slots per slice
  static __device__ __forceinline__ int of(int t) { return (t >> 1) + ((t & 1) ? ODD_BASE : 0); }
};

// Tile: 32*GS points along the major axis (lane l owns points GS*l .. GS*l+GS-1) x TN points along
// the minor axis x S slices.  MAJOR_B: the major axis is plane axis B (the contiguous one).
template <class G, bool IS3D, int S, int TN, int GS, int WIN, bool MAJOR_B, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
plane_forward_kernel(PlaneParams p, const float* __restrict__ in, float* __restrict__ sino) {
  static_assert(WIN % 32 == 0, "window is flushed 32 bins at a time");
  constexpr int Q = WIN / 32;
  constexpr int TM = 32 * GS;
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sgroups = (p.NS + S - 1) / S;
  const long long ntasks = (long long)sgroups * p.tilesA * p.tilesB;
  long long task = (long long)blockIdx.x * WARPS + warp;
  if (task >= ntasks) return;
  const int tb = (int)(task % p.tilesB);
  task /= p.tilesB;
  const int ta = (int)(task % p.tilesA);
  const int sg = (int)(task / p.tilesA);
  const int a0 = ta * (MAJOR_B ? TN : TM), b0 = tb * (MAJOR_B ? TM : TN), s0 = sg * S;
  using Slots = FwdSlots<WIN>;
  constexpr int WS = Slots::SIZE;
  float2* acc = reinterpret_cast<float2*>(smem) + warp * (S * WS);

  // this thread's voxels, kept for every view of the launch
  float x[TN][GS][S];
#pragma unroll
  for (int n = 0; n < TN; ++n)
#pragma unroll
    for (int d = 0; d < GS; ++d) {
      const int a = MAJOR_B ? a0 + n : a0 + GS * lane + d;
      const int b = MAJOR_B ? b0 + GS * lane + d : b0 + n;
      const bool ok = a < p.NA && b < p.NB;
#pragma unroll
      for (int s = 0; s < S; ++s)
        x[n][d][s] = (ok && s0 + s < p.NS)
                         ? __ldg(in + ((size_t)(s0 + s) * p.NA + a) * (size_t)p.NB + b)
                         : 0.f;
    }

  const int v_begin = blockIdx.y * p.views_per_chunk;
  const int v_end = min(p.n_list, v_begin + p.views_per_chunk);
  for (int vi = v_begin; vi < v_end; ++vi) {
    const int v = p.view_list ? __ldg(p.view_list + vi) : vi;
    const ViewRec vr = load_view(p.views + v);
    const int c0 = window_start<G>(vr, a0, a0 + (MAJOR_B ? TN : TM) - 1, b0, b0 + (MAJOR_B ? TM : TN) - 1);

#pragma unroll
    for (int s = 0; s < S; ++s)
      for (int i = lane; i < WS; i += 32) acc[s * WS + i] = make_float2(0.f, 0.f);
    __syncwarp();

    float hMaj[GS];
#pragma unroll
    for (int d = 0; d < GS; ++d)
      hMaj[d] = MAJOR_B ? G::hoistB(vr, b0 + GS * lane + d) : G::hoistA(vr, a0 + GS * lane + d);

#pragma unroll
    for (int n = 0; n < TN; ++n) {
      const float hMin = MAJOR_B ? G::hoistA(vr, a0 + n) : G::hoistB(vr, b0 + n);
#pragma unroll
      for (int d = 0; d < GS; ++d) {
        const float u = MAJOR_B ? G::combine(vr, hMin, hMaj[d]) : G::combine(vr, hMaj[d], hMin);
        int c;
        float w0, w1;
        G::bins(vr, u, c, w0, w1);
        int t = c - c0;
        t = min(max(t, 0), WIN - 2);
        float2* pa = acc + Slots::of(t);
        // lanes are GS voxels apart along the major axis => their bins differ by >= 1:
        // one lane per address, plain RMW is race free within this instruction
#pragma unroll
        for (int s = 0; s < S; ++s) {
          float2 ab = pa[s * WS];
          ab.x = fmaf(x[n][d][s], w0, ab.x);
          ab.y = fmaf(x[n][d][s], w1, ab.y);
          pa[s * WS] = ab;
        }
        __syncwarp();  // order this step's stores before the next step's loads of other lanes
      }
    }

    // flush: bin (c0 + t) = A[t] + B[t-1]; rows / weights from the slice's row record
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int sl = s0 + s;
      if (sl >= p.NS) break;
      RowRec rr;
      float* y0;
      if (IS3D) {
        rr = load_row(p.rows + (size_t)v * p.NS + sl);
        y0 = sino + ((size_t)v * p.D0 + rr.r0) * (size_t)p.D1;
      } else {
        rr.r0 = 0; rr.w0 = 1.f; rr.w1 = 0.f;
        y0 = sino + ((size_t)sl * p.V + v) * (size_t)p.D1;
      }
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int t = lane + 32 * q;
        float val = acc[s * WS + Slots::of(t)].x;
        if (t > 0) val += acc[s * WS + Slots::of(t - 1)].y;
        const int col = c0 + t;
        if (val != 0.f && col >= 0 && col < p.D1) {
          if (rr.w0 != 0.f) atomicAdd(y0 + col, rr.w0 * val);
          if (IS3D && rr.w1 != 0.f) atomicAdd(y0 + p.D1 + col, rr.w1 * val);
        }
      }
    }
    __syncwarp();  // all window reads done before the next view zeroes it
  }
}

}  // namespace xct
