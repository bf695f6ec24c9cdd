// Second-generation plane kernels ("walk" kernels): same problem decomposition as xct_plane.cuh
// (a stack of independent slices whose in-plane points (a, b) project to ONE detector coordinate
// u(a, b) per view with a 2-bin footprint), but each thread now WALKS along plane axis A and keeps
// the two bins under its current position in registers:
//
//   adjoint : lane = column b, thread owns TA rows x S slices (register accumulators), loops over
//             all views in order.  The S sinogram row segments a tile projects onto are copied
//             global -> shared with 16-byte cp.async (LDGSTS) into a STAGES-deep ring, no
//             registers and no branches in the fetch.  Walking down the rows the bin index moves
//             by at most one per step (|ca| <= 1), so the pair (z[c], z[c+1]) is carried in
//             registers and only ONE new tap is read from shared memory per voxel-view update
//             (xct_plane.cuh reads two): the kernel moves from LSU-bound to issue-bound.
//
// Race freedom / exactness: identical coordinate arithmetic to xct_geom.cuh (the bin index is the
// oracle's bit for bit); accumulation order per voxel is views ascending, tap c then tap c+1, as in
// the reference's lax.scan (_xray3d.py:189,200-203).  A bin jump other than the expected +-1
// (possible only when |ca| > 1) takes a cold path that reloads both taps, so the kernel is correct
// for any geometry inside the window envelope checked at plan creation.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is resolved at run time, see xct_api.cu)

#include <type_traits>

#include "xct_plane.cuh"

namespace xct {

struct Walk2Params {
  PlaneParams p;
  const long long* rowoff;  // [V][row_stride] element offset of the sinogram row slice s reads in view v, -1 = none
  float out_scale;          // 3D: 2.0 (= 4 * 0.5, the axis-0 weight of a full row); 2D: 1
  // slice sub-range launches (host pipeline, xct_api.cu): the kernel sees p.NS slices starting at
  // slice s_base of the plan; `in` / `out` volume pointers are already offset by the caller, the
  // per-(view, slice) tables are indexed with the plan's stride.
  int row_stride;           // slices per view in rowoff / p.rows (the plan's n0)
  int s_base;               // first slice of this launch
};

__device__ __forceinline__ void cp_async16_zfill(float* smem_dst, const float* gmem_src, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers for the walk adjoint's sinogram window ----
// (barriers and destinations are passed as 32-bit shared-window addresses, computed once per warp)
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// one elected lane of a converged warp (elect.sync: lets ptxas issue the uniform-datapath TMA
// instruction directly instead of looping over the lanes that might hold different operands)
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// box (x: WIN bins, y: S rows, z: 1 view) of the (V, D0, D1) sinogram -> shared; out-of-bounds
// elements (negative bins, bins >= D1, rows outside [0, D0)) arrive as zeros
__device__ __forceinline__ void tma_load_3d(unsigned smem_dst, const CUtensorMap* tmap, int x, int y, int z, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_dst),
      "l"(tmap), "r"(bar), "r"(x), "r"(y), "r"(z)
      : "memory");
}

// Window start rounded down to a multiple of 4 bins so that every 16-byte chunk of a staged row is
// either entirely inside [0, D1) or entirely outside (D1 % 4 == 0 is required by the launcher).
template <class G>
__device__ __forceinline__ int window_start4(const ViewRec& vr, int a_lo, int a_hi, int b_lo, int b_hi) {
  return window_start<G>(vr, a_lo, a_hi, b_lo, b_hi) & ~3;
}

// Tile: TA rows (axis A) x 32 columns (axis B, lane = column) x S slices.
// smem per warp: STAGES * S * WIN floats.
// TMA: detector rows of consecutive slices are consecutive (row = slice + ViewRec::krow for every
// view), so the S x WIN window of a view is ONE box of the (V, D0, D1) sinogram: lane 0 issues a
// single cp.async.bulk.tensor per view that completes on a per-stage mbarrier (hardware zero fill
// outside the detector) instead of every lane issuing 16-byte cp.async with its own row lookups.
// ROUTE: every result slice goes to the memory of the row block that owns it (view-block sharding fused with
// its exchange, xct_adjoint_scatter): one owner lookup per slice, then TA stores (or system-scope RED.ADD)
// at a fixed stride, as in plane_adjoint_kernel's routed epilogue.  The view loop is the plain kernel's.
template <class G, bool IS3D, int S, int TA, int WIN, int STAGES, int WARPS, bool TMA, bool ROUTE = false>
__global__ void __launch_bounds__(WARPS * 32, 2)  // two CTAs (16 warps) per SM: at most 128 registers
walk_adjoint_kernel(Walk2Params wp, const float* __restrict__ sino, float* __restrict__ out,
                    const __grid_constant__ CUtensorMap tmap, const __grid_constant__ OutRoute route) {
  static_assert(WIN % 4 == 0, "rows are staged in 16-byte chunks");
  constexpr int CPR = WIN / 4;                    // 16-byte chunks per staged row
  constexpr int CHUNKS = S * CPR;                 // chunks per view
  constexpr int CPL = (CHUNKS + 31) / 32;         // chunks per lane
  const PlaneParams& p = wp.p;
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
  float* ring = smem + (size_t)warp * (STAGES * S * WIN);
  // TMA: one mbarrier per (warp, stage), behind the rings
  const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring);
  const unsigned bars_sa =
      (unsigned)__cvta_generic_to_shared(smem + (size_t)WARPS * (STAGES * S * WIN)) + warp * STAGES * 8u;
  if (TMA) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < STAGES; ++i) mbar_init(bars_sa + 8u * i, 1);
      mbar_fence_init();
    }
    __syncwarp();
  }

  // accumulators as slice PAIRS: the two taps of two slices are one FFMA2 each (fma.rn.f32x2,
  // sm_100: two IEEE fp32 fmas per issue slot; per element identical to fmaf)
  static_assert(S % 2 == 0, "slices are processed in pairs");
  constexpr int H = S / 2;
  float2 acc[TA][H];
#pragma unroll
  for (int n = 0; n < TA; ++n)
#pragma unroll
    for (int h = 0; h < H; ++h) acc[n][h] = make_float2(0.f, 0.f);

  // ---- producer: one view's S row segments, global -> shared, asynchronous.
  // Chunk -> (slice, 4-bin group) assignment is fixed per lane; only the row offset and the
  // window start change from view to view.
  int ck_soff[CPL];   // smem float offset of the chunk inside a stage
  int ck_col[CPL];    // 4 * k
  int ck_sl[CPL];     // slice, clamped into [0, NS)
  bool ck_live[CPL];  // slice exists and the chunk id is in range
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    const int id = lane + 32 * q;
    const int s = id / CPR, k = id % CPR;
    ck_soff[q] = s * WIN + 4 * k;
    ck_col[q] = 4 * k;
    ck_live[q] = id < CHUNKS && s0 + s < p.NS;
    ck_sl[q] = min(s0 + s, p.NS - 1);
  }
  auto fetch = [&](int v, int c0) {
    float* zb = ring + (v % STAGES) * (S * WIN);
    const long long* ro = wp.rowoff + (size_t)v * wp.row_stride + wp.s_base;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const long long off = __ldg(ro + ck_sl[q]);  // element offset of the row, -1 = no row
      const int col = c0 + ck_col[q];
      const bool ok = ck_live[q] && off >= 0 && (unsigned)col < (unsigned)p.D1;
      const float* src = sino + (ok ? off + col : 0);
      if (CHUNKS % 32 == 0 || lane + 32 * q < CHUNKS) cp_async16_zfill(zb + ck_soff[q], src, ok);
    }
  };
  auto view_c0 = [&](int v) {
    const ViewRec vr = load_view(p.views + v);
    return window_start4<G>(vr, a0, a0 + TA - 1, b0, b0 + 31);
  };
  // TMA producer: the whole S x WIN window of view v in one bulk tensor copy (lane 0)
  const int row_base = wp.s_base + s0;
  auto fetch_tma = [&](int v, int stage) {
    const ViewRec vr = load_view(p.views + v);
    const int c0 = window_start4<G>(vr, a0, a0 + TA - 1, b0, b0 + 31);
    if (elect_one()) {
      const unsigned bar = bars_sa + 8u * stage;
      mbar_expect_tx(bar, S * WIN * (unsigned)sizeof(float));
      tma_load_3d(ring_sa + stage * (S * WIN * (unsigned)sizeof(float)), &tmap, c0, row_base + vr.krow, v, bar);
    }
    return c0;
  };

  // ---- consumer: walk down the TA rows of this thread's column, carrying (z[c], z[c+1]).
  // COLD (per view, warp-uniform: ViewRec::jump set on the host when |ca| is so close to 1 that
  // rounding can move the bin by two in one row step) reloads the other tap after such a jump.
  const float xa0 = G::coordA(a0);
  auto walk = [&](auto up_c, auto cold_c, const float* zb, const ViewRec& vr, int c0, float hB) {
    constexpr bool UP = decltype(up_c)::value;
    constexpr bool COLD = decltype(cold_c)::value;
    float2 lo[H], hi[H];
    int tp = 0;
    static_assert(TA % 2 == 0, "rows are evaluated in pairs");
    int cc[2];
    float2 ww0, ww1;
#pragma unroll
    for (int n = 0; n < TA; ++n) {
      if ((n & 1) == 0) {  // coordinates and weights of rows n, n + 1 in packed fp32
        const float2 hA = make_float2(G::hoistA_x(vr, xa0 + (float)n), G::hoistA_x(vr, xa0 + (float)(n + 1)));
        G::bins2(vr, G::combine2(vr, hA, make_float2(hB, hB)), cc[0], cc[1], ww0, ww1);
      }
      const int c = cc[n & 1];
      const float w0 = (n & 1) ? ww0.y : ww0.x, w1 = (n & 1) ? ww1.y : ww1.x;
      const int t = (int)min((unsigned)(c - c0), (unsigned)(WIN - 2));
      const float* z = zb + t;
      if (n == 0) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          lo[h] = make_float2(z[(2 * h) * WIN], z[(2 * h + 1) * WIN]);
          hi[h] = make_float2(z[(2 * h) * WIN + 1], z[(2 * h + 1) * WIN + 1]);
        }
      } else {
        const bool moved = t != tp;
        if (UP) {
#pragma unroll
          for (int h = 0; h < H; ++h) {
            lo[h].x = moved ? hi[h].x : lo[h].x;
            lo[h].y = moved ? hi[h].y : lo[h].y;
            hi[h] = make_float2(z[(2 * h) * WIN + 1], z[(2 * h + 1) * WIN + 1]);
          }
          if (COLD && moved && t != tp + 1) {
#pragma unroll
            for (int h = 0; h < H; ++h) lo[h] = make_float2(z[(2 * h) * WIN], z[(2 * h + 1) * WIN]);
          }
        } else {
#pragma unroll
          for (int h = 0; h < H; ++h) {
            hi[h].x = moved ? lo[h].x : hi[h].x;
            hi[h].y = moved ? lo[h].y : hi[h].y;
            lo[h] = make_float2(z[(2 * h) * WIN], z[(2 * h + 1) * WIN]);
          }
          if (COLD && moved && t != tp - 1) {
#pragma unroll
            for (int h = 0; h < H; ++h) hi[h] = make_float2(z[(2 * h) * WIN + 1], z[(2 * h + 1) * WIN + 1]);
          }
        }
      }
      tp = t;
      const float2 w0p = make_float2(w0, w0), w1p = make_float2(w1, w1);
#pragma unroll
      for (int h = 0; h < H; ++h) acc[n][h] = __ffma2_rn(hi[h], w1p, __ffma2_rn(lo[h], w0p, acc[n][h]));
    }
  };

  // c0 of the views in flight (register ring, rotated once per view)
  int c0q[STAGES];
#pragma unroll
  for (int i = 0; i < STAGES; ++i) c0q[i] = 0;
#pragma unroll
  for (int i = 0; i < STAGES - 1; ++i) {
    if (i < p.n_list) {
      if (TMA) {
        c0q[i] = fetch_tma(i, i);
      } else {
        c0q[i] = view_c0(i);
        fetch(i, c0q[i]);
      }
    }
    if (!TMA) cp_async_commit();
  }

  int st_c = 0, st_p = STAGES - 1;  // stage consumed / produced this iteration (v % STAGES, (v + STAGES - 1) % STAGES)
  unsigned par = 0;                 // mbarrier phase parity of the consumed stage ((v / STAGES) & 1)
  for (int v = 0; v < p.n_list; ++v) {
    // all lanes have finished reading the stage that fetch(v + STAGES - 1) overwrites
    __syncwarp();
    if (v + STAGES - 1 < p.n_list) {
      if (TMA) {
        c0q[STAGES - 1] = fetch_tma(v + STAGES - 1, st_p);
      } else {
        c0q[STAGES - 1] = view_c0(v + STAGES - 1);
        fetch(v + STAGES - 1, c0q[STAGES - 1]);
      }
    }
    if (TMA) {
      mbar_wait(bars_sa + 8u * st_c, par);  // stage v has landed (all lanes observe it)
    } else {
      cp_async_commit();
      cp_async_wait<STAGES - 1>();
      __syncwarp();  // every lane's copies of stage v are complete and visible
    }

    const float* zb = ring + st_c * (S * WIN);
    st_p = st_c;
    if (++st_c == STAGES) { st_c = 0; par ^= 1u; }
    const ViewRec vr = load_view(p.views + v);
    const int c0 = c0q[0];
#pragma unroll
    for (int i = 0; i + 1 < STAGES; ++i) c0q[i] = c0q[i + 1];
    const float hB = G::hoistB(vr, b);
    if (vr.jump != 0.f) {  // rare (|ca| within rounding distance of 1)
      if (vr.ca >= 0.f) walk(std::true_type{}, std::true_type{}, zb, vr, c0, hB);
      else walk(std::false_type{}, std::true_type{}, zb, vr, c0, hB);
    } else if (vr.ca >= 0.f) {
      walk(std::true_type{}, std::false_type{}, zb, vr, c0, hB);  // bins non-decreasing down the rows
    } else {
      walk(std::false_type{}, std::false_type{}, zb, vr, c0, hB);
    }
  }
  if (!TMA) cp_async_wait<0>();

  if constexpr (ROUTE) {
    if (b >= p.NB) return;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (s0 + s >= p.NS) break;
      RouteCursor cur;
      cur.seek(route, wp.s_base + s0 + s, (long long)a0 * p.NB + b);  // the slice picks the owner
      float* q = cur.q;
      if (route.store) {
#pragma unroll
        for (int n = 0; n < TA; ++n)
          if (a0 + n < p.NA) q[(size_t)n * p.NB] = ((s & 1) ? acc[n][s / 2].y : acc[n][s / 2].x) * wp.out_scale;
      } else {
#pragma unroll
        for (int n = 0; n < TA; ++n)
          if (a0 + n < p.NA) atomicAdd_system(q + (size_t)n * p.NB, ((s & 1) ? acc[n][s / 2].y : acc[n][s / 2].x) * wp.out_scale);
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
        if (a0 + n < p.NA)
          out[((size_t)(s0 + s) * p.NA + (a0 + n)) * (size_t)p.NB + b] =
              ((s & 1) ? acc[n][s / 2].y : acc[n][s / 2].x) * wp.out_scale;
      }
    }
  }
}

// ---------------------------------------------------------- adjoint, slice-interleaved window
// walk_adjoint_kernel reads one scalar tap per (row step, slice): with S slices per thread a row step is S LDS.32.
// Here the sinogram is first re-laid out by sino_interleave4_kernel as (V, G, D1, 4): the four detector rows that
// four consecutive slices project onto become the four components of one float4 per bin.  A view's window is then
// ONE TMA box of (4 WIN floats, S / 4 groups, 1 view) and a row step reads the new tap of FOUR slices with one
// LDS.128 (a quarter-warp's 8 lanes sit on <= 8 consecutive bins = 128 contiguous bytes: conflict-free), i.e.
// S / 4 shared loads per row step instead of S -- same bytes, a quarter of the LSU issue slots.  The window needs
// no 4-bin alignment any more (bin c sits at byte 16 c of a row), so it shrinks to TA + 31 + 2 bins.  Everything
// else -- the carried (z[c], z[c+1]) pair, the packed row-pair coordinates, the per-view jump variant, the order
// of accumulation (views ascending, tap c then c + 1) -- is walk_adjoint_kernel's: results are bit-identical.
//
// out[v][g][c][j] = sino[v][4 g + j + krow(v)][c]  (0 outside the detector; g = slice group of the PLAN)
__global__ void __launch_bounds__(256)
sino_interleave4_kernel(const ViewRec* __restrict__ views, const float* __restrict__ sino, float* __restrict__ out, int D0,
                        int D1, int g_begin, int g_count, int g_total) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = g_begin + blockIdx.y, v = blockIdx.z;
  if (c >= D1) return;
  const int krow = load_view(views + v).krow;
  const float* y = sino + (size_t)v * D0 * (size_t)D1 + c;
  float q[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = 4 * g + j + krow;
    q[j] = (unsigned)r < (unsigned)D0 ? __ldg(y + (size_t)r * D1) : 0.f;
  }
  reinterpret_cast<float4*>(out)[((size_t)v * g_total + g) * (size_t)D1 + c] = make_float4(q[0], q[1], q[2], q[3]);
}

template <class G, int S, int TA, int WIN, int STAGES, int WARPS, bool ROUTE = false>
__global__ void __launch_bounds__(WARPS * 32, 2)  // two CTAs (16 warps) per SM: at most 128 registers
walk_adjoint_vec_kernel(Walk2Params wp, float* __restrict__ out, const __grid_constant__ CUtensorMap tmap,
                        const __grid_constant__ OutRoute route) {
  static_assert(S % 4 == 0 && TA % 2 == 0 && 4 * WIN <= 256, "float4 slice groups, row pairs, TMA box <= 256 elements");
  constexpr int NG = S / 4, H = S / 2;
  constexpr int STAGE_VEC = NG * WIN;  // float4 slots per stage: [NG][WIN]
  const PlaneParams& p = wp.p;
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
  const float4* ring = reinterpret_cast<const float4*>(smem) + (size_t)warp * (STAGES * STAGE_VEC);
  const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring);
  const unsigned bars_sa =
      (unsigned)__cvta_generic_to_shared(reinterpret_cast<const float4*>(smem) + (size_t)WARPS * (STAGES * STAGE_VEC)) +
      warp * STAGES * 8u;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) mbar_init(bars_sa + 8u * i, 1);
    mbar_fence_init();
  }
  __syncwarp();

  float2 acc[TA][H];
#pragma unroll
  for (int n = 0; n < TA; ++n)
#pragma unroll
    for (int h = 0; h < H; ++h) acc[n][h] = make_float2(0.f, 0.f);

  // producer: the whole [NG][WIN] x 4 window of view v in one bulk tensor copy (one elected lane)
  const int g_base = (wp.s_base + s0) >> 2;  // slice group of the plan (the launcher guarantees s_base % 4 == 0)
  auto fetch_tma = [&](int v, int stage) {
    const ViewRec vr = load_view(p.views + v);
    const int c0 = window_start<G>(vr, a0, a0 + TA - 1, b0, b0 + 31);
    if (elect_one()) {
      const unsigned bar = bars_sa + 8u * stage;
      mbar_expect_tx(bar, STAGE_VEC * (unsigned)sizeof(float4));
      tma_load_3d(ring_sa + stage * (STAGE_VEC * (unsigned)sizeof(float4)), &tmap, 4 * c0, g_base, v, bar);
    }
    return c0;
  };

  const float xa0 = G::coordA(a0);
  auto walk = [&](auto up_c, auto cold_c, const float4* zb, const ViewRec& vr, int c0, float hB) {
    constexpr bool UP = decltype(up_c)::value;
    constexpr bool COLD = decltype(cold_c)::value;
    float4 lo[NG], hi[NG];
    int tp = 0;
    int cc[2];
    float2 ww0, ww1;
#pragma unroll
    for (int n = 0; n < TA; ++n) {
      if ((n & 1) == 0) {  // coordinates and weights of rows n, n + 1 in packed fp32
        const float2 hA = make_float2(G::hoistA_x(vr, xa0 + (float)n), G::hoistA_x(vr, xa0 + (float)(n + 1)));
        G::bins2(vr, G::combine2(vr, hA, make_float2(hB, hB)), cc[0], cc[1], ww0, ww1);
      }
      const int c = cc[n & 1];
      const float w0 = (n & 1) ? ww0.y : ww0.x, w1 = (n & 1) ? ww1.y : ww1.x;
      const int t = (int)min((unsigned)(c - c0), (unsigned)(WIN - 2));
      const float4* z = zb + t;
      if (n == 0) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          lo[g] = z[g * WIN];
          hi[g] = z[g * WIN + 1];
        }
      } else {
        const bool moved = t != tp;
        if (UP) {
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            lo[g].x = moved ? hi[g].x : lo[g].x;
            lo[g].y = moved ? hi[g].y : lo[g].y;
            lo[g].z = moved ? hi[g].z : lo[g].z;
            lo[g].w = moved ? hi[g].w : lo[g].w;
            hi[g] = z[g * WIN + 1];
          }
          if (COLD && moved && t != tp + 1) {
#pragma unroll
            for (int g = 0; g < NG; ++g) lo[g] = z[g * WIN];
          }
        } else {
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            hi[g].x = moved ? lo[g].x : hi[g].x;
            hi[g].y = moved ? lo[g].y : hi[g].y;
            hi[g].z = moved ? lo[g].z : hi[g].z;
            hi[g].w = moved ? lo[g].w : hi[g].w;
            lo[g] = z[g * WIN];
          }
          if (COLD && moved && t != tp - 1) {
#pragma unroll
            for (int g = 0; g < NG; ++g) hi[g] = z[g * WIN + 1];
          }
        }
      }
      tp = t;
      const float2 w0p = make_float2(w0, w0), w1p = make_float2(w1, w1);
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        acc[n][2 * g] = __ffma2_rn(make_float2(hi[g].x, hi[g].y), w1p, __ffma2_rn(make_float2(lo[g].x, lo[g].y), w0p, acc[n][2 * g]));
        acc[n][2 * g + 1] =
            __ffma2_rn(make_float2(hi[g].z, hi[g].w), w1p, __ffma2_rn(make_float2(lo[g].z, lo[g].w), w0p, acc[n][2 * g + 1]));
      }
    }
  };

  int c0q[STAGES];
#pragma unroll
  for (int i = 0; i < STAGES; ++i) c0q[i] = 0;
#pragma unroll
  for (int i = 0; i < STAGES - 1; ++i)
    if (i < p.n_list) c0q[i] = fetch_tma(i, i);

  int st_c = 0, st_p = STAGES - 1;
  unsigned par = 0;
  for (int v = 0; v < p.n_list; ++v) {
    __syncwarp();  // all lanes have finished reading the stage that is refilled now
    if (v + STAGES - 1 < p.n_list) c0q[STAGES - 1] = fetch_tma(v + STAGES - 1, st_p);
    mbar_wait(bars_sa + 8u * st_c, par);
    const float4* zb = ring + st_c * STAGE_VEC;
    st_p = st_c;
    if (++st_c == STAGES) { st_c = 0; par ^= 1u; }
    const ViewRec vr = load_view(p.views + v);
    const int c0 = c0q[0];
#pragma unroll
    for (int i = 0; i + 1 < STAGES; ++i) c0q[i] = c0q[i + 1];
    const float hB = G::hoistB(vr, b);
    if (vr.jump != 0.f) {  // rare (|ca| within rounding distance of 1)
      if (vr.ca >= 0.f) walk(std::true_type{}, std::true_type{}, zb, vr, c0, hB);
      else walk(std::false_type{}, std::true_type{}, zb, vr, c0, hB);
    } else if (vr.ca >= 0.f) {
      walk(std::true_type{}, std::false_type{}, zb, vr, c0, hB);
    } else {
      walk(std::false_type{}, std::false_type{}, zb, vr, c0, hB);
    }
  }

  if (b >= p.NB) return;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    if (s0 + s >= p.NS) break;
    if constexpr (ROUTE) {
      RouteCursor cur;
      cur.seek(route, wp.s_base + s0 + s, (long long)a0 * p.NB + b);  // the slice picks the owner
      float* q = cur.q;
#pragma unroll
      for (int n = 0; n < TA; ++n) {
        if (a0 + n >= p.NA) break;
        const float val = ((s & 1) ? acc[n][s / 2].y : acc[n][s / 2].x) * wp.out_scale;
        if (route.store) q[(size_t)n * p.NB] = val;
        else atomicAdd_system(q + (size_t)n * p.NB, val);
      }
    } else {
#pragma unroll
      for (int n = 0; n < TA; ++n) {
        if (a0 + n < p.NA)
          out[((size_t)(s0 + s) * p.NA + (a0 + n)) * (size_t)p.NB + b] = ((s & 1) ? acc[n][s / 2].y : acc[n][s / 2].x) * wp.out_scale;
      }
    }
  }
}

// ------------------------------------------------------------------------------------ forward
// Tile: 32*GS points along the view's MAJOR axis (lane l owns points GS*l .. GS*l+GS-1) x TN points
// along the MINOR axis x S slices, voxels register-stationary for every view of the launch (as in
// plane_forward_kernel).  Each lane WALKS its GS columns along the minor axis, where the bin index
// moves by at most one per step (|c_minor| <= 0.71 |c_major|...1), and carries the partial sums of
// the two bins under its current position in registers: shared memory sees one read-modify-write
// per bin CHANGE instead of one per voxel, which removes most of the LSU traffic (and the bank
// conflicts) that bound plane_forward_kernel.  Lanes sit GS voxels apart along the major axis, so
// the bins two lanes flush in one instruction differ by >= 1: plain RMW, no shared atomics.
// MINOR_UP: bins are non-decreasing along the walk (the launch holds views of one sign class).
// Window layout: win[t][S] (slice fastest), flushed to the sinogram with RED once per view.
template <int S>
struct SliceVec;
template <>
struct SliceVec<2> { using type = float2; };
template <>
struct SliceVec<4> { using type = float4; };

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// COLD : keep the path for bin jumps larger than one (needed only when |c_minor| can reach 1).
// UNIT4: 3D unit-row mode with 16-byte aligned rows (D1 % 4 == 0, aligned sinogram pointer): the
//        window starts at a multiple of 4 bins and is flushed with red.global.add.v4.f32, one lane per
//        4 consecutive bins of one slice (REDG.E.ADD.F32x4: 1.8x the scalar RED rate, 4x fewer ops).
template <class G, bool IS3D, int S, int TN, int GS, int WIN, bool MAJOR_B, bool MINOR_UP, bool COLD, bool UNIT4,
          int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
walk_forward_kernel(Walk2Params wp, const float* __restrict__ in, float* __restrict__ sino) {
  static_assert(WIN % 32 == 0, "window is flushed 32 bins at a time");
  static_assert(S == 2 || S == 4, "window slots are float2 / float4");
  static_assert(!UNIT4 || (S == 4 && IS3D), "the vector flush is written for 4 slices");
  using Vec = typename SliceVec<S>::type;
  constexpr int Q = WIN / 32;
  constexpr int TM = 32 * GS;
  const PlaneParams& p = wp.p;
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
  float* win = smem + (size_t)warp * (WIN * S);
  Vec* winv = reinterpret_cast<Vec*>(win);

  // this thread's voxels, kept for every view of the launch, as slice PAIRS: the two bin sums of
  // two slices advance with one FFMA2 each (fma.rn.f32x2; per element identical to fmaf)
  constexpr int H = S / 2;
  float2 x[GS][TN][H];
#pragma unroll
  for (int d = 0; d < GS; ++d)
#pragma unroll
    for (int n = 0; n < TN; ++n) {
      const int a = MAJOR_B ? a0 + n : a0 + GS * lane + d;
      const int b = MAJOR_B ? b0 + GS * lane + d : b0 + n;
      const bool ok = a < p.NA && b < p.NB;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float val = (ok && s0 + s < p.NS) ? __ldg(in + ((size_t)(s0 + s) * p.NA + a) * (size_t)p.NB + b) : 0.f;
        if (s & 1) x[d][n][s / 2].y = val;
        else x[d][n][s / 2].x = val;
      }
    }

  // minor-axis coordinate of walk step 0 (a for MAJOR_B, b otherwise); + n is exact
  const float xmin0 = MAJOR_B ? G::coordA(a0) : G::coordB(b0);

  auto rmw = [&](int t, const float2 (&v)[H]) {  // win[t][:] += v   (one lane per address)
    Vec cur = winv[t];
    float2* c = reinterpret_cast<float2*>(&cur);
#pragma unroll
    for (int h = 0; h < H; ++h) c[h] = __fadd2_rn(c[h], v[h]);
    winv[t] = cur;
  };
  Vec vzero;
#pragma unroll
  for (int s = 0; s < S; ++s) reinterpret_cast<float*>(&vzero)[s] = 0.f;
  const float2 zero2 = make_float2(0.f, 0.f);

  const int v_begin = blockIdx.y * p.views_per_chunk;
  const int v_end = min(p.n_list, v_begin + p.views_per_chunk);
  for (int vi = v_begin; vi < v_end; ++vi) {
    const int v = p.view_list ? __ldg(p.view_list + vi) : vi;
    const ViewRec vr = load_view(p.views + v);
    int c0 = window_start<G>(vr, a0, a0 + (MAJOR_B ? TN : TM) - 1, b0, b0 + (MAJOR_B ? TM : TN) - 1);
    if (UNIT4) c0 &= ~3;

#pragma unroll
    for (int q = 0; q < Q; ++q) winv[lane + 32 * q] = vzero;
    __syncwarp();

#pragma unroll
    for (int d = 0; d < GS; ++d) {
      const float hMaj = MAJOR_B ? G::hoistB(vr, b0 + GS * lane + d) : G::hoistA(vr, a0 + GS * lane + d);
      float2 lo[H], hi[H];
      int tp = 0;
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        const float xm = xmin0 + (float)n;
        const float u = MAJOR_B ? G::combine(vr, G::hoistA_x(vr, xm), hMaj) : G::combine(vr, hMaj, G::hoistB_x(vr, xm));
        int c;
        float w0, w1;
        G::bins(vr, u, c, w0, w1);
        const float2 w0p = make_float2(w0, w0), w1p = make_float2(w1, w1);
        const int t = (int)min((unsigned)(c - c0), (unsigned)(WIN - 2));
        if (n == 0) {
#pragma unroll
          for (int h = 0; h < H; ++h) {
            lo[h] = __fmul2_rn(x[d][n][h], w0p);
            hi[h] = __fmul2_rn(x[d][n][h], w1p);
          }
        } else {
          if (t != tp) {
            // the bin that falls out of the carried pair is complete for this walk: flush it
            if (MINOR_UP) rmw(tp, lo);
            else rmw(tp + 1, hi);
            if (!COLD || (MINOR_UP ? (t == tp + 1) : (t == tp - 1))) {
#pragma unroll
              for (int h = 0; h < H; ++h) {
                if (MINOR_UP) { lo[h] = hi[h]; hi[h] = zero2; }
                else { hi[h] = lo[h]; lo[h] = zero2; }
              }
            } else {  // the bin moved by more than one (|c_minor| > 1): flush the other one too
              if (MINOR_UP) rmw(tp + 1, hi);
              else rmw(tp, lo);
#pragma unroll
              for (int h = 0; h < H; ++h) lo[h] = hi[h] = zero2;
            }
          }
          // order this step's stores before the next step's loads of other lanes
          __syncwarp();
#pragma unroll
          for (int h = 0; h < H; ++h) {
            lo[h] = __ffma2_rn(x[d][n][h], w0p, lo[h]);
            hi[h] = __ffma2_rn(x[d][n][h], w1p, hi[h]);
          }
        }
        tp = t;
      }
      rmw(tp, lo);
      __syncwarp();
      rmw(tp + 1, hi);
      __syncwarp();
    }

    // ---- flush the window to the sinogram
    if (UNIT4) {
      // lane j owns bins 4j .. 4j+3; column group is entirely inside or outside [0, D1)
      const int col = c0 + 4 * lane;
      if (lane < WIN / 4 && (unsigned)col < (unsigned)p.D1) {
        float blk[4][S];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Vec r = winv[4 * lane + k];
#pragma unroll
          for (int s = 0; s < S; ++s) blk[k][s] = reinterpret_cast<const float*>(&r)[s];
        }
        const long long* ro = wp.rowoff + (size_t)v * wp.row_stride + wp.s_base;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int sl = s0 + s;
          if (sl >= p.NS) break;
          const long long off = __ldg(ro + sl);
          const bool any = blk[0][s] != 0.f || blk[1][s] != 0.f || blk[2][s] != 0.f || blk[3][s] != 0.f;
          if (off >= 0 && any)
            red_add_v4(sino + off + col, wp.out_scale * blk[0][s], wp.out_scale * blk[1][s],
                       wp.out_scale * blk[2][s], wp.out_scale * blk[3][s]);
        }
      }
    } else {
      // bin (c0 + t) of slice s; rows / weights from the slice's row record
      RowRec rr[S];
      float* y0[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int sl = min(s0 + s, p.NS - 1);
        if (IS3D) {
          rr[s] = load_row(p.rows + (size_t)v * wp.row_stride + wp.s_base + sl);
          y0[s] = sino + ((size_t)v * p.D0 + rr[s].r0) * (size_t)p.D1;
        } else {
          rr[s].r0 = 0; rr[s].w0 = 1.f; rr[s].w1 = 0.f;
          y0[s] = sino + ((size_t)sl * p.V + v) * (size_t)p.D1;
        }
        if (s0 + s >= p.NS) rr[s].w0 = rr[s].w1 = 0.f;
      }
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int t = lane + 32 * q;
        const Vec r = winv[t];
        const int col = c0 + t;
        if (col >= 0 && col < p.D1) {
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float val = reinterpret_cast<const float*>(&r)[s];
            if (val != 0.f) {
              if (rr[s].w0 != 0.f) atomicAdd(y0[s] + col, rr[s].w0 * val);
              if (IS3D && rr[s].w1 != 0.f) atomicAdd(y0[s] + p.D1 + col, rr[s].w1 * val);
            }
          }
        }
      }
    }
    __syncwarp();  // all window reads done before the next view zeroes it
  }
}

// ------------------------------------------------------------------------ forward, joint columns
enum { ROWS_TABLE = 0, ROWS_KROW = 1, ROWS_MIX = 2 };
// Same tile and window as walk_forward_kernel, but a lane walks its TWO major-axis columns
// TOGETHER and carries THREE bin sums in registers: the column with the smaller coordinate (F) sits
// on bins (t, t+1), its neighbour (G, one voxel further along the major axis, |c_major| <= 1) on
// (t+e, t+e+1) with e in {0, 1}, so both fit bins t .. t+2.  One shared-memory read-modify-write
// per bin change then serves both columns: the shared-memory wavefronts that bound
// walk_forward_kernel (ncu: 14 per walk step, 39 % of them bank-conflict replays) drop by ~45 %.
// Views whose MAJOR coefficient is within rounding distance of 1 (ViewRec::fjump != 0, decided on the
// host with a rounding margin; up to 1.5) take a per-view variant in which G may land two bins after F.
// Only views whose MINOR coefficient can reach 1 as well (never the case for a rotation) go through
// walk_forward_kernel.  3D unit-row geometry with the 16-byte vector flush only.
// MAJ_POS: the major-axis coefficient is positive (F is the lane's first column).
// ROWS selects how a slice finds its detector row(s) at the flush:
//   ROWS_KROW   in every view the local row of slice i is i + ViewRec::krow (or outside the detector): the
//               four row pointers come from one base instead of four table loads;
//   ROWS_TABLE  unit rows looked up in the per-(view, slice) offset table;
//   ROWS_MIX    a slice spreads over rows r0 and r0 + 1 with the masked axis-0 weights of its RowRec
//               (detector row pitch != voxel pitch along axis 0); voxels are then not pre-scaled.
__device__ __forceinline__ void red_add_v4_if(bool pred, float* p, float a, float b, float c, float d) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %0, 0;\n"
      "@q red.global.add.v4.f32 [%1], {%2, %3, %4, %5};\n"
      "}\n" ::"r"((int)pred),
      "l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
      : "memory");
}

template <class G, int S, int TN, int WIN, bool MAJOR_B, bool MINOR_UP, bool MAJ_POS, int ROWS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2)  // two CTAs (16 warps) per SM: at most 128 registers
walk_forward_joint_kernel(Walk2Params wp, const float* __restrict__ in, float* __restrict__ sino) {
  static_assert(WIN % 32 == 0 && S == 4, "float4 window slots, flushed 4 bins per lane");
  using Vec = float4;
  constexpr int GS = 2, H = S / 2, Q = WIN / 32, TM = 32 * GS;
  constexpr int DF = MAJ_POS ? 0 : 1, DG = 1 - DF;
  const PlaneParams& p = wp.p;
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sgroups = (p.NS + S - 1) / S;
  const long long ntasks = (long long)sgroups * p.tilesA * p.tilesB;
  long long task = (long long)blockIdx.x * WARPS + warp;
  if (task >= ntasks) return;
  const int tb_ = (int)(task % p.tilesB);
  task /= p.tilesB;
  const int ta = (int)(task % p.tilesA);
  const int sg = (int)(task / p.tilesA);
  const int a0 = ta * (MAJOR_B ? TN : TM), b0 = tb_ * (MAJOR_B ? TM : TN), s0 = sg * S;
  Vec* winv = reinterpret_cast<Vec*>(smem + (size_t)warp * (WIN * S));

  // voxels are pre-multiplied by the axis-0 row weight (2, a power of two: bit-identical to scaling
  // the bin sums afterwards)
  float2 x[GS][TN][H];
  auto put = [&](int d, int n, int s, float val) {
    if (s & 1) x[d][n][s / 2].y = wp.out_scale * val;
    else x[d][n][s / 2].x = wp.out_scale * val;
  };
  // interior tiles of 16-byte aligned volumes are loaded with vector loads: the lane's 8 consecutive
  // columns as two LDG.128 (major axis A) or its two adjacent columns as one LDG.64 (major axis B)
  const bool interior = a0 + (MAJOR_B ? TN : TM) <= p.NA && b0 + (MAJOR_B ? TM : TN) <= p.NB && s0 + S <= p.NS &&
                        (p.NB & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  if (interior) {
    static_assert(TN == 8 && GS == 2, "vector tile load is written for 8 minor points and column pairs");
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float* slice = in + (size_t)(s0 + s) * p.NA * (size_t)p.NB;
      if (MAJOR_B) {
#pragma unroll
        for (int n = 0; n < TN; ++n) {
          const float2 v = __ldg(reinterpret_cast<const float2*>(slice + (size_t)(a0 + n) * p.NB + b0 + GS * lane));
          put(0, n, s, v.x);
          put(1, n, s, v.y);
        }
      } else {
#pragma unroll
        for (int d = 0; d < GS; ++d) {
          const float4* row = reinterpret_cast<const float4*>(slice + (size_t)(a0 + GS * lane + d) * p.NB + b0);
          const float4 v0 = __ldg(row), v1 = __ldg(row + 1);
          put(d, 0, s, v0.x); put(d, 1, s, v0.y); put(d, 2, s, v0.z); put(d, 3, s, v0.w);
          put(d, 4, s, v1.x); put(d, 5, s, v1.y); put(d, 6, s, v1.z); put(d, 7, s, v1.w);
        }
      }
    }
  } else {
#pragma unroll
    for (int d = 0; d < GS; ++d)
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        const int a = MAJOR_B ? a0 + n : a0 + GS * lane + d;
        const int b = MAJOR_B ? b0 + GS * lane + d : b0 + n;
        const bool ok = a < p.NA && b < p.NB;
#pragma unroll
        for (int s = 0; s < S; ++s)
          put(d, n, s, (ok && s0 + s < p.NS) ? __ldg(in + ((size_t)(s0 + s) * p.NA + a) * (size_t)p.NB + b) : 0.f);
      }
  }
  const float xmin0 = MAJOR_B ? G::coordA(a0) : G::coordB(b0);

  // (An XOR-swizzled slot layout that makes the final read conflict-free was measured: the two extra
  // address instructions per RMW cost more than the wavefronts saved -- the kernel is issue-bound.)
  auto rmw = [&](int t, const float2 (&v)[H]) {  // win[t][:] += v   (one lane per address)
    Vec* q = winv + t;
    Vec cur = *q;
    float2* c = reinterpret_cast<float2*>(&cur);
#pragma unroll
    for (int h = 0; h < H; ++h) c[h] = __fadd2_rn(c[h], v[h]);
    *q = cur;
  };
  const Vec vzero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float2 zero2 = make_float2(0.f, 0.f);

  const int v_begin = blockIdx.y * p.views_per_chunk;
  const int v_end = min(p.n_list, v_begin + p.views_per_chunk);
  for (int vi = v_begin; vi < v_end; ++vi) {
    const int v = p.view_list ? __ldg(p.view_list + vi) : vi;
    const ViewRec vr = load_view(p.views + v);
    const int c0 = window_start<G>(vr, a0, a0 + (MAJOR_B ? TN : TM) - 1, b0, b0 + (MAJOR_B ? TM : TN) - 1) & ~3;

#pragma unroll
    for (int q = 0; q < Q; ++q) winv[lane + 32 * q] = vzero;
    __syncwarp();

    const float hF = MAJOR_B ? G::hoistB(vr, b0 + GS * lane + DF) : G::hoistA(vr, a0 + GS * lane + DF);
    const float hG = MAJOR_B ? G::hoistB(vr, b0 + GS * lane + DG) : G::hoistA(vr, a0 + GS * lane + DG);
    // E2 (per view, warp-uniform: ViewRec::fjump): the major-axis coefficient is within rounding
    // distance of 1 (or up to 1.5), so G can land TWO bins after F.  Those rare points bypass the
    // carried triple and add G's two taps to the window directly (lanes sit >= 1 bin apart, so the
    // lanes that do it in one instruction still touch distinct bins); everything else is unchanged.
    auto walk_view = [&](auto e2_c) {
      constexpr bool E2 = decltype(e2_c)::value;
      float2 A0[H], A1[H], A2[H];  // sums of bins tb, tb + 1, tb + 2
      int tb = 0;
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        const float xm = xmin0 + (float)n;
        const float hm = MAJOR_B ? G::hoistA_x(vr, xm) : G::hoistB_x(vr, xm);
        const float uF = MAJOR_B ? G::combine(vr, hm, hF) : G::combine(vr, hF, hm);
        const float uG = MAJOR_B ? G::combine(vr, hm, hG) : G::combine(vr, hG, hm);
        int cF, cG;
        float wF0, wF1, wG0, wG1;
        G::bins(vr, uF, cF, wF0, wF1);
        G::bins(vr, uG, cG, wG0, wG1);
        const int tF = (int)min((unsigned)(cF - c0), (unsigned)(WIN - (E2 ? 4 : 3)));
        const bool e = cG != cF;  // G one bin further
        if (E2) {
          const bool far = cG - cF >= 2;
          if (__any_sync(0xffffffffu, far)) {
            if (far) {
              float2 g0[H], g1[H];
#pragma unroll
              for (int h = 0; h < H; ++h) {
                g0[h] = __fmul2_rn(x[DG][n][h], make_float2(wG0, wG0));
                g1[h] = __fmul2_rn(x[DG][n][h], make_float2(wG1, wG1));
              }
              rmw(tF + 2, g0);
              rmw(tF + 3, g1);
            }
            __syncwarp();
          }
          if (far) wG0 = wG1 = 0.f;  // G is accounted for
        }
        const float wa = e ? 0.f : wG0, wb = e ? wG0 : wG1, wc = e ? wG1 : 0.f;
        const float2 wF0p = make_float2(wF0, wF0), wF1p = make_float2(wF1, wF1);
        const float2 wap = make_float2(wa, wa), wbp = make_float2(wb, wb), wcp = make_float2(wc, wc);
        if (n == 0) {
          tb = tF;
#pragma unroll
          for (int h = 0; h < H; ++h) {
            A0[h] = __ffma2_rn(x[DG][n][h], wap, __fmul2_rn(x[DF][n][h], wF0p));
            A1[h] = __ffma2_rn(x[DG][n][h], wbp, __fmul2_rn(x[DF][n][h], wF1p));
            A2[h] = __fmul2_rn(x[DG][n][h], wcp);
          }
        } else {
          if (tF != tb) {  // F moved by one bin: the bin leaving the carried triple is complete for this walk
            if (MINOR_UP) {
              rmw(tb, A0);
#pragma unroll
              for (int h = 0; h < H; ++h) { A0[h] = A1[h]; A1[h] = A2[h]; A2[h] = zero2; }
            } else {
              rmw(tb + 2, A2);
#pragma unroll
              for (int h = 0; h < H; ++h) { A2[h] = A1[h]; A1[h] = A0[h]; A0[h] = zero2; }
            }
            tb = tF;
          }
          __syncwarp();  // order this step's stores before the next step's loads of other lanes
#pragma unroll
          for (int h = 0; h < H; ++h) {
            A0[h] = __ffma2_rn(x[DG][n][h], wap, __ffma2_rn(x[DF][n][h], wF0p, A0[h]));
            A1[h] = __ffma2_rn(x[DG][n][h], wbp, __ffma2_rn(x[DF][n][h], wF1p, A1[h]));
            A2[h] = __ffma2_rn(x[DG][n][h], wcp, A2[h]);
          }
        }
      }
      rmw(tb, A0);
      __syncwarp();
      rmw(tb + 1, A1);
      __syncwarp();
      rmw(tb + 2, A2);
      __syncwarp();
    };
    if (vr.fjump != 0.f) walk_view(std::true_type{});
    else walk_view(std::false_type{});

    // ---- flush the window: lane j owns bins 4j .. 4j+3 (entirely inside or outside [0, D1));
    // predicated REDs, no branches per slice
    const int col = c0 + 4 * lane;
    if (lane < WIN / 4 && (unsigned)col < (unsigned)p.D1) {
      float blk[4][S];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const Vec r = winv[4 * lane + k];
        blk[k][0] = r.x; blk[k][1] = r.y; blk[k][2] = r.z; blk[k][3] = r.w;
      }
      if (ROWS == ROWS_KROW) {
        const int r0 = wp.s_base + s0 + vr.krow;  // local detector row of the group's first slice
        float* y = sino + ((long long)v * p.D0 + r0) * (long long)p.D1 + col;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const unsigned any = __float_as_uint(blk[0][s]) | __float_as_uint(blk[1][s]) | __float_as_uint(blk[2][s]) |
                               __float_as_uint(blk[3][s]);  // all four +0: nothing to add
          const bool live = s0 + s < p.NS && (unsigned)(r0 + s) < (unsigned)p.D0 && any != 0u;
          red_add_v4_if(live, y + (long long)s * p.D1, blk[0][s], blk[1][s], blk[2][s], blk[3][s]);
        }
      } else if (ROWS == ROWS_TABLE) {
        const long long* ro = wp.rowoff + (size_t)v * wp.row_stride + wp.s_base;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int sl = min(s0 + s, p.NS - 1);
          const long long off = __ldg(ro + sl);
          const unsigned any = __float_as_uint(blk[0][s]) | __float_as_uint(blk[1][s]) | __float_as_uint(blk[2][s]) |
                               __float_as_uint(blk[3][s]);
          const bool live = s0 + s < p.NS && off >= 0 && any != 0u;
          red_add_v4_if(live, sino + (live ? off : 0) + col, blk[0][s], blk[1][s], blk[2][s], blk[3][s]);
        }
      } else {
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int sl = min(s0 + s, p.NS - 1);
          const RowRec rr = load_row(p.rows + (size_t)v * wp.row_stride + wp.s_base + sl);
          const unsigned any = __float_as_uint(blk[0][s]) | __float_as_uint(blk[1][s]) | __float_as_uint(blk[2][s]) |
                               __float_as_uint(blk[3][s]);
          const bool live = s0 + s < p.NS && any != 0u;
          // the masked weights are zero for rows outside the detector: r0 may then be -1 or D0 - 1
          float* y = sino + ((long long)v * p.D0 + rr.r0) * (long long)p.D1 + col;
          red_add_v4_if(live && rr.w0 != 0.f, y, rr.w0 * blk[0][s], rr.w0 * blk[1][s], rr.w0 * blk[2][s], rr.w0 * blk[3][s]);
          red_add_v4_if(live && rr.w1 != 0.f, y + p.D1, rr.w1 * blk[0][s], rr.w1 * blk[1][s], rr.w1 * blk[2][s],
                        rr.w1 * blk[3][s]);
        }
      }
    }
    __syncwarp();  // all window reads done before the next view zeroes it
  }
}

// ----------------------------------------------------------- forward, joint columns, CTA-shared tile
// walk_forward_joint_kernel keeps a warp's voxels in registers (64 of its 126), which caps a walk at TN = 8
// steps: per (view, tile) of 2048 updates the window is zeroed and flushed (25 % of the kernel's time), three
// end-of-walk read-modify-writes are paid (17 %) and the view preamble (8 %).  Here the CTA stages ONE tile of
// 64 (major) x TN = 64 (minor) x 4 slices in shared memory and its 12 warps run 12 different VIEWS on it, each with
// a private window: a walk is 64 steps, so the per-(view, tile) costs are spread over 16384 updates, and the
// registers that held voxels are free (more resident warps).  Price: the two columns' voxels of a step are two
// LDS.128 instead of register operands.  Layout of the tile: xt[n][column parity][lane] as float4 (4 slices):
// the lanes of one LDS.128 read 512 contiguous bytes.  Everything else -- the carried triple, the per-view E2
// variant, the vector flush -- is walk_forward_joint_kernel's.
// NVW views per warp pass (1 or 2): the kernel is bound by shared-memory wavefronts (ncu: l1tex data pipe ~90 % busy;
// per walk step of 256 updates 8 wavefronts are the two LDS.128 of the voxels and ~10 the window read-modify-writes).
// With NVW = 2 a warp walks TWO views of its class over the tile together: the voxels of a step are loaded once and
// feed both views' carried triples, i.e. 4 instead of 8 voxel wavefronts per 256 updates; the two walks are
// independent instruction streams (more ILP per warp), each with its own window.
template <class G, int S, int TN, int WIN, bool MAJOR_B, bool MINOR_UP, bool MAJ_POS, int ROWS, int WARPS, int MINB, int NVW = 1>
__global__ void __launch_bounds__(WARPS * 32, MINB)
walk_forward_tile_kernel(Walk2Params wp, const float* __restrict__ in, float* __restrict__ sino) {
  static_assert(WIN % 32 == 0 && WIN <= 128 && (S == 4 || S == 8), "float4 window slots, flushed 4 bins per lane in one pass");
  static_assert(NVW >= 1 && NVW <= 3, "one to three views per warp pass");
  using Vec = float4;
  // S slices = NV planes of 4: the tile and the windows are arrays of float4 per plane, so every shared access
  // stays a conflict-free LDS.128 / STS.128 over 512 contiguous bytes; the coordinates, bins and weights of a
  // walk step are evaluated once for all S slices (measured at C5: S = 8 is not faster than S = 4 -- the shared
  // wavefronts per update are the same)
  constexpr int NV = S / 4;
  constexpr int GS = 2, H = S / 2, Q = WIN / 32, TM = 32 * GS;
  constexpr int DF = MAJ_POS ? 0 : 1, DG = 1 - DF;
  constexpr int XPLANE = TN * 2 * 32;  // float4 slots per tile plane
  const PlaneParams& p = wp.p;
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Vec* xt = reinterpret_cast<Vec*>(smem);                                             // [NV][TN][2][32]
  Vec* win0 = reinterpret_cast<Vec*>(smem) + NV * XPLANE + warp * (NVW * NV * WIN);   // [NVW][NV][WIN]
  long long task = blockIdx.x;
  const int tb_ = (int)(task % p.tilesB);
  task /= p.tilesB;
  const int ta = (int)(task % p.tilesA);
  const int sg = (int)(task / p.tilesA);
  const int a0 = ta * (MAJOR_B ? TN : TM), b0 = tb_ * (MAJOR_B ? TM : TN), s0 = sg * S;

  // ---- stage the tile (pre-multiplied by the axis-0 row weight), every warp takes TN / WARPS minor rows
  for (int n = warp; n < TN; n += WARPS) {
#pragma unroll
    for (int d = 0; d < GS; ++d) {
      const int a = MAJOR_B ? a0 + n : a0 + GS * lane + d;
      const int b = MAJOR_B ? b0 + GS * lane + d : b0 + n;
      const bool ok = a < p.NA && b < p.NB;
      float v[S];
#pragma unroll
      for (int s = 0; s < S; ++s)
        v[s] = (ok && s0 + s < p.NS) ? wp.out_scale * __ldg(in + ((size_t)(s0 + s) * p.NA + a) * (size_t)p.NB + b) : 0.f;
#pragma unroll
      for (int pv = 0; pv < NV; ++pv)
        xt[pv * XPLANE + (n * 2 + d) * 32 + lane] = make_float4(v[4 * pv], v[4 * pv + 1], v[4 * pv + 2], v[4 * pv + 3]);
    }
  }
  __syncthreads();
  const float xmin0 = MAJOR_B ? G::coordA(a0) : G::coordB(b0);

  auto rmw = [&](Vec* winv, int t, const float2 (&v)[H]) {  // win[t][:] += v   (one lane per address)
#pragma unroll
    for (int pv = 0; pv < NV; ++pv) {
      Vec* q = winv + pv * WIN + t;
      Vec cur = *q;
      float2* c = reinterpret_cast<float2*>(&cur);
      c[0] = __fadd2_rn(c[0], v[2 * pv]);
      c[1] = __fadd2_rn(c[1], v[2 * pv + 1]);
      *q = cur;
    }
  };
  const Vec vzero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float2 zero2 = make_float2(0.f, 0.f);

  const int v_begin = blockIdx.y * p.views_per_chunk;
  const int v_end = min(p.n_list, v_begin + p.views_per_chunk);
  for (int vi = v_begin + NVW * warp; vi < v_end; vi += NVW * WARPS) {
    // the pass's views; a missing second view (odd count) repeats the first one and is not flushed
    int v[NVW];
    ViewRec vr[NVW];
    int c0[NVW];
    float hF[NVW], hG[NVW];
    bool e2 = false;
#pragma unroll
    for (int w = 0; w < NVW; ++w) {
      const int vj = min(vi + w, v_end - 1);
      v[w] = p.view_list ? __ldg(p.view_list + vj) : vj;
      vr[w] = load_view(p.views + v[w]);
      c0[w] = window_start<G>(vr[w], a0, a0 + (MAJOR_B ? TN : TM) - 1, b0, b0 + (MAJOR_B ? TM : TN) - 1) & ~3;
      hF[w] = MAJOR_B ? G::hoistB(vr[w], b0 + GS * lane + DF) : G::hoistA(vr[w], a0 + GS * lane + DF);
      hG[w] = MAJOR_B ? G::hoistB(vr[w], b0 + GS * lane + DG) : G::hoistA(vr[w], a0 + GS * lane + DG);
      e2 = e2 || vr[w].fjump != 0.f;
    }

#pragma unroll
    for (int q = 0; q < NVW * NV * Q; ++q) win0[lane + 32 * q] = vzero;
    __syncwarp();

    auto walk_views = [&](auto e2_c) {
      constexpr bool E2 = decltype(e2_c)::value;
      float2 A0[NVW][H], A1[NVW][H], A2[NVW][H];  // per view: sums of bins tb, tb + 1, tb + 2
      int tb[NVW];
#pragma unroll
      for (int w = 0; w < NVW; ++w) {
#pragma unroll
        for (int h = 0; h < H; ++h) A0[w][h] = A1[w][h] = A2[w][h] = zero2;
        // bin of F at the first step: the triple starts there (the loop below then never moves at n = 0)
        const float hm = MAJOR_B ? G::hoistA_x(vr[w], xmin0) : G::hoistB_x(vr[w], xmin0);
        const float uF = MAJOR_B ? G::combine(vr[w], hm, hF[w]) : G::combine(vr[w], hF[w], hm);
        tb[w] = (int)min((unsigned)(__float2int_rd(uF) - c0[w]), (unsigned)(WIN - (E2 ? 4 : 3)));
      }
      float xm = xmin0;  // minor-axis coordinate of the step; + 1 is exact
#pragma unroll(S == 4 && NVW == 1 ? 8 : 4)
      for (int n = 0; n < TN; ++n, xm += 1.0f) {
        float2 xF[H], xG[H];
#pragma unroll
        for (int pv = 0; pv < NV; ++pv) {
          const Vec xf4 = xt[pv * XPLANE + (n * 2 + DF) * 32 + lane], xg4 = xt[pv * XPLANE + (n * 2 + DG) * 32 + lane];
          xF[2 * pv] = make_float2(xf4.x, xf4.y);
          xF[2 * pv + 1] = make_float2(xf4.z, xf4.w);
          xG[2 * pv] = make_float2(xg4.x, xg4.y);
          xG[2 * pv + 1] = make_float2(xg4.z, xg4.w);
        }
#pragma unroll
        for (int w = 0; w < NVW; ++w) {
          Vec* winv = win0 + w * (NV * WIN);
          // the two columns' coordinates, bins and weights in packed fp32 (component-wise the scalar expressions of
          // Geom3::combine / bins; the products stay scalar, see the CAUTION in xct_geom.cuh)
          const float hm = MAJOR_B ? G::hoistA_x(vr[w], xm) : G::hoistB_x(vr[w], xm);
          const float2 hm2 = make_float2(hm, hm), hFG = make_float2(hF[w], hG[w]);
          const float2 u2 = MAJOR_B ? G::combine2(vr[w], hm2, hFG) : G::combine2(vr[w], hFG, hm2);
          int cF, cG;
          float2 w0_2, w1_2;
          G::bins2(vr[w], u2, cF, cG, w0_2, w1_2);
          const float wF0 = w0_2.x, wF1 = w1_2.x;
          float wG0 = w0_2.y, wG1 = w1_2.y;
          const int tF = (int)min((unsigned)(cF - c0[w]), (unsigned)(WIN - (E2 ? 4 : 3)));
          const bool e = cG != cF;  // G one bin further
          if (E2) {
            const bool far = cG - cF >= 2;
            if (__any_sync(0xffffffffu, far)) {
              if (far) {
                float2 g0[H], g1[H];
#pragma unroll
                for (int h = 0; h < H; ++h) {
                  g0[h] = __fmul2_rn(xG[h], make_float2(wG0, wG0));
                  g1[h] = __fmul2_rn(xG[h], make_float2(wG1, wG1));
                }
                rmw(winv, tF + 2, g0);
                rmw(winv, tF + 3, g1);
              }
              __syncwarp();
            }
            if (far) wG0 = wG1 = 0.f;  // G is accounted for
          }
          const float wa = e ? 0.f : wG0, wb = e ? wG0 : wG1, wc = e ? wG1 : 0.f;
          const float2 wF0p = make_float2(wF0, wF0), wF1p = make_float2(wF1, wF1);
          const float2 wap = make_float2(wa, wa), wbp = make_float2(wb, wb), wcp = make_float2(wc, wc);
          if (tF != tb[w]) {  // F moved by one bin: the bin leaving the carried triple is complete for this walk
            if (MINOR_UP) {
              rmw(winv, tb[w], A0[w]);
#pragma unroll
              for (int h = 0; h < H; ++h) { A0[w][h] = A1[w][h]; A1[w][h] = A2[w][h]; A2[w][h] = zero2; }
            } else {
              rmw(winv, tb[w] + 2, A2[w]);
#pragma unroll
              for (int h = 0; h < H; ++h) { A2[w][h] = A1[w][h]; A1[w][h] = A0[w][h]; A0[w][h] = zero2; }
            }
            tb[w] = tF;
          }
          // order this step's stores before the next step's loads of other lanes (windows are per view: one
          // barrier per step after the last view's stores is enough)
          if (w == NVW - 1) __syncwarp();
#pragma unroll
          for (int h = 0; h < H; ++h) {
            A0[w][h] = __ffma2_rn(xG[h], wap, __ffma2_rn(xF[h], wF0p, A0[w][h]));
            A1[w][h] = __ffma2_rn(xG[h], wbp, __ffma2_rn(xF[h], wF1p, A1[w][h]));
            A2[w][h] = __ffma2_rn(xG[h], wcp, A2[w][h]);
          }
        }
      }
#pragma unroll
      for (int w = 0; w < NVW; ++w) rmw(win0 + w * (NV * WIN), tb[w], A0[w]);
      __syncwarp();
#pragma unroll
      for (int w = 0; w < NVW; ++w) rmw(win0 + w * (NV * WIN), tb[w] + 1, A1[w]);
      __syncwarp();
#pragma unroll
      for (int w = 0; w < NVW; ++w) rmw(win0 + w * (NV * WIN), tb[w] + 2, A2[w]);
      __syncwarp();
    };
    if (e2) walk_views(std::true_type{});
    else walk_views(std::false_type{});

    // ---- flush the windows: lane j owns bins 4j .. 4j+3 (entirely inside or outside [0, D1))
#pragma unroll
    for (int w = 0; w < NVW; ++w) {
      if (w > 0 && vi + w >= v_end) break;  // repeated view of an odd tail
      const Vec* winv = win0 + w * (NV * WIN);
      const int col = c0[w] + 4 * lane;
      if (lane < WIN / 4 && (unsigned)col < (unsigned)p.D1) {
#pragma unroll
        for (int pv = 0; pv < NV; ++pv) {
          float blk[4][4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const Vec r = winv[pv * WIN + 4 * lane + k];
            blk[k][0] = r.x; blk[k][1] = r.y; blk[k][2] = r.z; blk[k][3] = r.w;
          }
          if (ROWS == ROWS_KROW) {
            const int r0 = wp.s_base + s0 + 4 * pv + vr[w].krow;  // local detector row of the plane's first slice
            float* y = sino + ((long long)v[w] * p.D0 + r0) * (long long)p.D1 + col;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const unsigned any = __float_as_uint(blk[0][s]) | __float_as_uint(blk[1][s]) | __float_as_uint(blk[2][s]) |
                                   __float_as_uint(blk[3][s]);  // all four +0: nothing to add
              const bool live = s0 + 4 * pv + s < p.NS && (unsigned)(r0 + s) < (unsigned)p.D0 && any != 0u;
              red_add_v4_if(live, y + (long long)s * p.D1, blk[0][s], blk[1][s], blk[2][s], blk[3][s]);
            }
          } else {
            const long long* ro = wp.rowoff + (size_t)v[w] * wp.row_stride + wp.s_base;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const int sl = min(s0 + 4 * pv + s, p.NS - 1);
              const long long off = __ldg(ro + sl);
              const unsigned any = __float_as_uint(blk[0][s]) | __float_as_uint(blk[1][s]) | __float_as_uint(blk[2][s]) |
                                   __float_as_uint(blk[3][s]);
              const bool live = s0 + 4 * pv + s < p.NS && off >= 0 && any != 0u;
              red_add_v4_if(live, sino + (live ? off : 0) + col, blk[0][s], blk[1][s], blk[2][s], blk[3][s]);
            }
          }
        }
      }
    }
    __syncwarp();  // all window reads done before the next pass zeroes them
  }
}

// ----------------------------------------------------------------- 2D forward, joint column pairs
// The joint-column walk for a single 2D image (S = 1, nothing to amortise coordinates over): a lane
// owns FOUR major-axis columns, i.e. two adjacent pairs, 4 voxels apart from its neighbour lane
// (bins >= 2 apart for the reference's default pixel size, |c_major| in [0.5, 0.71]: plain RMW, no
// atomics).  Each pair is walked along the minor axis with three carried bin sums exactly like
// walk_forward_joint_kernel (F on bins (t, t+1), G on (t+e, t+e+1), e in {0, 1}), the two columns'
// coordinates and weights evaluated in packed fp32 (component-wise identical to Geom2's scalar
// expressions; the row / column products stay scalar, see the CAUTION in xct_geom.cuh).  Against
// plane_forward_kernel (one RMW of an (A, B) slot pair per voxel, half of its shared wavefronts
// bank-conflict replays) the shared traffic drops to one RMW per bin change.
// Window: WIN bins (floats) per warp; flushed with scalar RED (2D rows have no 16-byte alignment).
// Only for plans whose every view has |coefficients| <= 1 - 5 ulp(u) (ViewRec::fjump == 0).
// (body shared by the per-class kernel and by the one-launch kernel for small problems below; block_x / chunk_y
// are the CTA's position in the class's own (blocks, view chunks) grid)
template <class G, int TN, int WIN, bool MAJOR_B, bool MINOR_UP, bool MAJ_POS, int WARPS>
__device__ __forceinline__ void walk2d_forward_joint_body(const PlaneParams& p, const float* __restrict__ in,
                                                          float* __restrict__ sino, float* smem, int block_x, int chunk_y) {
  static_assert(WIN % 32 == 0, "window is flushed 32 bins at a time");
  constexpr int GS = 4, TM = 32 * GS, Q = WIN / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ntasks = (long long)p.NS * p.tilesA * p.tilesB;
  long long task = (long long)block_x * WARPS + warp;
  if (task >= ntasks) return;
  const int tb_ = (int)(task % p.tilesB);
  task /= p.tilesB;
  const int ta = (int)(task % p.tilesA);
  const int sl = (int)(task / p.tilesA);  // image of the batch
  const int a0 = ta * (MAJOR_B ? TN : TM), b0 = tb_ * (MAJOR_B ? TM : TN);
  float* win = smem + (size_t)warp * WIN;

  float x[GS][TN];
#pragma unroll
  for (int d = 0; d < GS; ++d)
#pragma unroll
    for (int n = 0; n < TN; ++n) {
      const int a = MAJOR_B ? a0 + n : a0 + GS * lane + d;
      const int b = MAJOR_B ? b0 + GS * lane + d : b0 + n;
      x[d][n] = (a < p.NA && b < p.NB) ? __ldg(in + ((size_t)sl * p.NA + a) * (size_t)p.NB + b) : 0.f;
    }
  const float xmin0 = MAJOR_B ? G::coordA(a0) : G::coordB(b0);

  auto rmw = [&](int t, float v) { win[t] += v; };  // one lane per address

  const int v_begin = chunk_y * p.views_per_chunk;
  const int v_end = min(p.n_list, v_begin + p.views_per_chunk);
  for (int vi = v_begin; vi < v_end; ++vi) {
    const int v = p.view_list ? __ldg(p.view_list + vi) : vi;
    const ViewRec vr = load_view(p.views + v);
    const int c0 = window_start<G>(vr, a0, a0 + (MAJOR_B ? TN : TM) - 1, b0, b0 + (MAJOR_B ? TM : TN) - 1);
#pragma unroll
    for (int q = 0; q < Q; ++q) win[lane + 32 * q] = 0.f;
    __syncwarp();

#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {  // the lane's two column pairs
      constexpr int DF0 = MAJ_POS ? 0 : 1;
      const int dF = 2 * pr + DF0, dG = 2 * pr + (1 - DF0);
      const float hF = MAJOR_B ? G::hoistB(vr, b0 + GS * lane + dF) : G::hoistA(vr, a0 + GS * lane + dF);
      const float hG = MAJOR_B ? G::hoistB(vr, b0 + GS * lane + dG) : G::hoistA(vr, a0 + GS * lane + dG);
      const float2 hFG = make_float2(hF, hG);
      float A0 = 0.f, A1 = 0.f, A2 = 0.f;  // sums of bins tb, tb + 1, tb + 2
      int tb = 0;
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        const float xm = xmin0 + (float)n;
        const float hm = MAJOR_B ? G::hoistA_x(vr, xm) : G::hoistB_x(vr, xm);
        const float2 u = __fadd2_rn(hFG, make_float2(hm, hm));  // Geom2::combine: hA + hB (commutative)
        int cF, cG;
        float2 w0, w1;
        G::bins2(vr, u, cF, cG, w0, w1);
        const int tF = (int)min((unsigned)(cF - c0), (unsigned)(WIN - 3));
        const bool e = cG != cF;  // G one bin further
        const float wa = e ? 0.f : w0.y, wb = e ? w0.y : w1.y, wc = e ? w1.y : 0.f;
        const float xF = x[dF][n], xG = x[dG][n];
        if (n == 0) {
          tb = tF;
          A0 = fmaf(xG, wa, xF * w0.x);
          A1 = fmaf(xG, wb, xF * w1.x);
          A2 = xG * wc;
        } else {
          if (tF != tb) {  // F moved by one bin: the bin leaving the carried triple is complete for this walk
            if (MINOR_UP) { rmw(tb, A0); A0 = A1; A1 = A2; A2 = 0.f; }
            else { rmw(tb + 2, A2); A2 = A1; A1 = A0; A0 = 0.f; }
            tb = tF;
          }
          __syncwarp();  // order this step's stores before the next step's loads of other lanes
          A0 = fmaf(xG, wa, fmaf(xF, w0.x, A0));
          A1 = fmaf(xG, wb, fmaf(xF, w1.x, A1));
          A2 = fmaf(xG, wc, A2);
        }
      }
      rmw(tb, A0);
      __syncwarp();
      rmw(tb + 1, A1);
      __syncwarp();
      rmw(tb + 2, A2);
      __syncwarp();
    }

    float* y0 = sino + ((size_t)sl * p.V + v) * (size_t)p.D1;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int t = lane + 32 * q;
      const float val = win[t];
      const int col = c0 + t;
      if (val != 0.f && (unsigned)col < (unsigned)p.D1) atomicAdd(y0 + col, val);
    }
    __syncwarp();  // all window reads done before the next view zeroes it
  }
}

template <class G, int TN, int WIN, bool MAJOR_B, bool MINOR_UP, bool MAJ_POS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
walk2d_forward_joint_kernel(PlaneParams p, const float* __restrict__ in, float* __restrict__ sino) {
  extern __shared__ __align__(128) float smem[];
  walk2d_forward_joint_body<G, TN, WIN, MAJOR_B, MINOR_UP, MAJ_POS, WARPS>(p, in, sino, smem, blockIdx.x, blockIdx.y);
}

// ----------------------------------------------------- 2D forward, joint column pairs, CTA-shared tile
// walk2d_forward_joint_kernel keeps a warp's pixels in registers, which caps a walk at TN = 8 steps: per (view, tile)
// of 1024 updates the window is zeroed and flushed, six end-of-walk read-modify-writes are paid and the view
// preamble -- ~16 % of its instructions (it is issue-bound: ncu issue active 85 %).  Here, as in
// walk_forward_tile_kernel, the CTA stages ONE tile of 128 (major) x TN = 64 (minor) pixels in shared memory
// (xt[step][lane] as float4 = the lane's four columns: one conflict-free LDS.128 per step) and its warps run different
// VIEWS on it, each with a private window: the per-(view, tile) costs are spread over 8192 updates.  The two column
// pairs of a lane are walked together step by step (pair-major order would load the tile twice).
template <class G, int TN, int WIN, bool MAJOR_B, bool MINOR_UP, bool MAJ_POS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
walk2d_forward_tile_kernel(PlaneParams p, const float* __restrict__ in, float* __restrict__ sino) {
  static_assert(WIN % 32 == 0, "window is flushed 32 bins at a time");
  constexpr int GS = 4, TM = 32 * GS, Q = WIN / 32;
  constexpr int DF0 = MAJ_POS ? 0 : 1;
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* xt = reinterpret_cast<float4*>(smem);       // [TN][32]
  float* win = smem + TN * 32 * 4 + warp * WIN;       // [WIN]
  long long task = blockIdx.x;
  const int tb_ = (int)(task % p.tilesB);
  task /= p.tilesB;
  const int ta = (int)(task % p.tilesA);
  const int sl = (int)(task / p.tilesA);  // image of the batch
  const int a0 = ta * (MAJOR_B ? TN : TM), b0 = tb_ * (MAJOR_B ? TM : TN);

  // ---- stage the tile: xt[n][l] = the four major-axis columns 4l .. 4l+3 of minor row n
  const float* img = in + (size_t)sl * p.NA * (size_t)p.NB;
  if (MAJOR_B) {  // columns are consecutive in memory: one (vector) load per slot
    const bool vec = (p.NB & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    for (int n = warp; n < TN; n += WARPS) {
      const int a = a0 + n, b = b0 + GS * lane;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a < p.NA) {
        const float* q = img + (size_t)a * p.NB + b;
        if (vec && b + 3 < p.NB) {
          v = __ldg(reinterpret_cast<const float4*>(q));
        } else {
          if (b < p.NB) v.x = __ldg(q);
          if (b + 1 < p.NB) v.y = __ldg(q + 1);
          if (b + 2 < p.NB) v.z = __ldg(q + 2);
          if (b + 3 < p.NB) v.w = __ldg(q + 3);
        }
      }
      xt[n * 32 + lane] = v;
    }
  } else {  // minor axis is the contiguous one: coalesced reads along it, scattered shared stores (once per tile)
    float* xs = smem;
    for (int e = threadIdx.x; e < TM * TN; e += WARPS * 32) {
      const int n = e % TN, am = e / TN;  // minor index (contiguous in memory), major index
      const int a = a0 + am, b = b0 + n;
      xs[(n * 32 + (am >> 2)) * 4 + (am & 3)] = (a < p.NA && b < p.NB) ? __ldg(img + (size_t)a * p.NB + b) : 0.f;
    }
  }
  __syncthreads();
  const float xmin0 = MAJOR_B ? G::coordA(a0) : G::coordB(b0);

  auto rmw = [&](int t, float v) { win[t] += v; };  // one lane per address

  const int v_begin = blockIdx.y * p.views_per_chunk;
  const int v_end = min(p.n_list, v_begin + p.views_per_chunk);
  for (int vi = v_begin + warp; vi < v_end; vi += WARPS) {
    const int v = p.view_list ? __ldg(p.view_list + vi) : vi;
    const ViewRec vr = load_view(p.views + v);
    const int c0 = window_start<G>(vr, a0, a0 + (MAJOR_B ? TN : TM) - 1, b0, b0 + (MAJOR_B ? TM : TN) - 1);
#pragma unroll
    for (int q = 0; q < Q; ++q) win[lane + 32 * q] = 0.f;
    __syncwarp();

    float2 hFG[2];
    float A0[2], A1[2], A2[2];  // per pair: sums of bins tb, tb + 1, tb + 2
    int tb[2];
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      const int dF = 2 * pr + DF0, dG = 2 * pr + (1 - DF0);
      const float hF = MAJOR_B ? G::hoistB(vr, b0 + GS * lane + dF) : G::hoistA(vr, a0 + GS * lane + dF);
      const float hG = MAJOR_B ? G::hoistB(vr, b0 + GS * lane + dG) : G::hoistA(vr, a0 + GS * lane + dG);
      hFG[pr] = make_float2(hF, hG);
      A0[pr] = A1[pr] = A2[pr] = 0.f;
      // bin of F at the first step: the triple starts there (the loop below then never moves at n = 0)
      const float hm = MAJOR_B ? G::hoistA_x(vr, xmin0) : G::hoistB_x(vr, xmin0);
      tb[pr] = (int)min((unsigned)(__float2int_rd(G::combine(vr, MAJOR_B ? hm : hF, MAJOR_B ? hF : hm)) - c0), (unsigned)(WIN - 3));
    }
    float xm = xmin0;  // minor-axis coordinate of the step; + 1 is exact
#pragma unroll 4
    for (int n = 0; n < TN; ++n, xm += 1.0f) {
      const float4 xv = xt[n * 32 + lane];
      const float hm = MAJOR_B ? G::hoistA_x(vr, xm) : G::hoistB_x(vr, xm);
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        const float x0 = pr ? xv.z : xv.x, x1 = pr ? xv.w : xv.y;
        const float xF = DF0 ? x1 : x0, xG = DF0 ? x0 : x1;
        const float2 u = __fadd2_rn(hFG[pr], make_float2(hm, hm));  // Geom2::combine: hA + hB (commutative)
        int cF, cG;
        float2 w0, w1;
        G::bins2(vr, u, cF, cG, w0, w1);
        const int tF = (int)min((unsigned)(cF - c0), (unsigned)(WIN - 3));
        const bool e = cG != cF;  // G one bin further
        const float wa = e ? 0.f : w0.y, wb = e ? w0.y : w1.y, wc = e ? w1.y : 0.f;
        if (tF != tb[pr]) {  // F moved by one bin: the bin leaving the carried triple is complete for this walk
          if (MINOR_UP) { rmw(tb[pr], A0[pr]); A0[pr] = A1[pr]; A1[pr] = A2[pr]; A2[pr] = 0.f; }
          else { rmw(tb[pr] + 2, A2[pr]); A2[pr] = A1[pr]; A1[pr] = A0[pr]; A0[pr] = 0.f; }
          tb[pr] = tF;
        }
        __syncwarp();  // order this pair's stores before the next read-modify-write of other lanes
        A0[pr] = fmaf(xG, wa, fmaf(xF, w0.x, A0[pr]));
        A1[pr] = fmaf(xG, wb, fmaf(xF, w1.x, A1[pr]));
        A2[pr] = fmaf(xG, wc, A2[pr]);
      }
    }
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      rmw(tb[pr], A0[pr]);
      __syncwarp();
      rmw(tb[pr] + 1, A1[pr]);
      __syncwarp();
      rmw(tb[pr] + 2, A2[pr]);
      __syncwarp();
    }

    float* y0 = sino + ((size_t)sl * p.V + v) * (size_t)p.D1;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int t = lane + 32 * q;
      const float val = win[t];
      const int col = c0 + t;
      if (val != 0.f && (unsigned)col < (unsigned)p.D1) atomicAdd(y0 + col, val);
    }
    __syncwarp();  // all window reads done before the next view zeroes it
  }
}

// Small problems (BASELINE.json configs[1]: 512^2 x 360 views is ~0.1 ms of work): ONE launch for all eight
// (major axis, minor sign, major sign) view classes instead of one launch each -- every class brings its own
// ramp-up and tail, and four to eight of them are a third of the operator's time at this size.  A CTA looks its
// class up in a prefix table of CTA counts and runs that class's instantiation of the body.
struct Walk2dAllParams {
  PlaneParams p[8];      // per class: view list, tile grid, views per chunk
  int block_begin[9];    // CTAs of class k: [block_begin[k], block_begin[k + 1])
  int blocks_x[8];       // class k's CTA grid is blocks_x[k] (tiles) x chunks (view chunks)
};
template <class G, int TN, int WIN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
walk2d_forward_joint_all_kernel(const __grid_constant__ Walk2dAllParams ap, const float* __restrict__ in, float* __restrict__ sino) {
  extern __shared__ __align__(128) float smem[];
  int k = 0;
#pragma unroll
  for (int q = 1; q < 8; ++q) k += (int)blockIdx.x >= ap.block_begin[q] ? 1 : 0;
  const int local = (int)blockIdx.x - ap.block_begin[k];
  const int bx = local % ap.blocks_x[k], cy = local / ap.blocks_x[k];
  switch (k) {  // k = 4 * major_b + 2 * minor_up + major_positive
    case 0: walk2d_forward_joint_body<G, TN, WIN, false, false, false, WARPS>(ap.p[0], in, sino, smem, bx, cy); break;
    case 1: walk2d_forward_joint_body<G, TN, WIN, false, false, true, WARPS>(ap.p[1], in, sino, smem, bx, cy); break;
    case 2: walk2d_forward_joint_body<G, TN, WIN, false, true, false, WARPS>(ap.p[2], in, sino, smem, bx, cy); break;
    case 3: walk2d_forward_joint_body<G, TN, WIN, false, true, true, WARPS>(ap.p[3], in, sino, smem, bx, cy); break;
    case 4: walk2d_forward_joint_body<G, TN, WIN, true, false, false, WARPS>(ap.p[4], in, sino, smem, bx, cy); break;
    case 5: walk2d_forward_joint_body<G, TN, WIN, true, false, true, WARPS>(ap.p[5], in, sino, smem, bx, cy); break;
    case 6: walk2d_forward_joint_body<G, TN, WIN, true, true, false, WARPS>(ap.p[6], in, sino, smem, bx, cy); break;
    default: walk2d_forward_joint_body<G, TN, WIN, true, true, true, WARPS>(ap.p[7], in, sino, smem, bx, cy); break;
  }
}

}  // namespace xct
