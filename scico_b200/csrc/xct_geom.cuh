// Geometry policies shared by every kernel: the fp32 expression trees of the reference's
// _calc_weights, written with explicit round-to-nearest intrinsics so that ptxas can never
// contract a product and a sum into an FMA.  Bin indices (and, in 3D, the "left edge is an exact
// integer" case of _xray3d.py:224) are discontinuous in the coordinate, so the coordinate must be
// bit-identical to the oracle's; weights only need to agree to an ulp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace xct {

// One record per view, built on the host (see xct_api.cu) and read through the constant/L1 path.
struct __align__(16) ViewRec {
  float ca;      // coefficient of plane axis A   3D: M[1,1] (voxel axis 1)   2D: Pdx0 (image axis 0)
  float cb;      // coefficient of plane axis B   3D: M[1,2] (voxel axis 2)   2D: Pdx1 (image axis 1)
  float off;     // 3D: M[1,3]                                           2D: Pxmin
  float width;   // 2D: projected pixel width                             3D: unused
  float rwidth;  // 2D: 1/width (correctly rounded)                       3D: unused
  float jump;    // walk adjoint: != 0 when rounding can move the bin by two per row step (|ca| ~ 1)
  float fjump;   // walk forward: 0 = bins move by <= 1 per step; 1 = the major coefficient alone is within
                 // rounding distance of 1 (or up to 1.5): joint kernel's E2 variant; 2 = minor too: 2-bin walk
  int krow;      // walk adjoint (TMA): local detector row of slice i is i + krow in this view
};

// 3D separable geometry, plane = (voxel axis 1, voxel axis 2) -> detector column.
//   Px = ((M10*x0 + M11*x1) + M12*x2) + M13 with M10 == 0   (_xray3d.py:217)
//   left = Px - w/2, w = 0.5                                 (_xray3d.py:223)
struct Geom3 {
  static __device__ __forceinline__ float hoistA(const ViewRec& v, int a) {
    return __fmul_rn(v.ca, (float)a + 0.5f);
  }
  static __device__ __forceinline__ float hoistB(const ViewRec& v, int b) {
    return __fmul_rn(v.cb, (float)b + 0.5f);
  }
  // voxel-centre coordinate of row a, and hoistA from it (a + 0.5 + n is exact: walk kernels)
  static __device__ __forceinline__ float coordA(int a) { return (float)a + 0.5f; }
  static __device__ __forceinline__ float hoistA_x(const ViewRec& v, float xa) { return __fmul_rn(v.ca, xa); }
  static __device__ __forceinline__ float coordB(int b) { return (float)b + 0.5f; }
  static __device__ __forceinline__ float hoistB_x(const ViewRec& v, float xb) { return __fmul_rn(v.cb, xb); }
  static __device__ __forceinline__ float combine(const ViewRec& v, float hA, float hB) {
    return __fadd_rn(__fadd_rn(__fadd_rn(hA, hB), v.off), -0.25f);
  }
  // Two coordinates at once in packed fp32 (component-wise the IEEE round-to-nearest sums of combine /
  // bins; the products hA, hB stay scalar: ptxas would contract a packed product into a packed sum)
  static __device__ __forceinline__ float2 combine2(const ViewRec& v, float2 hA, float2 hB) {
    return __fadd2_rn(__fadd2_rn(__fadd2_rn(hA, hB), make_float2(v.off, v.off)), make_float2(-0.25f, -0.25f));
  }
  static __device__ __forceinline__ void bins2(const ViewRec&, float2 u, int& c0, int& c1, float2& w0, float2& w1) {
    c0 = __float2int_rd(u.x);
    c1 = __float2int_rd(u.y);
    const float2 d = __fadd2_rn(make_float2(ceilf(u.x), ceilf(u.y)), make_float2(-u.x, -u.y));
    w0 = make_float2(fminf(d.x, 0.5f), fminf(d.y, 0.5f));
    w1 = __fadd2_rn(make_float2(0.5f, 0.5f), make_float2(-w0.x, -w0.y));
  }
  // c = floor(left); w0 = to_next = min(ceil(left)-left, 0.5) (0 when left is an integer,
  // _xray3d.py:224); w1 = 0.5 - to_next.  The common factor 1/w^2 = 4 lives in the row weights.
  static __device__ __forceinline__ void bins(const ViewRec&, float u, int& c, float& w0, float& w1) {
    c = __float2int_rd(u);
    w0 = fminf(__fadd_rn(ceilf(u), -u), 0.5f);
    w1 = __fadd_rn(0.5f, -w0);
  }
};

// 2D geometry, plane = (image axis 0, image axis 1) -> detector bin.
//   Px = (Pxmin + Pdx0*i) + Pdx1*j                            (_xray2d.py:331-335)
struct Geom2 {
  static __device__ __forceinline__ float hoistA(const ViewRec& v, int a) {
    return __fadd_rn(v.off, __fmul_rn(v.ca, (float)a));
  }
  static __device__ __forceinline__ float hoistB(const ViewRec& v, int b) {
    return __fmul_rn(v.cb, (float)b);
  }
  static __device__ __forceinline__ float coordA(int a) { return (float)a; }
  static __device__ __forceinline__ float hoistA_x(const ViewRec& v, float xa) {
    return __fadd_rn(v.off, __fmul_rn(v.ca, xa));
  }
  static __device__ __forceinline__ float coordB(int b) { return (float)b; }
  static __device__ __forceinline__ float hoistB_x(const ViewRec& v, float xb) { return __fmul_rn(v.cb, xb); }
  static __device__ __forceinline__ float combine(const ViewRec&, float hA, float hB) {
    return __fadd_rn(hA, hB);
  }
  // Two rows at once with the packed fp32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: each
  // component is the IEEE round-to-nearest result, bit-identical to the scalar expression below;
  // fl - u is the exact negation of u - fl, so 1 + (fl - u) == 1 - (u - fl) bit for bit).
  // CAUTION: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (it does NOT do that to the
  // scalar .rn forms), so a packed product must never feed a packed sum where the reference rounds
  // twice: the row term off + ca * a stays scalar.
  static __device__ __forceinline__ float2 hoistA2(const ViewRec& v, float2 xa) {
    return make_float2(hoistA_x(v, xa.x), hoistA_x(v, xa.y));
  }
  static __device__ __forceinline__ void bins2(const ViewRec& v, float2 u, int& c0, int& c1, float2& w0, float2& w1) {
    c0 = __float2int_rd(u.x);
    c1 = __float2int_rd(u.y);
    // floor(u) as a float from the integer (I2FP, ALU pipe) instead of a second XU-pipe rounding op:
    // exact for |u| < 2^24, far beyond any detector (the 2D adjoint ran the XU pipe at 71 %)
    const float2 fl = make_float2(__int2float_rn(c0), __int2float_rn(c1));
    const float2 one = make_float2(1.0f, 1.0f), wd = make_float2(v.width, v.width), rw = make_float2(v.rwidth, v.rwidth);
    const float2 d = __fadd2_rn(one, __fadd2_rn(fl, make_float2(-u.x, -u.y)));
    const float2 m = make_float2(fminf(d.x, v.width), fminf(d.y, v.width));
    const float2 q = __fmul2_rn(m, rw);
    const float2 e = __ffma2_rn(make_float2(-q.x, -q.y), wd, m);
    w0 = __ffma2_rn(e, rw, q);
    w1 = __fadd2_rn(one, make_float2(-w0.x, -w0.y));
  }
  // inds = floor(Px); weights = min(1 - (Px - inds), width) / width   (_xray2d.py:338,348-349)
  static __device__ __forceinline__ void bins(const ViewRec& v, float u, int& c, float& w0, float& w1) {
    c = __float2int_rd(u);
    float fl = __int2float_rn(c);  // == floorf(u) for |u| < 2^24 (see bins2)
    float m = fminf(__fadd_rn(1.0f, -__fadd_rn(u, -fl)), v.width);
    // correctly rounded m/width from the host-rounded reciprocal plus one Markstein step
    float q = __fmul_rn(m, v.rwidth);
    float e = __fmaf_rn(-q, v.width, m);
    w0 = __fmaf_rn(e, v.rwidth, q);
    w1 = __fadd_rn(1.0f, -w0);
  }
};

// Row record of the 3D separable path: where voxel slice i of view v lands on the detector's
// axis 0 and with which weights (already multiplied by 1/w^2 = 4 and masked to 0 out of bounds).
struct __align__(16) RowRec {
  int32_t r0;  // LOCAL detector row of the first tap (may be -1 or d0-1 with a zero weight beside it)
  float w0;    // weight of row r0      = 4*to_next0          (0 if row r0 is out of bounds)
  float w1;    // weight of row r0 + 1  = 4*(0.5 - to_next0)  (0 if row r0+1 is out of bounds)
  int32_t pad;
};

// ---- routed output of the adjoint (view-block sharding, xct_adjoint_scatter) --------------------
// The adjoint's result is cut along its axis 0 into `nparts` row blocks; block k lives in the memory
// behind ptr[k] -- a buffer of the GPU that owns that block, reached through a CUDA-IPC mapping over
// NVLink when it is not the local one.  Two ways for the partial back projections of all view blocks
// to meet in the owner's memory, neither of which writes a partial volume for a later collective:
//   store (default of sharded.PeerBlocks): ptr[k] is THIS rank's slot in the owner's staging area; plain
//          posted stores (full NVLink write bandwidth); the owner sums the slots in rank order afterwards
//          (sum_slots_kernel: deterministic);
//   add:   ptr[k] is the block itself and every rank adds into it (RED.ADD.F32, system scope).
constexpr int kMaxRouteParts = 16;
struct OutRoute {
  float* ptr[kMaxRouteParts];
  int row_begin[kMaxRouteParts + 1];  // block k holds rows [row_begin[k], row_begin[k + 1])
  int nparts;
  int store;        // 1: plain stores (each element is written exactly once per launch); 0: RED.ADD
  long long inner;  // elements per row (product of the trailing dims)
};
// out[row][rest] (+)= val, in the memory of the part that owns `row`
__device__ __forceinline__ void route_add(const OutRoute& r, int row, long long rest, float val) {
  int k = 0;
#pragma unroll 1
  for (int q = 1; q < r.nparts; ++q) k += row >= r.row_begin[q] ? 1 : 0;
  float* q = r.ptr[k] + (long long)(row - r.row_begin[k]) * r.inner + rest;
  if (r.store) *q = val;
  else atomicAdd_system(q, val);  // system scope: the other GPUs of the node add to the same block
}

// Owner lookup done once for a run of rows that share it (see the routed epilogue of plane_adjoint_kernel).
struct RouteCursor {
  float* q;  // address of (row, rest) in the owner's memory
  int next;  // first row of the next block
  __device__ __forceinline__ void seek(const OutRoute& r, int row, long long rest) {
    int k = 0;
#pragma unroll 1
    for (int j = 1; j < r.nparts; ++j) k += row >= r.row_begin[j] ? 1 : 0;
    next = r.row_begin[k + 1];
    q = r.ptr[k] + (long long)(row - r.row_begin[k]) * r.inner + rest;
  }
};

// ---- rendezvous of the fused exchange through flags in peer memory (no collective library call) ----
// Every rank owns `n` flag words (one per peer) in memory all GPUs of the node have mapped.  After its routed
// back projection a rank SIGNALS: it writes the call's epoch into its word on every peer (release, system
// scope; the kernel boundary before it has already made the back projection's peer stores visible).  It then
// WAITS until all of its own words carry that epoch (acquire, system scope): every peer's back projection into
// this rank's staging area has completed.  Epochs only grow, so the words are never reset.
struct PeerFlagPtrs {
  int* ptr[kMaxRouteParts];  // ptr[k]: this rank's flag word in rank k's flag array
  int n;
};
__global__ void peer_signal_kernel(const __grid_constant__ PeerFlagPtrs f, int epoch) {
  const int k = threadIdx.x;
  if (k < f.n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f.ptr[k]), "r"(epoch) : "memory");
  }
}
// flags[k] >= epoch for every k < n, or *timed_out = 1 after `timeout_ns` (a peer died: the caller reports it
// instead of hanging the GPU)
__global__ void peer_wait_kernel(const int* flags, int n, int epoch, unsigned long long timeout_ns, int* timed_out) {
  const int k = threadIdx.x;
  if (k < n) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      int v;
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + k) : "memory");
      if (v - epoch >= 0) break;
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns) {
        *timed_out = 1;
        break;
      }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

// dst[i] = ((slot_0[i] + slot_1[i]) + slot_2[i]) + ... : the owner's side of the store-mode exchange.
__global__ void __launch_bounds__(256)
sum_slots_kernel(float* __restrict__ dst, const float* __restrict__ slots, int nslots, size_t n, size_t pitch) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((n & 3) == 0 && (pitch & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(slots)) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(slots);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const size_t n4 = n >> 2, p4 = pitch >> 2;
    for (; i < n4; i += stride) {
      float4 a = s4[i];
      for (int s = 1; s < nslots; ++s) {
        const float4 b = s4[(size_t)s * p4 + i];
        a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
      }
      d4[i] = a;
    }
    return;
  }
  for (; i < n; i += stride) {
    float a = slots[i];
    for (int s = 1; s < nslots; ++s) a = __fadd_rn(a, slots[(size_t)s * pitch + i]);
    dst[i] = a;
  }
}

}  // namespace xct
