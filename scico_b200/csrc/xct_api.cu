// C ABI of the B200-native X-ray projector pair: plan creation (host geometry analysis), kernel
// dispatch, host-buffer entry points.  See include/scico_b200_xray.h for the contract.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/scico_b200_xray.h"
#include "xct_general.cuh"
#include "xct_plane.cuh"
#include "xct_plane2.cuh"
#include "xct_brick.cuh"
#include "xct_tv.cuh"
#include "xct_solver.cuh"

namespace {

thread_local std::string g_err;
thread_local int64_t g_launches = 0;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define XCT_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(XCT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
  } while (0)

// ---- compile-time tile configuration of the plane kernels (see xct_plane.cuh) ----
constexpr int kWarps = 8;
// adjoint
constexpr int kAdjWin = 64;
constexpr int kAdj3S = 4, kAdj3TA = 16;
constexpr int kAdj2S = 1, kAdj2TA = 32;
constexpr int kAdj2bS = 4, kAdj2bTA = 16;  // 2D image batches (>= 3 images): coordinates and weights of a (view, pixel) serve 4 images
// forward
constexpr int kFwdWin = 96;
constexpr int kFwd3S = 2, kFwd3TN = 16;
constexpr int kFwd2S = 1, kFwd2TN = 16;
// walk kernels (xct_plane2.cuh)
// walk adjoint tile: 32 columns x TA rows x S slices per warp.  The coordinates of a row step are shared by its S
// slices and the per-view preamble by its TA rows; measured at C5 (tools/bench_adj_ab.py, ms per application):
// S x TA x ring stages  8 x 8 x 4: 229.6   12 x 6 x 3: 219.2   16 x 4 x 4: 214.9   16 x 4 x 3: 212.9   16 x 4 x 2: 212.6
// (window of 40 / 48 / 64 bins: no difference -- the L2 -> shared traffic of the windows is not the limiter)
#ifndef XCT_WADJ_S
#define XCT_WADJ_S 16
#define XCT_WADJ_TA 4
#define XCT_WADJ_STAGES 2
#endif
#ifndef XCT_WADJ_WIN
#define XCT_WADJ_WIN 64
#endif
constexpr int kWAdjS = XCT_WADJ_S, kWAdjTA = XCT_WADJ_TA, kWAdjWin = XCT_WADJ_WIN, kWAdjStages = XCT_WADJ_STAGES;
// slice-interleaved walk adjoint (walk_adjoint_vec_kernel): window of TA + 31 + 2 bins, no 4-bin alignment
#ifndef XCT_WVEC_S
#define XCT_WVEC_S 16
#define XCT_WVEC_TA 4
#define XCT_WVEC_WIN 40
#define XCT_WVEC_STAGES 3
#endif
constexpr int kWVecS = XCT_WVEC_S, kWVecTA = XCT_WVEC_TA, kWVecWin = XCT_WVEC_WIN, kWVecStages = XCT_WVEC_STAGES;
constexpr int kWFwdS = 4, kWFwdTN = 8;     // walk forward tile: 64 (major) x 8 (minor) x 4 slices
// CTA-shared-tile joint forward: 64 (major) x kWTileTN (minor) x 4 slices per CTA, kWTileWarps views in flight per CTA
#ifndef XCT_TILE_TN
#define XCT_TILE_TN 64    // measured at C5 (tools/bench_fwd_ab.py, ms per application): TN 16: 359, 32: 319, 64: 304
#define XCT_TILE_WIN 128  // window of the 64 x 64 tile: 63 (|c_major| + |c_minor|) + 7 <= 128 for every rotation
#define XCT_TILE_WARPS 10 // warps per CTA, 2 CTAs per SM; with two views per warp pass 20 views are in flight per CTA
#define XCT_TILE_MINB 2
#endif
#ifndef XCT_TILE_S
#define XCT_TILE_S 4      // slices per tile (4 or 8): coordinates / bins / weights of a walk step are shared by all of them
#endif
#ifndef XCT_TILE_NVW
// views a warp walks together over the tile (1 or 2): the voxel loads of a step are shared.  Measured at C5
// (ms per application; one view x 12 warps: 305.4 -> 301.1 with the two columns' coordinates in packed fp32):
// two views x 12 warps 295.8, two views x 10 warps 291.3; 8 slices x two views x 12 warps (1 CTA per SM) 296.3;
// THREE views x 8 warps (127 registers, 2 CTAs per SM) 290.6 against 291.2 in the same run, x 7 warps 306.1: no gain;
// shared-memory data pipe 89 % -> 71 % busy, issue slots 78 % -> 80 % (profiles/ncu_r02_walk.md)
#define XCT_TILE_NVW 2
#endif
constexpr int kWTileTN = XCT_TILE_TN, kWTileWin = XCT_TILE_WIN, kWTileWarps = XCT_TILE_WARPS, kWTileMinB = XCT_TILE_MINB;
constexpr int kWTileS = XCT_TILE_S, kWTileNVW = XCT_TILE_NVW;
constexpr int kW2dTN = 8, kW2dWin = 160;  // 2D joint forward tile: 128 (major) x 8 (minor), one image
// 2D joint forward on a CTA-shared tile: 128 (major) x kW2dTileTN (minor) pixels, kW2dTileWarps views in flight per CTA
#ifndef XCT_TILE2D_TN
#define XCT_TILE2D_TN 64
#define XCT_TILE2D_WIN 192   // 127 |c_major| + 63 |c_minor| + 4 <= 192 for every pixel size up to one bin per pixel
#define XCT_TILE2D_WARPS 8
#endif
constexpr int kW2dTileTN = XCT_TILE2D_TN, kW2dTileWin = XCT_TILE2D_WIN, kW2dTileWarps = XCT_TILE2D_WARPS;
// brick kernels (xct_brick.cuh): general 3D matrices
constexpr int kBrAdjWR = 20, kBrAdjWC = 24, kBrAdjStages = 4;  // adjoint window of an 8^3 brick (columns start at a multiple of 4), ring depth
constexpr int kBrFwdWR = 24, kBrFwdWC = 24;                    // forward window of an 8 x 16 x 4 brick

}  // namespace

struct xct_plan {
  bool dry = false;  // analysis only (xct*_plan_analyse): every decision is made, nothing is uploaded
  int adj_jump_views = 0;  // views that take the walk adjoint's jump-by-two variant
  int ndim = 0;
  int path = 0;          // XCT_PATH_*
  bool fwd_plane = false;  // forward uses the plane kernel (else general)
  bool adj_plane = false;
  int gs = 0;            // forward lane stride (2 or 3)
  int device = 0;
  int V = 0;
  // 2D: n0,n1 image, d1 = ny.  3D: n0,n1,n2 volume, d0,d1 detector
  int n0 = 0, n1 = 0, n2 = 1, d0 = 1, d1 = 0;
  int slice_offset = 0, row_off = 0, rows_total = 0;
  bool row_aligned = false;
  xct::ViewRec* d_views = nullptr;
  xct::RowRec* d_rows = nullptr;
  float* d_mats = nullptr;
  int* d_list[2] = {nullptr, nullptr};  // [0]: views whose major axis is A, [1]: major axis B
  // walk kernels: every (view, slice) lands in exactly one detector row with axis-0 weight 2
  bool rows_unit = false;
  bool adj_walk = false;
  bool rows_krow = false;  // local detector row of slice i is i + ViewRec::krow in every view (or none)
  bool adj_tma = false;    // walk adjoint stages its sinogram window with one TMA box per view (rows = slice + krow)
  bool adj_vec = false;    // slice-interleaved walk adjoint possible (adj_tma and every view's window fits kWVecWin)
  float* d_sinoT = nullptr;             // (V, ceil(n0 / 4), d1, 4) interleaved copy of the sinogram, made per call; allocated with the plan
  cudaEvent_t sinoT_ev = nullptr;       // last reader of d_sinoT (orders its reuse across streams)
  bool fwd_walk = false;
  bool fwd_cold = false;   // some view's minor-axis coefficient can move the bin by more than one per step
  bool fwd_unit4 = false;  // vector flush possible (unit rows, D1 % 4 == 0, window fits with 4-bin alignment)
  bool fwd_mix4 = false;   // same window / alignment conditions, rows mix two detector rows (joint kernel, ROWS_MIX)
  int* d_list4[4] = {nullptr, nullptr, nullptr, nullptr};  // walk forward: [2*major_b + minor_up]
  int n_list4[4] = {0, 0, 0, 0};
  // joint-column walk forward: views with fjump == 0 by [4*major_b + 2*minor_up + major_positive],
  // the remaining ("risky") views by [2*major_b + minor_up] for walk_forward_kernel
  bool fwd_joint = false;
  bool fwd_tile = false;     // joint forward with the CTA-shared tile (walk_forward_tile_kernel): unit rows, window fits TN = 32
  bool fwd_joint2d = false;  // 2D: every view inside walk2d_forward_joint_kernel's envelope
  bool fwd_tile2d = false;   // 2D: ... and inside walk2d_forward_tile_kernel's window (large problems run on the CTA-shared tile)
  bool fwd_2d_per_class = false;  // XCT_FLAG_2D_PER_CLASS: one launch per view class even for small problems (A/B)
  int* d_listJ[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int n_listJ[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int* d_listR[4] = {nullptr, nullptr, nullptr, nullptr};
  int n_listR[4] = {0, 0, 0, 0};
  long long* d_rowoff = nullptr;  // [V][n0] element offset of the (local) sinogram row, or -1
  // brick kernels (general 3D matrices): adjoint over all views; forward view lists by
  // [2 * depth_axis + needs_atomics]
  bool brick_adj = false, brick_adj_tma = false, brick_fwd = false;
  float* d_mats_t = nullptr;  // (V, 4, 2): matrices transposed (see xct_brick.cuh::load_mat)
  int* d_listB[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int n_listB[6] = {0, 0, 0, 0, 0, 0};
  int n_list[2] = {0, 0};
  // host-buffer staging (xct_*_host): one (input, output) pair of device buffers per direction, so that a
  // forward and an adjoint can be in flight together (xct_*_host_async)
  float* stage_in[2] = {nullptr, nullptr};   // [0]: forward's volume, [1]: adjoint's sinogram
  float* stage_out[2] = {nullptr, nullptr};  // [0]: forward's sinogram, [1]: adjoint's volume
  size_t cap_in[2] = {0, 0}, cap_out[2] = {0, 0};
  cudaStream_t hstream = nullptr;
  // pipelined host path (3D separable, unit rows monotone in the slice index): H2D of slice chunk
  // k+1 and D2H of chunk k-1 overlap the kernels of chunk k
  bool pipe_ok = false;
  std::vector<int> h_row_lo, h_row_hi;  // per slice: smallest / largest local detector row over views (-1: none)
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> events;
  cudaEvent_t done_ev[2] = {nullptr, nullptr};  // last D2H of the previous host call of each direction
};

namespace {

int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (dev < 0) return;  // analysis-only plan: no CUDA call at all
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(XCT_ERR_NO_DEVICE,
                "no CUDA device available: scico_b200 has no CPU fallback (" +
                    std::string(e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)) + ")");
  }
  if (device < 0 || device >= n) return fail(XCT_ERR_INVALID, "device ordinal out of range");
  return XCT_OK;
}

// Split the views into the two classes of the forward kernel and pick its lane stride.
// Returns false when some view is outside the plane kernels' envelope.
struct Envelope {
  bool adj_ok = true, fwd_ok = true, adj_walk_ok = true, adj_vec_ok = true, fwd_unit4_ok = true;
  float max_minor = 0.f;
  int gs = 2;
  std::vector<int> list[2];
  std::vector<int> list4[4];  // [2*major_b + minor_up]: minor-axis coefficient >= 0
  std::vector<int> listJ[8];  // fjump == 0: [4*major_b + 2*minor_up + major_positive]
  std::vector<int> listR[4];  // fjump != 0: [2*major_b + minor_up]
};

Envelope analyse_views(const std::vector<xct::ViewRec>& views, int adjTA, int fwdTN) {
  Envelope env;
  float min_major = 1e30f;
  std::vector<int> zero_minor[2];
  std::vector<int> zero_minor_j[4];  // joint lists: [2*major_b + major_positive]
  for (size_t v = 0; v < views.size(); ++v) {
    const float a = std::fabs(views[v].ca), b = std::fabs(views[v].cb);
    if (!std::isfinite(a) || !std::isfinite(b)) {
      env.adj_ok = env.fwd_ok = env.adj_walk_ok = env.adj_vec_ok = false;
      continue;
    }
    // adjoint window: tile adjTA x 32, needs floor(max u) - floor(min u) + 2 <= WIN
    if (a * (adjTA - 1) + b * 31.f + 3.f > (float)kAdjWin) env.adj_ok = false;
    // walk adjoint: window start is rounded down to a multiple of 4 bins (+3)
    if (a * (kWAdjTA - 1) + b * 31.f + 6.f > (float)kWAdjWin) env.adj_walk_ok = false;
    // slice-interleaved walk adjoint: exact window start, two taps, one bin of rounding slack
    if (a * (kWVecTA - 1) + b * 31.f + 3.f > (float)kWVecWin) env.adj_vec_ok = false;
    const bool major_b = b >= a;
    env.list[major_b ? 1 : 0].push_back((int)v);
    const float minor = major_b ? views[v].ca : views[v].cb;
    if (minor == 0.f) zero_minor[major_b ? 1 : 0].push_back((int)v);  // no bin movement: either sign class
    else env.list4[(major_b ? 2 : 0) + (minor > 0.f ? 1 : 0)].push_back((int)v);
    {
      const float major = major_b ? views[v].cb : views[v].ca;
      const int up = minor >= 0.f ? 1 : 0;
      // joint-column kernel: safe views, and views whose MAJOR coefficient alone is near (or up to 1.5x)
      // a bin per voxel (per-view E2 variant); a minor coefficient that can reach 1 stays on the 2-bin walk
      // (fjump: 0 = every bin step is at most one, 1 = only the major coefficient is near / above one bin
      // per voxel, up to 1.5: the kernel's per-view E2 variant, 2 = the minor coefficient can reach 1 too)
      if (views[v].fjump < 2.f) {
        if (minor == 0.f) zero_minor_j[(major_b ? 2 : 0) + (major > 0.f ? 1 : 0)].push_back((int)v);  // either sign class
        else env.listJ[(major_b ? 4 : 0) + 2 * up + (major > 0.f ? 1 : 0)].push_back((int)v);
      } else {
        env.listR[(major_b ? 2 : 0) + up].push_back((int)v);
      }
    }
    min_major = std::min(min_major, std::max(a, b));
  }
  for (int m = 0; m < 2; ++m) {  // views with a zero minor coefficient join the larger sign class
    auto& dst = env.list4[2 * m + (env.list4[2 * m + 1].size() >= env.list4[2 * m].size() ? 1 : 0)];
    dst.insert(dst.end(), zero_minor[m].begin(), zero_minor[m].end());
    std::sort(dst.begin(), dst.end());
  }
  for (int m = 0; m < 4; ++m) {  // joint lists: a view without bin movement joins the larger sign class (no launch of its own)
    const int base = (m >> 1) * 4 + (m & 1);  // 4*major_b + major_positive, minor_up adds 2
    auto& dst = env.listJ[base + (env.listJ[base + 2].size() >= env.listJ[base].size() ? 2 : 0)];
    dst.insert(dst.end(), zero_minor_j[m].begin(), zero_minor_j[m].end());
    std::sort(dst.begin(), dst.end());
  }
  // lanes GS voxels apart along the major axis must land >= 1 bin apart
  if (2.f * min_major >= 1.01f) env.gs = 2;
  else if (3.f * min_major >= 1.01f) env.gs = 3;
  else env.fwd_ok = false;
  if (env.fwd_ok) {
    for (const auto& vr : views) {
      const float a = std::fabs(vr.ca), b = std::fabs(vr.cb);
      const float mj = std::max(a, b), mn = std::min(a, b);
      if (mj * (32 * env.gs - 1) + mn * (fwdTN - 1) + 3.f > (float)kFwdWin) env.fwd_ok = false;
      // walk forward with the window start rounded down to a multiple of 4 bins
      if (mj * (32 * env.gs - 1) + mn * (kWFwdTN - 1) + 6.f > (float)kFwdWin) env.fwd_unit4_ok = false;
      env.max_minor = std::max(env.max_minor, mn);
    }
  }
  return env;
}

// device copy of a host table; a dry plan (analysis only, possibly without any CUDA device) skips it
template <class T>
cudaError_t dev_upload(const xct_plan* pl, T*& dst, const T* src, size_t count) {
  if (pl->dry || count == 0) return cudaSuccess;
  cudaError_t e = cudaMalloc(&dst, sizeof(T) * count);
  if (e == cudaSuccess) e = cudaMemcpy(dst, src, sizeof(T) * count, cudaMemcpyHostToDevice);
  return e;
}

int upload_lists(xct_plan* pl, const Envelope& env) {
  auto up = [&](const std::vector<int>& src, int*& dst, int& n) -> cudaError_t {
    n = (int)src.size();
    return dev_upload(pl, dst, src.data(), src.size());
  };
  for (int c = 0; c < 2; ++c) XCT_CUDA(up(env.list[c], pl->d_list[c], pl->n_list[c]));
  for (int c = 0; c < 4; ++c) XCT_CUDA(up(env.list4[c], pl->d_list4[c], pl->n_list4[c]));
  for (int c = 0; c < 8; ++c) XCT_CUDA(up(env.listJ[c], pl->d_listJ[c], pl->n_listJ[c]));
  for (int c = 0; c < 4; ++c) XCT_CUDA(up(env.listR[c], pl->d_listR[c], pl->n_listR[c]));
  return XCT_OK;
}

template <class K>
int launch_check(const char* name) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(XCT_ERR_CUDA, std::string(name) + " launch: " + cudaGetErrorString(e));
  return XCT_OK;
}
int launch_ok(const char* name) { return launch_check<void>(name); }

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link against libcuda,
// so the library still loads on a machine without a driver)
using TensorMapEncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      f = nullptr;
    }
    return reinterpret_cast<TensorMapEncodeFn>(f);
  }();
  return fn;
}

int general_grid(size_t n) {
  size_t blocks = (n + 255) / 256;
  return (int)std::min<size_t>(blocks, 148u * 64u);
}

xct::PlaneParams plane_params(const xct_plan* pl, int batch) {
  xct::PlaneParams p{};
  p.views = pl->d_views;
  p.rows = pl->d_rows;
  p.view_list = nullptr;
  p.n_list = pl->V;
  p.V = pl->V;
  if (pl->ndim == 3) {
    p.NA = pl->n1; p.NB = pl->n2; p.NS = pl->n0; p.D0 = pl->d0; p.D1 = pl->d1;
  } else {
    p.NA = pl->n0; p.NB = pl->n1; p.NS = batch; p.D0 = 1; p.D1 = pl->d1;
  }
  p.views_per_chunk = pl->V;
  return p;
}

// ------------------------------------------------------------------ plane launches
template <class G, bool IS3D, int S, int TA>
int launch_plane_adjoint(const xct_plan* pl, int batch, const float* in, float* out, cudaStream_t st,
                         const xct::OutRoute* route = nullptr) {
  xct::PlaneParams p = plane_params(pl, batch);
  p.tilesA = ceil_div(p.NA, TA);
  p.tilesB = ceil_div(p.NB, 32);
  const long long tasks = (long long)ceil_div(p.NS, S) * p.tilesA * p.tilesB;
  const int blocks = ceil_div(tasks, kWarps);
  // small problems: split the views over blockIdx.y (partial sums meet in the zeroed output via RED)
  long long target_warps = 148LL * 64;  // two waves of resident warps; measured at C2: 1184 / 2368 / 3552 / 4736 / 7104 / 9472 / 14208 -> 134 / 103 / 98 / 90 / 85 / 82 / 83 us
  if (const char* env = std::getenv("XCT_ADJ_TARGET_WARPS")) target_warps = std::max(1, std::atoi(env));  // tuning (tools/bench_c2.py)
  int chunks = 1;
  if (tasks < target_warps) chunks = (int)std::min<long long>((target_warps + tasks - 1) / tasks, std::max(1, p.n_list / 8));
  p.views_per_chunk = ceil_div(p.n_list, chunks);
  chunks = ceil_div(p.n_list, p.views_per_chunk);
  const size_t smem = (size_t)kWarps * 2 * S * kAdjWin * sizeof(float);
  if (route && route->store) {  // plain stores: every element must be written exactly once, so no view split
    chunks = 1;
    p.views_per_chunk = p.n_list;
  }
  if (route) {  // routed: every value goes to its row block's owner (add mode: the caller zeroed the blocks)
    xct::plane_adjoint_kernel<G, IS3D, S, TA, kAdjWin, kWarps, true>
        <<<dim3(blocks, chunks), kWarps * 32, smem, st>>>(p, in, nullptr, *route);
    return launch_ok("plane_adjoint_kernel<route>");
  }
  if (chunks > 1) {
    const size_t n_out = (size_t)p.NS * p.NA * p.NB;
    XCT_CUDA(cudaMemsetAsync(out, 0, n_out * sizeof(float), st));
  }
  xct::plane_adjoint_kernel<G, IS3D, S, TA, kAdjWin, kWarps><<<dim3(blocks, chunks), kWarps * 32, smem, st>>>(p, in, out, xct::OutRoute{});
  return launch_ok("plane_adjoint_kernel");
}

// Slice-interleaved walk adjoint: interleave the detector rows of the launch's slices four by four into the plan's
// scratch (sino_interleave4_kernel), then walk_adjoint_vec_kernel with one TMA box of that copy per (view, tile).
// Returns 1 when this path cannot be taken (tensor map not encodable): the caller then launches the scalar-tap kernel.
int launch_walk_adjoint_vec(const xct_plan* cpl, const float* in, float* out, cudaStream_t st, int s_begin, int s_count,
                            const xct::OutRoute* route) {
  xct_plan* pl = const_cast<xct_plan*>(cpl);  // the scratch and its event are caches, not plan state
  const int g_total = ceil_div(pl->n0, 4);
  if (!pl->d_sinoT) return 1;
  if (pl->V > 65535 || ceil_div(pl->n0, 4) > 65535) return 1;  // grid limits of the interleaving pass
  // inside a stream capture (CUDA graphs, XLA command buffers) the cross-stream event is neither waited on nor
  // recorded: a replayed graph is ordered by its own stream like any other kernel sequence
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  const bool capturing = cap != cudaStreamCaptureStatusNone;
  xct::Walk2Params wp{};
  wp.p = plane_params(pl, 1);
  wp.rowoff = pl->d_rowoff;
  wp.out_scale = 2.0f;
  wp.row_stride = pl->n0;
  wp.s_base = s_begin;
  xct::PlaneParams& p = wp.p;
  if (s_count >= 0) p.NS = s_count;
  if (out) out += (size_t)s_begin * pl->n1 * pl->n2;
  p.tilesA = ceil_div(p.NA, kWVecTA);
  p.tilesB = ceil_div(p.NB, 32);
  const long long tasks = (long long)ceil_div(p.NS, kWVecS) * p.tilesA * p.tilesB;
  const int blocks = ceil_div(tasks, kWarps);
  // (V, g_total, 4 * d1) fp32, box = 4 * kWVecWin floats x kWVecS / 4 groups x 1 view, zero fill out of bounds
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  {
    const cuuint64_t dims[3] = {(cuuint64_t)pl->d1 * 4, (cuuint64_t)g_total, (cuuint64_t)pl->V};
    const cuuint64_t strides[2] = {(cuuint64_t)pl->d1 * 4 * sizeof(float), (cuuint64_t)g_total * pl->d1 * 4 * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)kWVecWin * 4, (cuuint32_t)kWVecS / 4, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, pl->d_sinoT, dims, strides, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1;
  }
  // the previous reader of the scratch (possibly on another stream) must have finished
  if (!capturing) XCT_CUDA(cudaStreamWaitEvent(st, pl->sinoT_ev, 0));
  {
    const int g_begin = s_begin / 4, g_count = ceil_div(p.NS, 4);
    const dim3 grid(ceil_div(pl->d1, 256), g_count, pl->V);
    xct::sino_interleave4_kernel<<<grid, 256, 0, st>>>(pl->d_views, in, pl->d_sinoT, pl->d0, pl->d1, g_begin, g_count, g_total);
    int rc = launch_ok("sino_interleave4_kernel");
    if (rc) return rc;
  }
  const size_t smem = (size_t)kWarps * kWVecStages * kWVecS * kWVecWin * sizeof(float) + (size_t)kWarps * kWVecStages * sizeof(unsigned long long);
  const xct::OutRoute none{};
  if (route) {
    auto kern = xct::walk_adjoint_vec_kernel<xct::Geom3, kWVecS, kWVecTA, kWVecWin, kWVecStages, kWarps, true>;
    if (smem > 48 * 1024) XCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, kWarps * 32, smem, st>>>(wp, out, tmap, *route);
  } else {
    auto kern = xct::walk_adjoint_vec_kernel<xct::Geom3, kWVecS, kWVecTA, kWVecWin, kWVecStages, kWarps, false>;
    if (smem > 48 * 1024) XCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, kWarps * 32, smem, st>>>(wp, out, tmap, none);
  }
  int rc = launch_ok("walk_adjoint_vec_kernel");
  if (rc) return rc;
  if (!capturing) XCT_CUDA(cudaEventRecord(pl->sinoT_ev, st));
  return XCT_OK;
}

// walk adjoint (3D separable geometry with unit rows; `in` must be 16-byte aligned)
// Slices [s_begin, s_begin + s_count) of the plan only (s_count < 0: all); `out` is the full volume.
int launch_walk_adjoint(const xct_plan* pl, const float* in, float* out, cudaStream_t st, int s_begin = 0,
                        int s_count = -1, const xct::OutRoute* route = nullptr) {
  if (pl->adj_vec && (s_begin & 3) == 0) {
    const int rc = launch_walk_adjoint_vec(pl, in, out, st, s_begin, s_count, route);
    if (rc != 1) return rc;
  }
  xct::Walk2Params wp{};
  wp.p = plane_params(pl, 1);
  wp.rowoff = pl->d_rowoff;
  wp.out_scale = 2.0f;
  wp.row_stride = pl->n0;
  wp.s_base = s_begin;
  xct::PlaneParams& p = wp.p;
  if (s_count >= 0) p.NS = s_count;
  if (out) out += (size_t)s_begin * pl->n1 * pl->n2;
  p.tilesA = ceil_div(p.NA, kWAdjTA);
  p.tilesB = ceil_div(p.NB, 32);
  const long long tasks = (long long)ceil_div(p.NS, kWAdjS) * p.tilesA * p.tilesB;
  const int blocks = ceil_div(tasks, kWarps);
  const size_t smem = (size_t)kWarps * kWAdjStages * kWAdjS * kWAdjWin * sizeof(float);
  const xct::OutRoute none{};
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  bool tma = pl->adj_tma;
  if (tma) {
    // (V, d0, d1) fp32 sinogram, box = kWAdjWin bins x kWAdjS rows x 1 view, zero fill out of bounds
    const cuuint64_t dims[3] = {(cuuint64_t)pl->d1, (cuuint64_t)pl->d0, (cuuint64_t)pl->V};
    const cuuint64_t strides[2] = {(cuuint64_t)pl->d1 * sizeof(float), (cuuint64_t)pl->d0 * pl->d1 * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)kWAdjWin, (cuuint32_t)kWAdjS, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), dims, strides,
                                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) tma = false;  // e.g. a stride the tensor map cannot express: cp.async staging below
  }
  const size_t smem_all = smem + (tma ? (size_t)kWarps * kWAdjStages * sizeof(unsigned long long) : 0);
#define XCT_WALK_ADJ(TMA_, ROUTE_)                                                                                                \
  do {                                                                                                                            \
    auto kern = xct::walk_adjoint_kernel<xct::Geom3, true, kWAdjS, kWAdjTA, kWAdjWin, kWAdjStages, kWarps, TMA_, ROUTE_>;         \
    if (smem_all > 48 * 1024) XCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_all));   \
    kern<<<blocks, kWarps * 32, smem_all, st>>>(wp, in, out, tmap, route ? *route : none);                                        \
  } while (0)
  if (tma) { if (route) XCT_WALK_ADJ(true, true); else XCT_WALK_ADJ(true, false); }
  else { if (route) XCT_WALK_ADJ(false, true); else XCT_WALK_ADJ(false, false); }
#undef XCT_WALK_ADJ
  return launch_ok(tma ? "walk_adjoint_kernel<tma>" : "walk_adjoint_kernel");
}

template <class G, bool IS3D, int S, int TN, int GS, bool MAJOR_B>
int launch_plane_forward_class(const xct_plan* pl, int batch, const float* in, float* out, cudaStream_t st) {
  const int cls = MAJOR_B ? 1 : 0;
  if (pl->n_list[cls] == 0) return XCT_OK;
  xct::PlaneParams p = plane_params(pl, batch);
  p.view_list = pl->d_list[cls];
  p.n_list = pl->n_list[cls];
  constexpr int TM = 32 * GS;
  p.tilesA = ceil_div(p.NA, MAJOR_B ? TN : TM);
  p.tilesB = ceil_div(p.NB, MAJOR_B ? TM : TN);
  const long long tasks = (long long)ceil_div(p.NS, S) * p.tilesA * p.tilesB;
  const int blocks = ceil_div(tasks, kWarps);
  // small problems: split the view list over blockIdx.y so the grid fills the 148 SMs
  const long long target_warps = 148LL * 32;
  int chunks = 1;
  if (tasks < target_warps) chunks = (int)std::min<long long>((target_warps + tasks - 1) / tasks, std::max(1, p.n_list / 4));
  p.views_per_chunk = ceil_div(p.n_list, chunks);
  chunks = ceil_div(p.n_list, p.views_per_chunk);
  const size_t smem = (size_t)kWarps * S * xct::FwdSlots<kFwdWin>::SIZE * sizeof(float2);
  dim3 grid(blocks, chunks);
  xct::plane_forward_kernel<G, IS3D, S, TN, GS, kFwdWin, MAJOR_B, kWarps><<<grid, kWarps * 32, smem, st>>>(p, in, out);
  return launch_ok("plane_forward_kernel");
}

template <class G, bool IS3D, int S, int TN>
int launch_plane_forward(const xct_plan* pl, int batch, const float* in, float* out, cudaStream_t st) {
  int rc;
  if (pl->gs == 2) {
    if ((rc = launch_plane_forward_class<G, IS3D, S, TN, 2, true>(pl, batch, in, out, st))) return rc;
    return launch_plane_forward_class<G, IS3D, S, TN, 2, false>(pl, batch, in, out, st);
  }
  if ((rc = launch_plane_forward_class<G, IS3D, S, TN, 3, true>(pl, batch, in, out, st))) return rc;
  return launch_plane_forward_class<G, IS3D, S, TN, 3, false>(pl, batch, in, out, st);
}

// walk forward: one launch per (major axis, minor-axis sign) class
// Launch geometry shared by the walk forward kernels (tile = 64 major x TN minor x S slices per warp).
template <int S, int TN, int GS, bool MAJOR_B>
dim3 walk_forward_grid(xct::PlaneParams& p) {
  constexpr int TM = 32 * GS;
  p.tilesA = ceil_div(p.NA, MAJOR_B ? TN : TM);
  p.tilesB = ceil_div(p.NB, MAJOR_B ? TM : TN);
  const long long tasks = (long long)ceil_div(p.NS, S) * p.tilesA * p.tilesB;
  const int blocks = ceil_div(tasks, kWarps);
  const long long target_warps = 148LL * 32;
  int chunks = 1;
  if (tasks < target_warps) chunks = (int)std::min<long long>((target_warps + tasks - 1) / tasks, std::max(1, p.n_list / 4));
  p.views_per_chunk = ceil_div(p.n_list, chunks);
  chunks = ceil_div(p.n_list, p.views_per_chunk);
  return dim3(blocks, chunks);
}

// joint-column walk forward: one launch per (major axis, minor sign, major sign) class of safe views
template <bool MAJOR_B, bool MINOR_UP, bool MAJ_POS>
int launch_walk_forward_joint_class(const xct_plan* pl, const float* in, float* out, cudaStream_t st, int s_begin,
                                    int s_count) {
  const int cls = (MAJOR_B ? 4 : 0) + (MINOR_UP ? 2 : 0) + (MAJ_POS ? 1 : 0);
  if (pl->n_listJ[cls] == 0) return XCT_OK;
  xct::Walk2Params wp{};
  wp.p = plane_params(pl, 1);
  wp.rowoff = pl->d_rowoff;
  wp.out_scale = pl->rows_unit ? 2.0f : 1.0f;  // unit rows: the row weight 2 is folded into the voxels
  wp.row_stride = wp.p.NS;
  wp.s_base = s_begin;
  xct::PlaneParams& p = wp.p;
  if (s_count >= 0) p.NS = s_count;
  in += (size_t)s_begin * p.NA * p.NB;
  p.view_list = pl->d_listJ[cls];
  p.n_list = pl->n_listJ[cls];
  const dim3 grid = walk_forward_grid<kWFwdS, kWFwdTN, 2, MAJOR_B>(p);
  const size_t smem = (size_t)kWarps * kWFwdS * kFwdWin * sizeof(float);
  if (!pl->rows_unit)
    xct::walk_forward_joint_kernel<xct::Geom3, kWFwdS, kWFwdTN, kFwdWin, MAJOR_B, MINOR_UP, MAJ_POS, xct::ROWS_MIX, kWarps>
        <<<grid, kWarps * 32, smem, st>>>(wp, in, out);
  else if (pl->rows_krow)
    xct::walk_forward_joint_kernel<xct::Geom3, kWFwdS, kWFwdTN, kFwdWin, MAJOR_B, MINOR_UP, MAJ_POS, xct::ROWS_KROW, kWarps>
        <<<grid, kWarps * 32, smem, st>>>(wp, in, out);
  else
    xct::walk_forward_joint_kernel<xct::Geom3, kWFwdS, kWFwdTN, kFwdWin, MAJOR_B, MINOR_UP, MAJ_POS, xct::ROWS_TABLE, kWarps>
        <<<grid, kWarps * 32, smem, st>>>(wp, in, out);
  return launch_ok("walk_forward_joint_kernel");
}

// joint forward on a CTA-shared tile: one launch per (major axis, minor sign, major sign) class
template <bool MAJOR_B, bool MINOR_UP, bool MAJ_POS>
int launch_walk_forward_tile_class(const xct_plan* pl, const float* in, float* out, cudaStream_t st, int s_begin, int s_count) {
  const int cls = (MAJOR_B ? 4 : 0) + (MINOR_UP ? 2 : 0) + (MAJ_POS ? 1 : 0);
  if (pl->n_listJ[cls] == 0) return XCT_OK;
  xct::Walk2Params wp{};
  wp.p = plane_params(pl, 1);
  wp.rowoff = pl->d_rowoff;
  wp.out_scale = 2.0f;
  wp.row_stride = wp.p.NS;
  wp.s_base = s_begin;
  xct::PlaneParams& p = wp.p;
  if (s_count >= 0) p.NS = s_count;
  in += (size_t)s_begin * p.NA * p.NB;
  p.view_list = pl->d_listJ[cls];
  p.n_list = pl->n_listJ[cls];
  p.tilesA = ceil_div(p.NA, MAJOR_B ? kWTileTN : 64);
  p.tilesB = ceil_div(p.NB, MAJOR_B ? 64 : kWTileTN);
  const long long tiles = (long long)ceil_div(p.NS, kWTileS) * p.tilesA * p.tilesB;
  if (tiles > 0x7fffffffLL) return fail(XCT_ERR_INVALID, "volume too large for the tile forward grid");
  // small problems: split the view list over blockIdx.y so that the grid fills the SMs (8 views run per CTA at a time)
  int chunks = 1;
  if (tiles < 148LL * 3) chunks = (int)std::min<long long>((148LL * 3 + tiles - 1) / tiles, std::max(1, p.n_list / (2 * kWTileWarps * kWTileNVW)));
  p.views_per_chunk = ceil_div(p.n_list, chunks);
  chunks = ceil_div(p.n_list, p.views_per_chunk);
  const size_t smem = ((size_t)kWTileTN * 2 * 32 + (size_t)kWTileWarps * kWTileNVW * kWTileWin) * (kWTileS / 4) * sizeof(float4);
  const dim3 grid((unsigned)tiles, chunks);
  if (pl->rows_krow) {
    auto kern = xct::walk_forward_tile_kernel<xct::Geom3, kWTileS, kWTileTN, kWTileWin, MAJOR_B, MINOR_UP, MAJ_POS, xct::ROWS_KROW, kWTileWarps, kWTileMinB, kWTileNVW>;
    XCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kWTileWarps * 32, smem, st>>>(wp, in, out);
  } else {
    auto kern = xct::walk_forward_tile_kernel<xct::Geom3, kWTileS, kWTileTN, kWTileWin, MAJOR_B, MINOR_UP, MAJ_POS, xct::ROWS_TABLE, kWTileWarps, kWTileMinB, kWTileNVW>;
    XCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kWTileWarps * 32, smem, st>>>(wp, in, out);
  }
  return launch_ok("walk_forward_tile_kernel");
}

int launch_walk_forward_tile(const xct_plan* pl, const float* in, float* out, cudaStream_t st, int s_begin, int s_count) {
  int rc;
  if ((rc = launch_walk_forward_tile_class<true, true, true>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_tile_class<true, true, false>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_tile_class<true, false, true>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_tile_class<true, false, false>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_tile_class<false, true, true>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_tile_class<false, true, false>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_tile_class<false, false, true>(pl, in, out, st, s_begin, s_count))) return rc;
  return launch_walk_forward_tile_class<false, false, false>(pl, in, out, st, s_begin, s_count);
}

template <class G, bool IS3D, int S, int TN, int GS, bool MAJOR_B, bool MINOR_UP, bool COLD, bool UNIT4>
int launch_walk_forward_class(const xct_plan* pl, int batch, const float* in, float* out, cudaStream_t st,
                              int s_begin, int s_count, bool risky_only = false) {
  const int cls = (MAJOR_B ? 2 : 0) + (MINOR_UP ? 1 : 0);
  const int* list = risky_only ? pl->d_listR[cls] : pl->d_list4[cls];
  const int n_list = risky_only ? pl->n_listR[cls] : pl->n_list4[cls];
  if (n_list == 0) return XCT_OK;
  xct::Walk2Params wp{};
  wp.p = plane_params(pl, batch);
  wp.rowoff = pl->d_rowoff;
  wp.out_scale = 2.0f;
  wp.row_stride = wp.p.NS;
  wp.s_base = s_begin;
  xct::PlaneParams& p = wp.p;
  if (s_count >= 0) p.NS = s_count;
  in += (size_t)s_begin * p.NA * p.NB;
  p.view_list = list;
  p.n_list = n_list;
  const dim3 grid = walk_forward_grid<S, TN, GS, MAJOR_B>(p);
  const size_t smem = (size_t)kWarps * S * kFwdWin * sizeof(float);
  xct::walk_forward_kernel<G, IS3D, S, TN, GS, kFwdWin, MAJOR_B, MINOR_UP, COLD, UNIT4, kWarps>
      <<<grid, kWarps * 32, smem, st>>>(wp, in, out);
  return launch_ok("walk_forward_kernel");
}

template <class G, bool IS3D, int S, int TN, bool COLD, bool UNIT4>
int launch_walk_forward_v(const xct_plan* pl, int batch, const float* in, float* out, cudaStream_t st, int s_begin,
                          int s_count, bool risky_only = false) {
  int rc;
  if ((rc = launch_walk_forward_class<G, IS3D, S, TN, 2, true, true, COLD, UNIT4>(pl, batch, in, out, st, s_begin, s_count, risky_only))) return rc;
  if ((rc = launch_walk_forward_class<G, IS3D, S, TN, 2, true, false, COLD, UNIT4>(pl, batch, in, out, st, s_begin, s_count, risky_only))) return rc;
  if ((rc = launch_walk_forward_class<G, IS3D, S, TN, 2, false, true, COLD, UNIT4>(pl, batch, in, out, st, s_begin, s_count, risky_only))) return rc;
  return launch_walk_forward_class<G, IS3D, S, TN, 2, false, false, COLD, UNIT4>(pl, batch, in, out, st, s_begin, s_count, risky_only);
}

// 2D joint-pair forward: one launch per (major axis, minor sign, major sign) class
template <bool MAJOR_B, bool MINOR_UP, bool MAJ_POS>
int launch_walk2d_forward_class(const xct_plan* pl, int batch, const float* in, float* out, cudaStream_t st) {
  const int cls = (MAJOR_B ? 4 : 0) + (MINOR_UP ? 2 : 0) + (MAJ_POS ? 1 : 0);
  if (pl->n_listJ[cls] == 0) return XCT_OK;
  xct::PlaneParams p = plane_params(pl, batch);
  p.view_list = pl->d_listJ[cls];
  p.n_list = pl->n_listJ[cls];
  p.tilesA = ceil_div(p.NA, MAJOR_B ? kW2dTN : 128);
  p.tilesB = ceil_div(p.NB, MAJOR_B ? 128 : kW2dTN);
  const long long tasks = (long long)p.NS * p.tilesA * p.tilesB;
  const int blocks = ceil_div(tasks, kWarps);
  const long long target_warps = 148LL * 32;
  int chunks = 1;
  if (tasks < target_warps) chunks = (int)std::min<long long>((target_warps + tasks - 1) / tasks, std::max(1, p.n_list / 4));
  p.views_per_chunk = ceil_div(p.n_list, chunks);
  chunks = ceil_div(p.n_list, p.views_per_chunk);
  const size_t smem = (size_t)kWarps * kW2dWin * sizeof(float);
  xct::walk2d_forward_joint_kernel<xct::Geom2, kW2dTN, kW2dWin, MAJOR_B, MINOR_UP, MAJ_POS, kWarps>
      <<<dim3(blocks, chunks), kWarps * 32, smem, st>>>(p, in, out);
  return launch_ok("walk2d_forward_joint_kernel");
}

// Parameters of one class of the 2D joint forward (shared by the per-class and the one-launch path)
template <bool MAJOR_B>
int walk2d_class_params(const xct_plan* pl, int batch, int cls, xct::PlaneParams& p, int& blocks, int& chunks) {
  p = plane_params(pl, batch);
  p.view_list = pl->d_listJ[cls];
  p.n_list = pl->n_listJ[cls];
  p.tilesA = ceil_div(p.NA, MAJOR_B ? kW2dTN : 128);
  p.tilesB = ceil_div(p.NB, MAJOR_B ? 128 : kW2dTN);
  const long long tasks = (long long)p.NS * p.tilesA * p.tilesB;
  blocks = ceil_div(tasks, kWarps);
  const long long target_warps = 148LL * 32;
  chunks = 1;
  if (tasks < target_warps) chunks = (int)std::min<long long>((target_warps + tasks - 1) / tasks, std::max(1, p.n_list / 4));
  p.views_per_chunk = ceil_div(p.n_list, chunks);
  chunks = ceil_div(p.n_list, p.views_per_chunk);
  return XCT_OK;
}

// small problems: all view classes in ONE launch (walk2d_forward_joint_all_kernel)
int launch_walk2d_forward_all(const xct_plan* pl, int batch, const float* in, float* out, cudaStream_t st) {
  xct::Walk2dAllParams ap{};
  // view chunks: the classes fill the GPU TOGETHER, so the warps of all classes are sized to a whole number of
  // waves of resident warps (148 SMs x 24) instead of each class to its own wave
  long long total_tasks = 0;
  for (int cls = 0; cls < 8; ++cls) {
    if (pl->n_listJ[cls] == 0) continue;
    int blocks = 0, chunks = 0;
    if (cls & 4) walk2d_class_params<true>(pl, batch, cls, ap.p[cls], blocks, chunks);
    else walk2d_class_params<false>(pl, batch, cls, ap.p[cls], blocks, chunks);
    total_tasks += (long long)blocks * kWarps;
  }
  double waves = 1.5;  // measured at C2 (tools/bench_c2.py): 0.5 / 1 / 2 / 3 / 4 / 6 waves -> 101 / 97 / 99 / 99 / 115 / 110 us
  if (const char* env = std::getenv("XCT_2D_ALL_WAVES")) waves = std::max(0.25, std::atof(env));  // tuning (tools/bench_c2.py)
  const int want_chunks = total_tasks > 0 ? (int)std::max<long long>(1, (long long)(waves * 148 * 24 + total_tasks / 2) / total_tasks) : 1;
  int total = 0;
  for (int cls = 0; cls < 8; ++cls) {
    ap.block_begin[cls] = total;
    ap.blocks_x[cls] = 1;
    if (pl->n_listJ[cls] == 0) continue;
    xct::PlaneParams& p = ap.p[cls];
    const long long tasks = (long long)p.NS * p.tilesA * p.tilesB;
    const int blocks = ceil_div(tasks, kWarps);
    int chunks = std::min(want_chunks, std::max(1, p.n_list / 4));
    p.views_per_chunk = ceil_div(p.n_list, chunks);
    chunks = ceil_div(p.n_list, p.views_per_chunk);
    ap.blocks_x[cls] = blocks;
    total += blocks * chunks;
  }
  ap.block_begin[8] = total;
  if (total == 0) return XCT_OK;
  const size_t smem = (size_t)kWarps * kW2dWin * sizeof(float);
  xct::walk2d_forward_joint_all_kernel<xct::Geom2, kW2dTN, kW2dWin, kWarps><<<total, kWarps * 32, smem, st>>>(ap, in, out);
  return launch_ok("walk2d_forward_joint_all_kernel");
}

// 2D joint forward on a CTA-shared tile: one launch per (major axis, minor sign, major sign) class
template <bool MAJOR_B, bool MINOR_UP, bool MAJ_POS>
int launch_walk2d_forward_tile_class(const xct_plan* pl, int batch, const float* in, float* out, cudaStream_t st) {
  const int cls = (MAJOR_B ? 4 : 0) + (MINOR_UP ? 2 : 0) + (MAJ_POS ? 1 : 0);
  if (pl->n_listJ[cls] == 0) return XCT_OK;
  xct::PlaneParams p = plane_params(pl, batch);
  p.view_list = pl->d_listJ[cls];
  p.n_list = pl->n_listJ[cls];
  p.tilesA = ceil_div(p.NA, MAJOR_B ? kW2dTileTN : 128);
  p.tilesB = ceil_div(p.NB, MAJOR_B ? 128 : kW2dTileTN);
  const long long tiles = (long long)p.NS * p.tilesA * p.tilesB;
  if (tiles > 0x7fffffffLL) return fail(XCT_ERR_INVALID, "image batch too large for the 2D tile forward grid");
  // split the view list over blockIdx.y until the grid is ~8 waves of resident CTAs (5 per SM): a CTA that runs all
  // views of its class is long, and a last wave that is 3/4 full costs its whole duration
  long long target = 148LL * 40;
  if (const char* env = std::getenv("XCT_2D_TILE_TARGET_CTAS")) target = std::max(1, std::atoi(env));  // tuning (tools/bench_2d.py)
  int chunks = 1;
  if (tiles < target) chunks = (int)std::min<long long>((target + tiles - 1) / tiles, std::max(1, p.n_list / (2 * kW2dTileWarps)));
  p.views_per_chunk = ceil_div(p.n_list, chunks);
  chunks = ceil_div(p.n_list, p.views_per_chunk);
  const size_t smem = ((size_t)kW2dTileTN * 32 * 4 + (size_t)kW2dTileWarps * kW2dTileWin) * sizeof(float);
  auto kern = xct::walk2d_forward_tile_kernel<xct::Geom2, kW2dTileTN, kW2dTileWin, MAJOR_B, MINOR_UP, MAJ_POS, kW2dTileWarps>;
  if (smem > 48 * 1024) XCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((unsigned)tiles, chunks), kW2dTileWarps * 32, smem, st>>>(p, in, out);
  return launch_ok("walk2d_forward_tile_kernel");
}

int launch_walk2d_forward(const xct_plan* pl, int batch, const float* in, float* out, cudaStream_t st) {
  // below ~2 waves of CTAs per class the per-launch ramp-up and tail dominate: one launch for all classes
  if ((long long)batch * pl->n0 * pl->n1 * pl->V <= (1LL << 29) && !pl->fwd_2d_per_class)
    return launch_walk2d_forward_all(pl, batch, in, out, st);
  int rc;
  if (pl->fwd_tile2d) {  // large problems: the CTA-shared tile (64-step walks)
    if ((rc = launch_walk2d_forward_tile_class<true, true, true>(pl, batch, in, out, st))) return rc;
    if ((rc = launch_walk2d_forward_tile_class<true, true, false>(pl, batch, in, out, st))) return rc;
    if ((rc = launch_walk2d_forward_tile_class<true, false, true>(pl, batch, in, out, st))) return rc;
    if ((rc = launch_walk2d_forward_tile_class<true, false, false>(pl, batch, in, out, st))) return rc;
    if ((rc = launch_walk2d_forward_tile_class<false, true, true>(pl, batch, in, out, st))) return rc;
    if ((rc = launch_walk2d_forward_tile_class<false, true, false>(pl, batch, in, out, st))) return rc;
    if ((rc = launch_walk2d_forward_tile_class<false, false, true>(pl, batch, in, out, st))) return rc;
    return launch_walk2d_forward_tile_class<false, false, false>(pl, batch, in, out, st);
  }
  if ((rc = launch_walk2d_forward_class<true, true, true>(pl, batch, in, out, st))) return rc;
  if ((rc = launch_walk2d_forward_class<true, true, false>(pl, batch, in, out, st))) return rc;
  if ((rc = launch_walk2d_forward_class<true, false, true>(pl, batch, in, out, st))) return rc;
  if ((rc = launch_walk2d_forward_class<true, false, false>(pl, batch, in, out, st))) return rc;
  if ((rc = launch_walk2d_forward_class<false, true, true>(pl, batch, in, out, st))) return rc;
  if ((rc = launch_walk2d_forward_class<false, true, false>(pl, batch, in, out, st))) return rc;
  if ((rc = launch_walk2d_forward_class<false, false, true>(pl, batch, in, out, st))) return rc;
  return launch_walk2d_forward_class<false, false, false>(pl, batch, in, out, st);
}

int launch_walk_forward_joint(const xct_plan* pl, const float* in, float* out, cudaStream_t st, int s_begin, int s_count) {
  int rc;
  if ((rc = launch_walk_forward_joint_class<true, true, true>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_joint_class<true, true, false>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_joint_class<true, false, true>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_joint_class<true, false, false>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_joint_class<false, true, true>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_joint_class<false, true, false>(pl, in, out, st, s_begin, s_count))) return rc;
  if ((rc = launch_walk_forward_joint_class<false, false, true>(pl, in, out, st, s_begin, s_count))) return rc;
  return launch_walk_forward_joint_class<false, false, false>(pl, in, out, st, s_begin, s_count);
}

// 3D separable forward.  The vector flush needs unit rows, D1 % 4 == 0 and a 16-byte aligned sinogram.
// Slices [s_begin, s_begin + s_count) of the plan only (s_count < 0: all); `in` is the full volume.
int launch_walk_forward3(const xct_plan* pl, const float* in, float* out, cudaStream_t st, int s_begin = 0,
                         int s_count = -1) {
  const bool aligned16 = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const bool unit4 = pl->fwd_unit4 && aligned16;
  if (pl->fwd_cold) {
    return launch_walk_forward_v<xct::Geom3, true, kWFwdS, kWFwdTN, true, false>(pl, 1, in, out, st, s_begin, s_count);
  }
  if (pl->fwd_tile && aligned16) return launch_walk_forward_tile(pl, in, out, st, s_begin, s_count);
  if (pl->fwd_joint && aligned16 && (pl->fwd_unit4 || pl->fwd_mix4)) {
    // joint-column kernel; views whose minor coefficient can reach one bin per voxel: 2-bin walk
    int rc = launch_walk_forward_joint(pl, in, out, st, s_begin, s_count);
    if (rc) return rc;
    if (pl->fwd_unit4)
      return launch_walk_forward_v<xct::Geom3, true, kWFwdS, kWFwdTN, false, true>(pl, 1, in, out, st, s_begin, s_count, true);
    return launch_walk_forward_v<xct::Geom3, true, kWFwdS, kWFwdTN, false, false>(pl, 1, in, out, st, s_begin, s_count, true);
  }
  if (unit4) return launch_walk_forward_v<xct::Geom3, true, kWFwdS, kWFwdTN, false, true>(pl, 1, in, out, st, s_begin, s_count);
  return launch_walk_forward_v<xct::Geom3, true, kWFwdS, kWFwdTN, false, false>(pl, 1, in, out, st, s_begin, s_count);
}

size_t in_elems(const xct_plan* pl) { return (size_t)pl->n0 * pl->n1 * pl->n2; }
size_t out_elems(const xct_plan* pl) { return (size_t)pl->V * pl->d0 * pl->d1; }

// ------------------------------------------------------------------ brick kernels (general 3D)
// Host analysis of one plan: window extents of every view, and per view the depth axis of the forward
// lattice and whether the 2-voxel lane lattice keeps the lanes of one instruction on distinct bins.
struct BrickAnalysis {
  bool adj_ok = true, fwd_ok = true;
  std::vector<int> list[6];  // [2 * depth + needs_atomics]
};
BrickAnalysis analyse_bricks(const float* mats, int V, int n0, int n1, int n2, int slice_offset) {
  BrickAnalysis out;
  using BG = xct::BrickFwdGeom;
  const float idx_max = (float)std::max(std::max(n0 + std::abs(slice_offset), n1), n2) + 1.f;
  for (int v = 0; v < V; ++v) {
    const float* M = mats + 8 * (size_t)v;
    // coordinate magnitude -> rounding slack (each coordinate is one product + four sums: <= 4 ulp)
    float umax = 0.f;
    for (int r = 0; r < 2; ++r)
      umax = std::max(umax, (std::fabs(M[4 * r]) + std::fabs(M[4 * r + 1]) + std::fabs(M[4 * r + 2])) * idx_max + std::fabs(M[4 * r + 3]) + 2.f);
    if (!(umax < 4.0e6f)) { out.adj_ok = out.fwd_ok = false; continue; }  // int / fp32 index headroom
    const float slack = 16.f * std::ldexp(1.f, std::ilogb(umax) - 23) + 0.01f;
    // adjoint: 8^3 brick
    for (int r = 0; r < 2; ++r) {
      const float e = (std::fabs(M[4 * r]) + std::fabs(M[4 * r + 1]) + std::fabs(M[4 * r + 2])) * (float)(xct::kBrick - 1);
      if (!(e + slack < (float)(r == 0 ? kBrAdjWR - 2 : kBrAdjWC - 5))) out.adj_ok = false;  // columns: 4-aligned start
    }
    // forward: depth axis = smallest projected length first; the lane lattice must separate all lane pairs
    float len[3];
    for (int q = 0; q < 3; ++q) len[q] = std::hypot(M[q], M[4 + q]);
    int order[3] = {0, 1, 2};
    std::sort(order, order + 3, [&](int a, int b) { return len[a] < len[b]; });
    int pick = -1, pick_atomic = -1;
    for (int t = 0; t < 3 && pick < 0; ++t) {
      const int depth = order[t];
      const int ax_a = depth == 0 ? 1 : 0, ax_b = depth == 2 ? 1 : 2;
      const float er = std::fabs(M[ax_a]) * (BG::EA - 1) + std::fabs(M[ax_b]) * (BG::EB - 1) + std::fabs(M[depth]) * (BG::EC - 1);
      const float ec = std::fabs(M[4 + ax_a]) * (BG::EA - 1) + std::fabs(M[4 + ax_b]) * (BG::EB - 1) + std::fabs(M[4 + depth]) * (BG::EC - 1);
      if (!(er + slack < (float)(kBrFwdWR - 2)) || !(ec + slack + 3.f < (float)(kBrFwdWC - 2))) continue;  // +3: 4-bin aligned start
      if (pick_atomic < 0) pick_atomic = depth;
      bool distinct = true;
      for (int na = -(BG::LA - 1); na < BG::LA && distinct; ++na)
        for (int nb = -(BG::LB - 1); nb < BG::LB; ++nb) {
          if (na == 0 && nb == 0) continue;
          const float d0 = 2.f * (na * M[ax_a] + nb * M[ax_b]), d1 = 2.f * (na * M[4 + ax_a] + nb * M[4 + ax_b]);
          if (!(std::max(std::fabs(d0), std::fabs(d1)) >= 1.f + slack)) { distinct = false; break; }
        }
      if (distinct) pick = depth;
    }
    if (pick >= 0) out.list[2 * pick].push_back(v);
    else if (pick_atomic >= 0) out.list[2 * pick_atomic + 1].push_back(v);
    else out.fwd_ok = false;
  }
  return out;
}

xct::BrickParams brick_params(const xct_plan* pl) {
  xct::BrickParams b{};
  b.mats = pl->d_mats_t;
  b.view_list = nullptr;
  b.n_list = pl->V;
  b.V = pl->V;
  b.N0 = pl->n0; b.N1 = pl->n1; b.N2 = pl->n2; b.D0 = pl->d0; b.D1 = pl->d1;
  b.slice_offset = pl->slice_offset; b.row_off = pl->row_off;
  b.views_per_chunk = pl->V;
  return b;
}

int view_chunks(long long tasks, int n_views, int min_views, int& views_per_chunk) {
  const long long target_warps = 148LL * 32;
  int chunks = 1;
  if (tasks < target_warps) chunks = (int)std::min<long long>((target_warps + tasks - 1) / tasks, std::max(1, n_views / min_views));
  views_per_chunk = ceil_div(n_views, chunks);
  return ceil_div(n_views, views_per_chunk);
}

int launch_brick_adjoint(const xct_plan* pl, const float* in, float* out, cudaStream_t st, const xct::OutRoute* route = nullptr) {
  xct::BrickParams b = brick_params(pl);
  b.nb0 = ceil_div(pl->n0, xct::kBrick); b.nb1 = ceil_div(pl->n1, xct::kBrick); b.nb2 = ceil_div(pl->n2, xct::kBrick);
  const long long tasks = (long long)b.nb0 * b.nb1 * b.nb2;
  int chunks = view_chunks(tasks, b.n_list, 8, b.views_per_chunk);
  if (route && route->store) { chunks = 1; b.views_per_chunk = b.n_list; }  // plain stores: each element exactly once
  const int blocks = ceil_div(tasks, kWarps);
  constexpr int stage_floats = ((kBrAdjWR * kBrAdjWC + 31) / 32) * 32;
  const size_t smem = (size_t)kWarps * kBrAdjStages * stage_floats * sizeof(float) + (size_t)kWarps * kBrAdjStages * sizeof(unsigned long long);
  if (!route && chunks > 1) XCT_CUDA(cudaMemsetAsync(out, 0, in_elems(pl) * sizeof(float), st));
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  bool tma = pl->brick_adj_tma && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  if (tma) {
    const cuuint64_t dims[3] = {(cuuint64_t)pl->d1, (cuuint64_t)pl->d0, (cuuint64_t)pl->V};
    const cuuint64_t strides[2] = {(cuuint64_t)pl->d1 * sizeof(float), (cuuint64_t)pl->d0 * pl->d1 * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)kBrAdjWC, (cuuint32_t)kBrAdjWR, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), dims, strides, box,
                                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) tma = false;  // e.g. a stride the tensor map cannot express: cp.async staging
  }
  const dim3 grid(blocks, chunks);
  const xct::OutRoute none{};
#define XCT_BRICK_ADJ(TMA_, ROUTE_)                                                                                            \
  do {                                                                                                                         \
    auto kern = xct::brick_adjoint_kernel<kBrAdjWR, kBrAdjWC, kBrAdjStages, kWarps, TMA_, ROUTE_>;                              \
    XCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                              \
    kern<<<grid, kWarps * 32, smem, st>>>(b, in, out, tmap, route ? *route : none);                                            \
  } while (0)
  if (tma) { if (route) XCT_BRICK_ADJ(true, true); else XCT_BRICK_ADJ(true, false); }
  else { if (route) XCT_BRICK_ADJ(false, true); else XCT_BRICK_ADJ(false, false); }
#undef XCT_BRICK_ADJ
  return launch_ok("brick_adjoint_kernel");
}

template <int DEPTH, bool ATOMIC>
int launch_brick_forward_class(const xct_plan* pl, const float* in, float* out, cudaStream_t st) {
  const int cls = 2 * DEPTH + (ATOMIC ? 1 : 0);
  if (pl->n_listB[cls] == 0) return XCT_OK;
  using BG = xct::BrickFwdGeom;
  xct::BrickParams b = brick_params(pl);
  b.view_list = pl->d_listB[cls];
  b.n_list = pl->n_listB[cls];
  const int dims[3] = {pl->n0, pl->n1, pl->n2};
  const int ax_a = DEPTH == 0 ? 1 : 0, ax_b = DEPTH == 2 ? 1 : 2;
  b.nb0 = ceil_div(dims[ax_a], BG::EA); b.nb1 = ceil_div(dims[ax_b], BG::EB); b.nb2 = ceil_div(dims[DEPTH], BG::EC);
  const long long tasks = (long long)b.nb0 * b.nb1 * b.nb2;
  const int chunks = view_chunks(tasks, b.n_list, 4, b.views_per_chunk);
  const dim3 grid(ceil_div(tasks, kWarps), chunks);
  const size_t smem = (size_t)kWarps * kBrFwdWR * kBrFwdWC * sizeof(float);
  const bool vec4 = (pl->d1 % 4 == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec4) xct::brick_forward_kernel<DEPTH, kBrFwdWR, kBrFwdWC, kWarps, ATOMIC, true><<<grid, kWarps * 32, smem, st>>>(b, in, out);
  else xct::brick_forward_kernel<DEPTH, kBrFwdWR, kBrFwdWC, kWarps, ATOMIC, false><<<grid, kWarps * 32, smem, st>>>(b, in, out);
  return launch_ok("brick_forward_kernel");
}

int launch_brick_forward(const xct_plan* pl, const float* in, float* out, cudaStream_t st) {
  int rc;
  if ((rc = launch_brick_forward_class<2, false>(pl, in, out, st))) return rc;
  if ((rc = launch_brick_forward_class<1, false>(pl, in, out, st))) return rc;
  if ((rc = launch_brick_forward_class<0, false>(pl, in, out, st))) return rc;
  if ((rc = launch_brick_forward_class<2, true>(pl, in, out, st))) return rc;
  if ((rc = launch_brick_forward_class<1, true>(pl, in, out, st))) return rc;
  return launch_brick_forward_class<0, true>(pl, in, out, st);
}

xct::Gen3Params gen3_params(const xct_plan* pl) {
  xct::Gen3Params g{};
  g.mats = pl->d_mats;
  g.V = pl->V; g.N0 = pl->n0; g.N1 = pl->n1; g.N2 = pl->n2; g.D0 = pl->d0; g.D1 = pl->d1;
  g.slice_offset = pl->slice_offset; g.row_off = pl->row_off; g.rows_total = pl->rows_total;
  return g;
}
xct::Gen2Params gen2_params(const xct_plan* pl, int batch) {
  xct::Gen2Params g{};
  g.views = pl->d_views;
  g.V = pl->V; g.N0 = pl->n0; g.N1 = pl->n1; g.ny = pl->d1; g.batch = batch;
  return g;
}


int check_call(const xct_plan* pl, const void* a, const void* b, int batch) {
  if (!pl) return fail(XCT_ERR_INVALID, "null plan");
  if (pl->dry) return fail(XCT_ERR_INVALID, "analysis-only plan (xct*_plan_analyse) cannot compute");
  if (!a || !b) return fail(XCT_ERR_INVALID, "null buffer");
  if (batch < 1) return fail(XCT_ERR_INVALID, "batch must be >= 1");
  if (pl->ndim == 3 && batch != 1) return fail(XCT_ERR_INVALID, "3D plans take batch == 1");
  return XCT_OK;
}

int ensure_stage(xct_plan* pl, int dir, size_t n_in, size_t n_out) {
  if (!pl->hstream) XCT_CUDA(cudaStreamCreateWithFlags(&pl->hstream, cudaStreamNonBlocking));
  if (!pl->s_in) XCT_CUDA(cudaStreamCreateWithFlags(&pl->s_in, cudaStreamNonBlocking));
  if (!pl->s_out) XCT_CUDA(cudaStreamCreateWithFlags(&pl->s_out, cudaStreamNonBlocking));
  auto grow = [&](float*& buf, size_t& cap, size_t n) -> cudaError_t {
    if (n <= cap) return cudaSuccess;
    // a larger batch than before: everything queued on the plan's streams may still use the old buffer
    cudaStreamSynchronize(pl->s_in); cudaStreamSynchronize(pl->hstream); cudaStreamSynchronize(pl->s_out);
    if (buf) cudaFree(buf);
    buf = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&buf, n * sizeof(float));
    if (e == cudaSuccess) cap = n;
    return e;
  };
  XCT_CUDA(grow(pl->stage_in[dir], pl->cap_in[dir], n_in));
  XCT_CUDA(grow(pl->stage_out[dir], pl->cap_out[dir], n_out));
  return XCT_OK;
}

}  // namespace

// =====================================================================================
extern "C" {

int xct_version(void) { return XCT_VERSION; }
const char* xct_last_error(void) { return g_err.c_str(); }
int64_t xct_launch_count(void) { return g_launches; }
void xct_launch_count_reset(void) { g_launches = 0; }

int xct_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static int plan2d_create_impl(xct_plan** out, const xct2d_geom* g, bool dry) {
  if (!out || !g) return fail(XCT_ERR_INVALID, "null argument");
  *out = nullptr;
  if (g->n0 < 1 || g->n1 < 1 || g->num_views < 1 || g->det_count < 1 || !g->view_table)
    return fail(XCT_ERR_INVALID, "xct2d_plan_create: bad shape or null view_table");
  if ((long long)g->n0 * g->n1 > (1LL << 40)) return fail(XCT_ERR_INVALID, "image too large");
  int rc = dry ? XCT_OK : check_device(g->device);
  if (rc) return rc;
  DeviceGuard guard(dry ? -1 : g->device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");

  xct_plan* pl = new (std::nothrow) xct_plan();
  if (!pl) return fail(XCT_ERR_INVALID, "out of host memory");
  pl->dry = dry;
  pl->ndim = 2; pl->device = g->device; pl->V = g->num_views;
  pl->n0 = g->n0; pl->n1 = g->n1; pl->n2 = 1; pl->d0 = 1; pl->d1 = g->det_count;

  std::vector<xct::ViewRec> views(g->num_views);
  bool finite = true;
  for (int v = 0; v < g->num_views; ++v) {
    const float* t = g->view_table + 4 * (size_t)v;
    xct::ViewRec r{};
    r.off = t[0]; r.ca = t[1]; r.cb = t[2]; r.width = t[3];
    r.rwidth = 1.0f / t[3];
    {  // same rounding margin as the 3D walk kernels: the bin may advance by two only above 1 - 5 ulp(u)
      const float umax = std::fabs(r.ca) * g->n0 + std::fabs(r.cb) * g->n1 + std::fabs(r.off) + 2.f;
      const float ulp = std::ldexp(1.f, std::ilogb(umax) - 23);
      r.fjump = (std::max(std::fabs(r.ca), std::fabs(r.cb)) + 5.f * ulp > 1.f) ? 1.f : 0.f;
    }
    if (!(std::isfinite(t[0]) && std::isfinite(t[1]) && std::isfinite(t[2]) && t[3] > 0.f)) finite = false;
    views[v] = r;
  }
  if (!finite) {
    delete pl;
    return fail(XCT_ERR_INVALID, "xct2d_plan_create: non-finite view table or width <= 0");
  }
  Envelope env = analyse_views(views, kAdj2TA, kFwd2TN);
  const bool force_general = (g->flags & XCT_FLAG_FORCE_GENERAL) != 0;
  pl->adj_plane = env.adj_ok && !force_general;
  pl->fwd_plane = env.fwd_ok && !force_general;
  pl->gs = pl->fwd_plane ? env.gs : 0;
  pl->path = (pl->adj_plane && pl->fwd_plane) ? XCT_PATH_2D_PLANE : XCT_PATH_2D_GENERAL;
  {  // joint-pair forward: lanes 4 voxels apart must land >= 1 bin apart, the window must hold the tile
    bool ok = pl->fwd_plane && !(g->flags & (XCT_FLAG_NO_WALK | XCT_FLAG_NO_JOINT));
    for (const auto& vr : views) {
      const float a = std::fabs(vr.ca), b = std::fabs(vr.cb);
      const float mj = std::max(a, b), mn = std::min(a, b);
      if (vr.fjump != 0.f || 4.f * mj < 1.01f || mj * 127.f + mn * (kW2dTN - 1) + 4.f > (float)kW2dWin) ok = false;
    }
    pl->fwd_joint2d = ok;
    pl->fwd_2d_per_class = (g->flags & XCT_FLAG_2D_PER_CLASS) != 0;
    bool tile_ok = ok && !(g->flags & XCT_FLAG_NO_TILE);
    for (const auto& vr : views) {
      const float mj = std::max(std::fabs(vr.ca), std::fabs(vr.cb)), mn = std::min(std::fabs(vr.ca), std::fabs(vr.cb));
      if (!(mj * 127.f + mn * (kW2dTileTN - 1) + 4.f <= (float)kW2dTileWin)) tile_ok = false;
    }
    pl->fwd_tile2d = tile_ok;
  }

  auto cleanup = [&](int code) { xct_plan_destroy(pl); return code; };
  cudaError_t e = dev_upload(pl, pl->d_views, views.data(), views.size());
  if (e != cudaSuccess) return cleanup(fail(XCT_ERR_CUDA, std::string("view table upload: ") + cudaGetErrorString(e)));
  if ((rc = upload_lists(pl, env))) return cleanup(rc);
  *out = pl;
  return XCT_OK;
}

static int plan3d_create_impl(xct_plan** out, const xct3d_geom* g, bool dry) {
  if (!out || !g) return fail(XCT_ERR_INVALID, "null argument");
  *out = nullptr;
  if (g->n0 < 1 || g->n1 < 1 || g->n2 < 1 || g->d0 < 1 || g->d1 < 1 || g->num_views < 1 || !g->matrices)
    return fail(XCT_ERR_INVALID, "xct3d_plan_create: bad shape or null matrices");
  if ((long long)g->d0 * g->d1 >= (1LL << 31))
    return fail(XCT_ERR_INVALID, "detector too large for int32 row*col offsets");
  int rc = dry ? XCT_OK : check_device(g->device);
  if (rc) return rc;
  DeviceGuard guard(dry ? -1 : g->device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");

  xct_plan* pl = new (std::nothrow) xct_plan();
  if (!pl) return fail(XCT_ERR_INVALID, "out of host memory");
  pl->dry = dry;
  pl->ndim = 3; pl->device = g->device; pl->V = g->num_views;
  pl->n0 = g->n0; pl->n1 = g->n1; pl->n2 = g->n2; pl->d0 = g->d0; pl->d1 = g->d1;
  pl->slice_offset = g->slice_offset; pl->row_off = g->det_row_offset;
  pl->rows_total = g->det_rows_total > 0 ? g->det_rows_total : g->d0;
  auto cleanup = [&](int code) { xct_plan_destroy(pl); return code; };

  const int V = g->num_views;
  bool sep = true, finite = true;
  for (int v = 0; v < V; ++v) {
    const float* M = g->matrices + 8 * (size_t)v;
    for (int q = 0; q < 8; ++q) finite = finite && std::isfinite(M[q]);
    // rows depend on voxel axis 0 only, columns on axes 1, 2 only
    if (!(M[1] == 0.f && M[2] == 0.f && M[4] == 0.f)) sep = false;
  }
  if (!finite) return cleanup(fail(XCT_ERR_INVALID, "xct3d_plan_create: non-finite matrix entry"));

  cudaError_t e = dev_upload(pl, pl->d_mats, g->matrices, 8 * (size_t)V);
  if (e != cudaSuccess) return cleanup(fail(XCT_ERR_CUDA, std::string("matrix upload: ") + cudaGetErrorString(e)));

  pl->path = XCT_PATH_3D_GENERAL;
  if (sep && !(g->flags & XCT_FLAG_FORCE_GENERAL)) {
    std::vector<xct::ViewRec> views(V);
    for (int v = 0; v < V; ++v) {
      const float* M = g->matrices + 8 * (size_t)v;
      xct::ViewRec r{};
      r.ca = M[5]; r.cb = M[6]; r.off = M[7]; r.width = 1.f; r.rwidth = 1.f;
      // walk kernels: neighbouring rows / columns move the coordinate by the coefficient plus at most
      // 4 ulp of |u| of rounding (one product and three sums on each side, half an ulp each); only
      // when that can exceed 1 can the bin advance by two in one step (margin: 5 ulp)
      const float umax = std::fabs(r.ca) * g->n1 + std::fabs(r.cb) * g->n2 + std::fabs(r.off) + 2.f;
      const float ulp = std::ldexp(1.f, std::ilogb(umax) - 23);
      r.jump = (std::fabs(r.ca) + 5.f * ulp > 1.f) ? 1.f : 0.f;
      {
        const float mj = std::max(std::fabs(r.ca), std::fabs(r.cb)), mn = std::min(std::fabs(r.ca), std::fabs(r.cb));
        r.fjump = mj + 5.f * ulp <= 1.f ? 0.f : ((mn + 5.f * ulp <= 1.f && mj <= 1.5f) ? 1.f : 2.f);
      }
      views[v] = r;
    }
    Envelope env = analyse_views(views, kAdj3TA, kFwd3TN);
    if (env.gs != 2) env.fwd_ok = false;  // only the stride-2 forward is instantiated for 3D
    if (env.adj_ok || env.fwd_ok) {
      // Row records: voxel slice i of view v -> detector row(s) and axis-0 weights.
      // Same expression tree as _xray3d.py:216 with the two zero coefficients kept, so that the
      // result (including signed zeros) is what the general formula gives.  This translation unit
      // is compiled with -ffp-contract=off for host code: every product and sum rounds to fp32.
      std::vector<xct::RowRec> rows((size_t)V * g->n0);
      std::vector<long long> rowoff((size_t)V * g->n0);
      bool aligned = true, unit = true;
      for (int v = 0; v < V; ++v) {
        const float* M = g->matrices + 8 * (size_t)v;
        for (int i = 0; i < g->n0; ++i) {
          volatile float xi = ((float)i + 0.5f) + (float)g->slice_offset;
          volatile float t0 = M[0] * xi;
          volatile float t1 = M[1] * 0.5f;
          volatile float t2 = M[2] * 0.5f;
          volatile float s = t0 + t1;
          s = s + t2;
          s = s + M[3];
          volatile float left = s - 0.25f;
          volatile float d = std::ceil(left) - left;
          const float tn = std::fmin(d, 0.5f);
          const long long rg = (long long)std::floor(left);
          xct::RowRec rr{};
          long long rl = rg - g->det_row_offset;
          rl = std::max<long long>(-(1LL << 30), std::min<long long>(rl, 1LL << 30));  // int range only
          rr.r0 = (int)rl;
          volatile float un = 0.5f - tn;
          rr.w0 = tn * 4.0f;
          rr.w1 = un * 4.0f;
          const bool in0 = rg >= 0 && rg < pl->rows_total && rl >= 0 && rl < g->d0;
          const bool in1 = rg + 1 >= 0 && rg + 1 < pl->rows_total && rl + 1 >= 0 && rl + 1 < g->d0;
          if (!in0) rr.w0 = 0.f;
          if (!in1) rr.w1 = 0.f;
          if (rr.w0 != 0.f && rr.w1 != 0.f) aligned = false;
          int ri = -1;
          if (rr.w0 == 2.f && rr.w1 == 0.f) ri = rr.r0;
          else if (rr.w0 == 0.f && rr.w1 == 2.f) ri = rr.r0 + 1;
          else if (!(rr.w0 == 0.f && rr.w1 == 0.f)) unit = false;
          rowoff[(size_t)v * g->n0 + i] = ri < 0 ? -1LL : ((long long)v * g->d0 + ri) * (long long)g->d1;
          rows[(size_t)v * g->n0 + i] = rr;
        }
      }
      pl->row_aligned = aligned;
      pl->rows_unit = unit;
      // TMA staging of the walk adjoint: per view, local row = local slice + krow wherever a row
      // exists, and slices without a row fall outside [0, d0) under the same rule (zero fill)
      bool tma_ok = unit;
      for (int v = 0; v < V && tma_ok; ++v) {
        bool have = false;
        long long k = 0;
        for (int i = 0; i < g->n0; ++i) {
          const long long off = rowoff[(size_t)v * g->n0 + i];
          if (off < 0) continue;
          const long long r = off / g->d1 - (long long)v * g->d0;
          if (!have) { k = r - i; have = true; }
          else if (r - i != k) tma_ok = false;
        }
        if (!have) k = -(long long)(g->n0 + g->d0 + 64);
        for (int i = 0; i < g->n0 && tma_ok; ++i)
          if (rowoff[(size_t)v * g->n0 + i] < 0 && i + k >= 0 && i + k < g->d0) tma_ok = false;
        if (k < -(1LL << 30) || k > (1LL << 30)) tma_ok = false;
        views[v].krow = (int)k;
      }
      if (!tma_ok)
        for (auto& vr : views) vr.krow = 0;
      pl->rows_krow = tma_ok;
      pl->adj_tma = tma_ok && (dry || tensor_map_encoder() != nullptr) && !(g->flags & XCT_FLAG_NO_TMA);
      if (unit) {
        // per-slice detector row range over the views, and whether rows never decrease with the slice
        // index (any rotation about axis 0 with a positive axis-0 scale): the host pipeline relies on it
        pl->h_row_lo.assign(g->n0, -1);
        pl->h_row_hi.assign(g->n0, -1);
        bool mono = true;
        for (int v = 0; v < V; ++v) {
          long long prev = -1;
          for (int i = 0; i < g->n0; ++i) {
            const long long off = rowoff[(size_t)v * g->n0 + i];
            if (off < 0) continue;
            const int r = (int)(off / g->d1 - (long long)v * g->d0);
            if (r < prev) mono = false;
            prev = r;
            pl->h_row_lo[i] = pl->h_row_lo[i] < 0 ? r : std::min(pl->h_row_lo[i], r);
            pl->h_row_hi[i] = std::max(pl->h_row_hi[i], r);
          }
        }
        pl->pipe_ok = mono;
      }
      if (unit) {
        e = dev_upload(pl, pl->d_rowoff, rowoff.data(), rowoff.size());
        if (e != cudaSuccess) return cleanup(fail(XCT_ERR_CUDA, std::string("row index upload: ") + cudaGetErrorString(e)));
      }
      e = dev_upload(pl, pl->d_views, views.data(), views.size());
      if (e == cudaSuccess) e = dev_upload(pl, pl->d_rows, rows.data(), rows.size());
      for (const auto& vr : views) pl->adj_jump_views += vr.jump != 0.f ? 1 : 0;
      if (e != cudaSuccess) return cleanup(fail(XCT_ERR_CUDA, std::string("table upload: ") + cudaGetErrorString(e)));
      if ((rc = upload_lists(pl, env))) return cleanup(rc);
      pl->adj_plane = env.adj_ok;
      pl->fwd_plane = env.fwd_ok;
      pl->fwd_walk = env.fwd_ok && env.gs == 2 && !(g->flags & XCT_FLAG_NO_WALK);
      pl->fwd_cold = env.max_minor > 0.98f;
      pl->fwd_unit4 = env.fwd_unit4_ok && unit && (g->d1 % 4 == 0);
      pl->fwd_mix4 = env.fwd_unit4_ok && !unit && (g->d1 % 4 == 0);
      pl->fwd_joint = pl->fwd_walk && (pl->fwd_unit4 || pl->fwd_mix4) && !pl->fwd_cold && !(g->flags & XCT_FLAG_NO_JOINT);
      {  // CTA-shared-tile variant of the joint forward: the window must hold 64 major x 32 minor voxels
        bool ok = pl->fwd_joint && pl->fwd_unit4 && !(g->flags & XCT_FLAG_NO_TILE);
        for (const auto& vr : views) {
          const float mj = std::max(std::fabs(vr.ca), std::fabs(vr.cb)), mn = std::min(std::fabs(vr.ca), std::fabs(vr.cb));
          if (!(mj * 63.f + mn * (kWTileTN - 1) + 7.f <= (float)kWTileWin)) ok = false;
        }
        for (int c = 0; c < 4; ++c) ok = ok && pl->n_listR[c] == 0;  // no view on the two-bin walk
        pl->fwd_tile = ok;
      }
      pl->adj_walk = env.adj_ok && env.adj_walk_ok && unit && (g->d1 % 4 == 0) && !(g->flags & XCT_FLAG_NO_WALK);
      pl->adj_tma = pl->adj_tma && pl->adj_walk;  // the TMA box is the walk adjoint's staging
      pl->adj_vec = pl->adj_tma && env.adj_vec_ok && !(g->flags & XCT_FLAG_NO_ADJ_VEC);
      if (pl->adj_vec && !dry) {
        // the interleaved copy of the sinogram is made by every adjoint call into this scratch: allocated HERE, so
        // that xct_adjoint itself never allocates (XLA execute stage, stream capture); without the memory the plan
        // keeps the scalar-tap kernel
        const size_t bytes = (size_t)V * ceil_div(g->n0, 4) * g->d1 * 4 * sizeof(float);
        if (cudaMalloc(&pl->d_sinoT, bytes) != cudaSuccess ||
            cudaEventCreateWithFlags(&pl->sinoT_ev, cudaEventDisableTiming) != cudaSuccess) {
          cudaGetLastError();
          cudaFree(pl->d_sinoT);
          pl->d_sinoT = nullptr;
          pl->adj_vec = false;
        }
      }
      pl->gs = env.fwd_ok ? env.gs : 0;
      pl->pipe_ok = pl->pipe_ok && pl->fwd_walk && pl->adj_walk && !(g->flags & XCT_FLAG_NO_HOST_PIPELINE);
      if (pl->adj_plane && pl->fwd_plane) pl->path = XCT_PATH_3D_SEP;
    }
  }
  if (pl->path == XCT_PATH_3D_GENERAL && !(g->flags & (XCT_FLAG_FORCE_GENERAL | XCT_FLAG_NO_BRICK)) &&
      g->det_row_offset >= 0 && g->det_row_offset + g->d0 <= pl->rows_total) {
    // general matrices: brick kernels wherever every view's brick windows fit (else thread-per-voxel)
    BrickAnalysis ba = analyse_bricks(g->matrices, V, g->n0, g->n1, g->n2, g->slice_offset);
    pl->brick_adj = ba.adj_ok;
    pl->brick_adj_tma = ba.adj_ok && (g->d1 % 4 == 0) && (dry || tensor_map_encoder() != nullptr) && !(g->flags & XCT_FLAG_NO_TMA);
    pl->brick_fwd = ba.fwd_ok;
    if (ba.adj_ok || ba.fwd_ok) {
      std::vector<float> mt(8 * (size_t)V);
      for (int v = 0; v < V; ++v)
        for (int q = 0; q < 4; ++q) {
          mt[8 * (size_t)v + 2 * q] = g->matrices[8 * (size_t)v + q];
          mt[8 * (size_t)v + 2 * q + 1] = g->matrices[8 * (size_t)v + 4 + q];
        }
      e = dev_upload(pl, pl->d_mats_t, mt.data(), mt.size());
      if (e != cudaSuccess) return cleanup(fail(XCT_ERR_CUDA, std::string("matrix upload: ") + cudaGetErrorString(e)));
    }
    if (ba.fwd_ok)
      for (int c = 0; c < 6; ++c) {
        pl->n_listB[c] = (int)ba.list[c].size();
        e = dev_upload(pl, pl->d_listB[c], ba.list[c].data(), ba.list[c].size());
        if (e != cudaSuccess) return cleanup(fail(XCT_ERR_CUDA, std::string("view list upload: ") + cudaGetErrorString(e)));
      }
  }
  *out = pl;
  return XCT_OK;
}

int xct2d_plan_create(xct_plan** out, const xct2d_geom* g) { return plan2d_create_impl(out, g, false); }
int xct3d_plan_create(xct_plan** out, const xct3d_geom* g) { return plan3d_create_impl(out, g, false); }

// Analysis only: every plan decision (kernel family, view classes, TMA / joint eligibility) from the
// geometry alone, without touching a CUDA device -- what the CPU-only tests exercise.
static void fill_classes(const xct_plan* pl, xct_plan_classes* c) {
  for (int k = 0; k < 8; ++k) c->joint_views[k] = pl->n_listJ[k];
  for (int k = 0; k < 4; ++k) c->two_bin_views[k] = (pl->fwd_joint || pl->fwd_joint2d) ? pl->n_listR[k] : pl->n_list4[k];
  c->adj_jump_views = pl->adj_jump_views;
  c->rows_unit = pl->rows_unit ? 1 : 0;
  c->rows_consecutive = pl->rows_krow ? 1 : 0;
  c->fwd_cold = pl->fwd_cold ? 1 : 0;
  for (int k = 0; k < 6; ++k) c->brick_views[k] = pl->n_listB[k];
  c->fwd_tile = (pl->fwd_tile || pl->fwd_tile2d) ? 1 : 0;
  c->adj_interleaved = pl->adj_vec ? 1 : 0;
}
int xct_plan_get_classes(const xct_plan* pl, xct_plan_classes* classes) {
  if (!pl || !classes) return fail(XCT_ERR_INVALID, "null argument");
  fill_classes(pl, classes);
  return XCT_OK;
}
int xct2d_plan_analyse(const xct2d_geom* g, xct_plan_info* info, xct_plan_classes* classes) {
  xct_plan* pl = nullptr;
  int rc = plan2d_create_impl(&pl, g, true);
  if (rc) return rc;
  if (info) rc = xct_plan_get_info(pl, info);
  if (!rc && classes) fill_classes(pl, classes);
  xct_plan_destroy(pl);
  return rc;
}
int xct3d_plan_analyse(const xct3d_geom* g, xct_plan_info* info, xct_plan_classes* classes) {
  xct_plan* pl = nullptr;
  int rc = plan3d_create_impl(&pl, g, true);
  if (rc) return rc;
  if (info) rc = xct_plan_get_info(pl, info);
  if (!rc && classes) fill_classes(pl, classes);
  xct_plan_destroy(pl);
  return rc;
}

void xct_plan_destroy(xct_plan* pl) {
  if (!pl) return;
  if (pl->dry) {
    delete pl;
    return;
  }
  DeviceGuard guard(pl->device);
  cudaFree(pl->d_views);
  cudaFree(pl->d_rows);
  cudaFree(pl->d_mats);
  cudaFree(pl->d_list[0]);
  cudaFree(pl->d_list[1]);
  cudaFree(pl->d_rowoff);
  for (int c = 0; c < 4; ++c) cudaFree(pl->d_list4[c]);
  for (int c = 0; c < 8; ++c) cudaFree(pl->d_listJ[c]);
  for (int c = 0; c < 4; ++c) cudaFree(pl->d_listR[c]);
  for (int c = 0; c < 6; ++c) cudaFree(pl->d_listB[c]);
  cudaFree(pl->d_mats_t);
  cudaFree(pl->d_sinoT);
  if (pl->sinoT_ev) cudaEventDestroy(pl->sinoT_ev);
  for (int d = 0; d < 2; ++d) { cudaFree(pl->stage_in[d]); cudaFree(pl->stage_out[d]); }
  if (pl->hstream) cudaStreamDestroy(pl->hstream);
  if (pl->s_in) cudaStreamDestroy(pl->s_in);
  if (pl->s_out) cudaStreamDestroy(pl->s_out);
  for (cudaEvent_t ev : pl->events) cudaEventDestroy(ev);
  for (int d = 0; d < 2; ++d)
    if (pl->done_ev[d]) cudaEventDestroy(pl->done_ev[d]);
  delete pl;
}

int xct_plan_get_info(const xct_plan* pl, xct_plan_info* info) {
  if (!pl || !info) return fail(XCT_ERR_INVALID, "null argument");
  info->ndim = pl->ndim;
  info->path = pl->path;
  info->num_views = pl->V;
  info->fwd_lane_stride = pl->gs;
  info->row_aligned = pl->row_aligned ? 1 : 0;
  info->device = pl->device;
  info->adj_kernel = pl->adj_walk ? XCT_KERNEL_WALK : (pl->adj_plane ? XCT_KERNEL_PLANE : (pl->brick_adj ? XCT_KERNEL_BRICK : XCT_KERNEL_GENERAL));
  info->fwd_kernel = (pl->fwd_walk || pl->fwd_joint2d) ? XCT_KERNEL_WALK
                     : (pl->fwd_plane ? XCT_KERNEL_PLANE : (pl->brick_fwd ? XCT_KERNEL_BRICK : XCT_KERNEL_GENERAL));
  info->fwd_joint = (pl->fwd_joint || pl->fwd_joint2d) ? 1 : 0;
  info->adj_tma = (pl->adj_tma || pl->brick_adj_tma) ? 1 : 0;
  info->in_elems = (int64_t)in_elems(pl);
  info->out_elems = (int64_t)out_elems(pl);
  info->updates = info->in_elems * pl->V;
  return XCT_OK;
}

int xct_forward(const xct_plan* pl, const float* in, float* out, int32_t batch, void* stream) {
  int rc = check_call(pl, in, out, batch);
  if (rc) return rc;
  DeviceGuard guard(pl->device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  // every forward kernel accumulates with RED: the (possibly uninitialised) output is zeroed first
  XCT_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * out_elems(pl) * batch, st));
  if (pl->ndim == 3) {
    if (pl->fwd_walk) return launch_walk_forward3(pl, in, out, st);
    if (pl->fwd_plane) return launch_plane_forward<xct::Geom3, true, kFwd3S, kFwd3TN>(pl, 1, in, out, st);
    if (pl->brick_fwd) return launch_brick_forward(pl, in, out, st);
    xct::gen3d_forward_kernel<<<general_grid(in_elems(pl)), 256, 0, st>>>(gen3_params(pl), in, out);
    return launch_ok("gen3d_forward_kernel");
  }
  if (pl->fwd_joint2d) return launch_walk2d_forward(pl, batch, in, out, st);
  if (pl->fwd_plane) return launch_plane_forward<xct::Geom2, false, kFwd2S, kFwd2TN>(pl, batch, in, out, st);
  xct::gen2d_forward_kernel<<<general_grid(in_elems(pl) * batch), 256, 0, st>>>(gen2_params(pl, batch), in, out);
  return launch_ok("gen2d_forward_kernel");
}

int xct_adjoint(const xct_plan* pl, const float* in, float* out, int32_t batch, void* stream) {
  int rc = check_call(pl, in, out, batch);
  if (rc) return rc;
  DeviceGuard guard(pl->device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  if (pl->ndim == 3) {
    if (pl->adj_walk && (reinterpret_cast<uintptr_t>(in) & 15) == 0) return launch_walk_adjoint(pl, in, out, st);
    if (pl->adj_plane) return launch_plane_adjoint<xct::Geom3, true, kAdj3S, kAdj3TA>(pl, 1, in, out, st);
    if (pl->brick_adj) return launch_brick_adjoint(pl, in, out, st);
    xct::gen3d_adjoint_kernel<false><<<general_grid(in_elems(pl)), 256, 0, st>>>(gen3_params(pl), in, out, xct::OutRoute{});
    return launch_ok("gen3d_adjoint_kernel");
  }
  if (pl->adj_plane) {
    // a batch shares the geometry: four images per thread amortise the coordinate / weight arithmetic that bounds the
    // single-image kernel (same taps in the same order per image: bit-identical to image-by-image calls)
    if (batch >= 3) return launch_plane_adjoint<xct::Geom2, false, kAdj2bS, kAdj2bTA>(pl, batch, in, out, st);
    return launch_plane_adjoint<xct::Geom2, false, kAdj2S, kAdj2TA>(pl, batch, in, out, st);
  }
  xct::gen2d_adjoint_kernel<false><<<general_grid(in_elems(pl) * batch), 256, 0, st>>>(gen2_params(pl, batch), in, out, xct::OutRoute{});
  return launch_ok("gen2d_adjoint_kernel");
}

// Back projection of a view block whose result rows go straight to their owners (view-block sharding):
// the kernels' epilogue adds each value into the row block that holds it, across NVLink for a peer's
// block, instead of writing a partial volume for a reduce-scatter.  Every adjoint family has a routed
// epilogue: walk (TMA or cp.async staging), plane, brick, thread-per-voxel.
int xct_adjoint_scatter(const xct_plan* pl, const float* in, const xct_out_route* r, void* stream) {
  if (!pl || !in || !r) return fail(XCT_ERR_INVALID, "null argument");
  if (pl->dry) return fail(XCT_ERR_INVALID, "analysis-only plan");
  const int rows = pl->n0;
  if (r->nparts < 1 || r->nparts > XCT_MAX_ROUTE_PARTS) return fail(XCT_ERR_INVALID, "route: nparts outside [1, XCT_MAX_ROUTE_PARTS]");
  if (r->row_begin[0] != 0 || r->row_begin[r->nparts] != rows)
    return fail(XCT_ERR_INVALID, "route: row blocks must cover [0, n0) of the adjoint's result");
  xct::OutRoute route{};
  static_assert(XCT_MAX_ROUTE_PARTS == xct::kMaxRouteParts, "header and kernels disagree");
  route.nparts = r->nparts;
  route.store = r->store ? 1 : 0;
  route.inner = pl->ndim == 3 ? (long long)pl->n1 * pl->n2 : (long long)pl->n1;
  for (int k = 0; k < r->nparts; ++k) {
    if (r->row_begin[k + 1] < r->row_begin[k]) return fail(XCT_ERR_INVALID, "route: row_begin must be non-decreasing");
    if (r->row_begin[k + 1] > r->row_begin[k] && !r->ptr[k]) return fail(XCT_ERR_INVALID, "route: null pointer for a non-empty block");
    route.ptr[k] = r->ptr[k];
    route.row_begin[k] = r->row_begin[k];
  }
  for (int k = r->nparts; k <= xct::kMaxRouteParts; ++k) route.row_begin[k] = rows;
  DeviceGuard guard(pl->device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  if (pl->ndim == 3) {
    if (pl->adj_walk && (reinterpret_cast<uintptr_t>(in) & 15) == 0) return launch_walk_adjoint(pl, in, nullptr, st, 0, -1, &route);
    if (pl->adj_plane) return launch_plane_adjoint<xct::Geom3, true, kAdj3S, kAdj3TA>(pl, 1, in, nullptr, st, &route);
    if (pl->brick_adj) return launch_brick_adjoint(pl, in, nullptr, st, &route);
    xct::gen3d_adjoint_kernel<true><<<general_grid(in_elems(pl)), 256, 0, st>>>(gen3_params(pl), in, nullptr, route);
    return launch_ok("gen3d_adjoint_kernel<route>");
  }
  if (pl->adj_plane) return launch_plane_adjoint<xct::Geom2, false, kAdj2S, kAdj2TA>(pl, 1, in, nullptr, st, &route);
  xct::gen2d_adjoint_kernel<true><<<general_grid(in_elems(pl)), 256, 0, st>>>(gen2_params(pl, 1), in, nullptr, route);
  return launch_ok("gen2d_adjoint_kernel<route>");
}

// ---- device buffers other processes of the node can map (CUDA IPC over NVLink / NVSwitch) ----
int xct_peer_alloc(int32_t device, size_t bytes, void** ptr, xct_ipc_handle* handle) {
  if (!ptr || !handle || bytes == 0) return fail(XCT_ERR_INVALID, "null argument or empty buffer");
  static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(handle->bytes), "xct_ipc_handle must hold a cudaIpcMemHandle_t");
  int rc = check_device(device);
  if (rc) return rc;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  void* p = nullptr;
  XCT_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(XCT_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  std::memcpy(handle->bytes, &h, sizeof(h));
  *ptr = p;
  return XCT_OK;
}
int xct_peer_open(int32_t device, const xct_ipc_handle* handle, void** ptr) {
  if (!ptr || !handle) return fail(XCT_ERR_INVALID, "null argument");
  int rc = check_device(device);
  if (rc) return rc;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle->bytes, sizeof(h));
  XCT_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return XCT_OK;
}
int xct_peer_zero(int32_t device, void* ptr, size_t bytes, void* stream) {
  if (!ptr) return fail(XCT_ERR_INVALID, "null argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  XCT_CUDA(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream));
  return XCT_OK;
}
int xct_peer_copy_out(int32_t device, void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return fail(XCT_ERR_INVALID, "null argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  XCT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return XCT_OK;
}
int xct_sum_slots(int32_t device, float* dst, const float* slots, int32_t nslots, size_t n, size_t pitch, void* stream) {
  if (!dst || !slots || nslots < 1 || pitch < n) return fail(XCT_ERR_INVALID, "xct_sum_slots: bad argument");
  if (n == 0) return XCT_OK;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  xct::sum_slots_kernel<<<general_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(dst, slots, nslots, n, pitch);
  return launch_ok("sum_slots_kernel");
}
int xct_peer_signal(int32_t device, int32_t* const* flag_ptrs, int32_t nranks, int32_t epoch, void* stream) {
  if (!flag_ptrs || nranks < 1 || nranks > XCT_MAX_ROUTE_PARTS) return fail(XCT_ERR_INVALID, "xct_peer_signal: bad argument");
  xct::PeerFlagPtrs f{};
  f.n = nranks;
  for (int k = 0; k < nranks; ++k) {
    if (!flag_ptrs[k]) return fail(XCT_ERR_INVALID, "xct_peer_signal: null flag pointer");
    f.ptr[k] = flag_ptrs[k];
  }
  DeviceGuard guard(device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  xct::peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, epoch);
  return launch_ok("peer_signal_kernel");
}
int xct_peer_wait(int32_t device, const int32_t* flags, int32_t nranks, int32_t epoch, double timeout_s, int32_t* timed_out_dev,
                  void* stream) {
  if (!flags || !timed_out_dev || nranks < 1 || nranks > XCT_MAX_ROUTE_PARTS) return fail(XCT_ERR_INVALID, "xct_peer_wait: bad argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  const unsigned long long ns = (unsigned long long)(std::max(timeout_s, 1e-3) * 1e9);
  xct::peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, nranks, epoch, ns, timed_out_dev);
  return launch_ok("peer_wait_kernel");
}
int xct_peer_close(int32_t device, void* ptr) {
  if (!ptr) return XCT_OK;
  DeviceGuard guard(device);
  XCT_CUDA(cudaIpcCloseMemHandle(ptr));
  return XCT_OK;
}
int xct_peer_free(int32_t device, void* ptr) {
  if (!ptr) return XCT_OK;
  DeviceGuard guard(device);
  XCT_CUDA(cudaFree(ptr));
  return XCT_OK;
}

// Pipelined host path for 3D separable plans with unit, monotone rows: the volume is cut into >= 8
// chunks of slices (32 for the headline shape, 16 for a 128-slice slab); chunk k's kernels (stream hstream) overlap the H2D copy of
// chunk k+1 (stream s_in) and the D2H copy of what chunk k-1 completed (stream s_out).  Detector rows of a
// slice chunk are a row block of every view: strided 2D copies with the view pitch.  Nothing here waits for
// the device: the three streams carry the calls of a plan in order, so a forward and an adjoint enqueued back
// to back (xct_*_host_async) keep all three busy across the call boundary -- the adjoint's first sinogram
// rows go up while the forward's last chunks compute and its last rows come down.
static int enqueue_host_pipelined(xct_plan* pl, const float* in_host, float* out_host, bool forward) {
  int rc;
  const int dir = forward ? 0 : 1;
  const size_t n_vol = in_elems(pl), n_sino = out_elems(pl);
  if ((rc = ensure_stage(pl, dir, forward ? n_vol : n_sino, forward ? n_sino : n_vol))) return rc;
  const int NS = pl->n0, D0 = pl->d0, D1 = pl->d1, V = pl->V;
  int chunk = std::max(16, (NS + 31) / 32);  // 1024 slices, one GPU: chunks of 8 / 16 / 32 / 64 slices -> 553.7 / 553.5 / 547.7 / 552.3 ms per pair
  if (const char* env = std::getenv("XCT_HOST_CHUNK_SLICES")) {  // tuning / diagnosis (tools/bench_host_pipeline.py)
    const int v = std::atoi(env);
    if (v > 0) chunk = v;
  }
  chunk = (chunk + 15) & ~15;  // whole slice groups of both kernels (4 slices forward, 16 adjoint)
  chunk = std::max(chunk, ((NS + 63) / 64 + 15) & ~15);  // at most 64 chunks (event table)
  // slice ranges: chunks of `chunk` slices, except that the first and the last one are halved (whole 16-slice groups):
  // the first H2D copy and the last D2H copy are the only ones nothing overlaps, so they should be short
  std::vector<int> bounds{0};
  {
    const int edge = (chunk >= 32 && NS >= 4 * chunk) ? ((chunk / 2 + 15) & ~15) : chunk;
    int a = 0;
    bool first = true;
    while (a < NS) {
      int left = NS - a;
      const int len = first ? edge : chunk;
      if (len >= left) {  // the end: one short chunk last (boundaries stay multiples of 16 slices)
        if (edge < chunk && left > edge) {
          const int head = (left - edge) & ~15;
          if (head > 0) { a += head; bounds.push_back(a); }
        }
        bounds.push_back(NS);
        break;
      }
      a += len;
      bounds.push_back(a);
      first = false;
    }
  }
  const int nchunks = (int)bounds.size() - 1;
  const size_t ev_base = dir ? 2 * 64 : 0;  // each direction owns its events
  while (pl->events.size() < 4 * 64) {
    cudaEvent_t ev;
    XCT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    pl->events.push_back(ev);
  }
  if (nchunks > 64) return fail(XCT_ERR_INVALID, "host pipeline: more than 64 chunks");
  const size_t slice = (size_t)pl->n1 * pl->n2;
  const size_t vpitch = (size_t)D0 * D1 * sizeof(float);  // bytes between the same row of two views
  float* vol_dev = forward ? pl->stage_in[0] : pl->stage_out[1];
  float* sino_dev = forward ? pl->stage_out[0] : pl->stage_in[1];
  auto ev = [&](int k, int which) { return pl->events[ev_base + 2 * k + which]; };
  auto rows_copy = [&](int r0, int r1, bool to_device, cudaStream_t st) -> cudaError_t {
    if (r1 <= r0) return cudaSuccess;
    const size_t off = (size_t)r0 * D1;
    const size_t width = (size_t)(r1 - r0) * D1 * sizeof(float);
    if (to_device)
      return cudaMemcpy2DAsync(sino_dev + off, vpitch, in_host + off, vpitch, width, V, cudaMemcpyHostToDevice, st);
    return cudaMemcpy2DAsync(out_host + off, vpitch, sino_dev + off, vpitch, width, V, cudaMemcpyDeviceToHost, st);
  };
  // the previous call of THIS direction may still be reading / draining these staging buffers: order after its
  // last D2H (the other direction has its own buffers and is not waited for)
  if (pl->done_ev[dir]) {
    XCT_CUDA(cudaStreamWaitEvent(pl->s_in, pl->done_ev[dir], 0));
    XCT_CUDA(cudaStreamWaitEvent(pl->hstream, pl->done_ev[dir], 0));
  } else {
    XCT_CUDA(cudaEventCreateWithFlags(&pl->done_ev[dir], cudaEventDisableTiming));
  }
  if (forward) {
    XCT_CUDA(cudaMemsetAsync(sino_dev, 0, n_sino * sizeof(float), pl->hstream));
    // first row any slice >= i touches (rows never decrease with the slice index)
    std::vector<int> first_from(NS + 1, D0);
    for (int i = NS - 1; i >= 0; --i) first_from[i] = pl->h_row_lo[i] >= 0 ? std::min(first_from[i + 1], pl->h_row_lo[i]) : first_from[i + 1];
    int rows_done = 0;
    for (int k = 0; k < nchunks; ++k) {
      const int a = bounds[k], b = bounds[k + 1];
      XCT_CUDA(cudaMemcpyAsync(vol_dev + a * slice, in_host + a * slice, (size_t)(b - a) * slice * sizeof(float),
                               cudaMemcpyHostToDevice, pl->s_in));
      XCT_CUDA(cudaEventRecord(ev(k, 0), pl->s_in));
      XCT_CUDA(cudaStreamWaitEvent(pl->hstream, ev(k, 0), 0));
      if ((rc = launch_walk_forward3(pl, vol_dev, sino_dev, pl->hstream, a, b - a))) return rc;
      XCT_CUDA(cudaEventRecord(ev(k, 1), pl->hstream));
      const int complete = b == NS ? D0 : std::max(rows_done, first_from[b]);  // rows no later chunk adds to
      if (complete > rows_done) {
        XCT_CUDA(cudaStreamWaitEvent(pl->s_out, ev(k, 1), 0));
        XCT_CUDA(rows_copy(rows_done, complete, false, pl->s_out));
        rows_done = complete;
      }
    }
  } else {
    int rows_up = 0;  // detector rows [0, rows_up) are on the device
    for (int k = 0; k < nchunks; ++k) {
      const int a = bounds[k], b = bounds[k + 1];
      int need = rows_up;
      for (int i = a; i < b; ++i) need = std::max(need, pl->h_row_hi[i] + 1);
      XCT_CUDA(rows_copy(rows_up, need, true, pl->s_in));
      rows_up = need;
      XCT_CUDA(cudaEventRecord(ev(k, 0), pl->s_in));
      XCT_CUDA(cudaStreamWaitEvent(pl->hstream, ev(k, 0), 0));
      if ((rc = launch_walk_adjoint(pl, sino_dev, vol_dev, pl->hstream, a, b - a))) return rc;
      XCT_CUDA(cudaEventRecord(ev(k, 1), pl->hstream));
      XCT_CUDA(cudaStreamWaitEvent(pl->s_out, ev(k, 1), 0));
      XCT_CUDA(cudaMemcpyAsync(out_host + a * slice, vol_dev + a * slice, (size_t)(b - a) * slice * sizeof(float),
                               cudaMemcpyDeviceToHost, pl->s_out));
    }
  }
  XCT_CUDA(cudaEventRecord(pl->done_ev[dir], pl->s_out));
  return XCT_OK;
}

// One H2D, the kernels, one D2H, all on hstream (every plan that is not pipelined).  The D2H stream is made to
// follow, so that xct_host_wait only has to drain the three streams.
static int enqueue_host_plain(xct_plan* pl, const float* in_host, float* out_host, int32_t batch, bool forward) {
  int rc;
  const int dir = forward ? 0 : 1;
  const size_t n_in = (forward ? in_elems(pl) : out_elems(pl)) * batch;
  const size_t n_out = (forward ? out_elems(pl) : in_elems(pl)) * batch;
  if ((rc = ensure_stage(pl, dir, n_in, n_out))) return rc;
  XCT_CUDA(cudaMemcpyAsync(pl->stage_in[dir], in_host, n_in * sizeof(float), cudaMemcpyHostToDevice, pl->hstream));
  rc = forward ? xct_forward(pl, pl->stage_in[dir], pl->stage_out[dir], batch, pl->hstream)
               : xct_adjoint(pl, pl->stage_in[dir], pl->stage_out[dir], batch, pl->hstream);
  if (rc) return rc;
  XCT_CUDA(cudaMemcpyAsync(out_host, pl->stage_out[dir], n_out * sizeof(float), cudaMemcpyDeviceToHost, pl->hstream));
  return XCT_OK;
}

static int enqueue_host(xct_plan* pl, const float* in_host, float* out_host, int32_t batch, bool forward) {
  int rc = check_call(pl, in_host, out_host, batch);
  if (rc) return rc;
  DeviceGuard guard(pl->device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  if (pl->pipe_ok && pl->n0 >= 64) return enqueue_host_pipelined(pl, in_host, out_host, forward);
  return enqueue_host_plain(pl, in_host, out_host, batch, forward);
}

int xct_host_wait(xct_plan* pl) {
  if (!pl) return fail(XCT_ERR_INVALID, "null plan");
  if (pl->dry) return fail(XCT_ERR_INVALID, "analysis-only plan");
  DeviceGuard guard(pl->device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  if (pl->s_out) XCT_CUDA(cudaStreamSynchronize(pl->s_out));
  if (pl->hstream) XCT_CUDA(cudaStreamSynchronize(pl->hstream));
  if (pl->s_in) XCT_CUDA(cudaStreamSynchronize(pl->s_in));
  return XCT_OK;
}

int xct_forward_host_async(xct_plan* pl, const float* in_host, float* out_host, int32_t batch) {
  return enqueue_host(pl, in_host, out_host, batch, true);
}
int xct_adjoint_host_async(xct_plan* pl, const float* in_host, float* out_host, int32_t batch) {
  return enqueue_host(pl, in_host, out_host, batch, false);
}
int xct_forward_host(xct_plan* pl, const float* in_host, float* out_host, int32_t batch) {
  int rc = enqueue_host(pl, in_host, out_host, batch, true);
  return rc ? rc : xct_host_wait(pl);
}
int xct_adjoint_host(xct_plan* pl, const float* in_host, float* out_host, int32_t batch) {
  int rc = enqueue_host(pl, in_host, out_host, batch, false);
  return rc ? rc : xct_host_wait(pl);
}


// ---------------------------------------------------------------- operator registry (XLA FFI callers)
// An XLA FFI handler gets plain integer attributes, may be executed on ANY device of the process and may
// outlive the Python object that described the geometry.  So the handler is given an operator id, not a
// plan pointer: the registry owns a copy of the geometry and one plan per device, created on first use for
// that device, and destroys them when the last reference is released.
namespace {
struct OpEntry {
  int ndim = 0;
  int refs = 1;
  xct2d_geom g2{};
  xct3d_geom g3{};
  std::vector<float> table;           // 2D view table / 3D matrices (owned copy)
  std::map<int, xct_plan*> plans;     // device ordinal -> plan
};
std::mutex g_op_mutex;
std::map<int64_t, OpEntry> g_ops;
std::atomic<int64_t> g_next_op{1};
}  // namespace

int xct_op_register_2d(const xct2d_geom* g, int64_t* op_id) {
  if (!g || !op_id) return fail(XCT_ERR_INVALID, "null argument");
  xct_plan_info info;
  int rc = xct2d_plan_analyse(g, &info, nullptr);  // validates the geometry without a device
  if (rc) return rc;
  OpEntry e;
  e.ndim = 2;
  e.g2 = *g;
  e.table.assign(g->view_table, g->view_table + 4 * (size_t)g->num_views);
  std::lock_guard<std::mutex> lock(g_op_mutex);
  const int64_t id = g_next_op++;
  g_ops[id] = std::move(e);
  *op_id = id;
  return XCT_OK;
}
int xct_op_register_3d(const xct3d_geom* g, int64_t* op_id) {
  if (!g || !op_id) return fail(XCT_ERR_INVALID, "null argument");
  xct_plan_info info;
  int rc = xct3d_plan_analyse(g, &info, nullptr);
  if (rc) return rc;
  OpEntry e;
  e.ndim = 3;
  e.g3 = *g;
  e.table.assign(g->matrices, g->matrices + 8 * (size_t)g->num_views);
  std::lock_guard<std::mutex> lock(g_op_mutex);
  const int64_t id = g_next_op++;
  g_ops[id] = std::move(e);
  *op_id = id;
  return XCT_OK;
}
int xct_op_retain(int64_t op_id) {
  std::lock_guard<std::mutex> lock(g_op_mutex);
  auto it = g_ops.find(op_id);
  if (it == g_ops.end()) return fail(XCT_ERR_INVALID, "unknown or released operator id");
  ++it->second.refs;
  return XCT_OK;
}
int xct_op_release(int64_t op_id) {
  std::map<int, xct_plan*> dead;
  {
    std::lock_guard<std::mutex> lock(g_op_mutex);
    auto it = g_ops.find(op_id);
    if (it == g_ops.end()) return fail(XCT_ERR_INVALID, "unknown or released operator id");
    if (--it->second.refs > 0) return XCT_OK;
    dead.swap(it->second.plans);
    g_ops.erase(it);
  }
  for (auto& kv : dead) xct_plan_destroy(kv.second);
  return XCT_OK;
}
// The plan of `op_id` on `device`; created (allocates and uploads the geometry tables) on the first call for
// that device -- call it from an initialize-stage handler, or once before the first execution.
int xct_op_plan(int64_t op_id, int32_t device, const xct_plan** plan) {
  if (!plan) return fail(XCT_ERR_INVALID, "null argument");
  *plan = nullptr;
  std::lock_guard<std::mutex> lock(g_op_mutex);
  auto it = g_ops.find(op_id);
  if (it == g_ops.end()) return fail(XCT_ERR_INVALID, "unknown or released operator id");
  OpEntry& e = it->second;
  auto pit = e.plans.find(device);
  if (pit == e.plans.end()) {
    xct_plan* pl = nullptr;
    int rc;
    if (e.ndim == 2) {
      xct2d_geom g = e.g2;
      g.view_table = e.table.data();
      g.device = device;
      rc = xct2d_plan_create(&pl, &g);
    } else {
      xct3d_geom g = e.g3;
      g.matrices = e.table.data();
      g.device = device;
      rc = xct3d_plan_create(&pl, &g);
    }
    if (rc) return rc;
    pit = e.plans.emplace(device, pl).first;
  }
  *plan = pit->second;
  return XCT_OK;
}
// One application on `device`: looks the plan up (it must exist: no allocation here), derives the batch from
// the element count -- leading batch axes of any rank, as jax.vmap with vmap_method="expand_dims" produces --
// and enqueues on `stream`.  2D plans take the batch natively; 3D plans run the items one after the other.
int xct_op_apply(int64_t op_id, int32_t device, int32_t forward, const float* in, float* out, int64_t in_count, void* stream) {
  const xct_plan* pl = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_op_mutex);
    auto it = g_ops.find(op_id);
    if (it == g_ops.end()) return fail(XCT_ERR_INVALID, "unknown or released operator id");
    auto pit = it->second.plans.find(device);
    if (pit == it->second.plans.end())
      return fail(XCT_ERR_INVALID, "operator has no plan on this device (xct_op_plan was not called for it)");
    pl = pit->second;
  }
  const int64_t per_in = forward ? (int64_t)in_elems(pl) : (int64_t)out_elems(pl);
  const int64_t per_out = forward ? (int64_t)out_elems(pl) : (int64_t)in_elems(pl);
  if (in_count < per_in || in_count % per_in != 0) return fail(XCT_ERR_INVALID, "operand size is not a multiple of the operator's input");
  const int64_t batch = in_count / per_in;
  if (batch > INT32_MAX) return fail(XCT_ERR_INVALID, "batch too large");
  if (pl->ndim == 2) return forward ? xct_forward(pl, in, out, (int32_t)batch, stream) : xct_adjoint(pl, in, out, (int32_t)batch, stream);
  for (int64_t b = 0; b < batch; ++b) {
    const int rc = forward ? xct_forward(pl, in + b * per_in, out + b * per_out, 1, stream)
                           : xct_adjoint(pl, in + b * per_in, out + b * per_out, 1, stream);
    if (rc) return rc;
  }
  return XCT_OK;
}

// ---------------------------------------------------------------- TV / PDHG kernels (xct_tv.cuh)
static int tv_check(const xct_tv_block* b, const void* p0, const void* p1) {
  if (!b || !p0 || !p1) return fail(XCT_ERR_INVALID, "null argument");
  if (b->n0 < 1 || b->n1 < 1 || b->n2 < 1) return fail(XCT_ERR_INVALID, "bad volume block shape");
  return XCT_OK;
}
static xct::TvDims tv_dims(const xct_tv_block* b) {
  xct::TvDims d{};
  d.n0 = b->n0; d.n1 = b->n1; d.n2 = b->n2; d.first = b->is_first ? 1 : 0; d.last = b->is_last ? 1 : 0;
  return d;
}
static int tv_grid(size_t n) { return (int)std::min<size_t>((n + 255) / 256, 148u * 32u); }

static int tv_primal_impl(const xct_tv_block* b, float* x, float* xbar, const float* atz, const float* z1,
                          const float* lo_halo, float tau, float alpha, int32_t nonneg, double* stat, void* stream) {
  int rc = tv_check(b, x, xbar);
  if (rc) return rc;
  if (!atz || !z1) return fail(XCT_ERR_INVALID, "null argument");
  const size_t n = (size_t)b->n0 * b->n1 * b->n2;
  if (stat)
    xct::tv_primal_kernel<true><<<tv_grid(n), 256, 0, (cudaStream_t)stream>>>(tv_dims(b), x, xbar, atz, z1, lo_halo, tau, alpha, nonneg, stat);
  else
    xct::tv_primal_kernel<false><<<tv_grid(n), 256, 0, (cudaStream_t)stream>>>(tv_dims(b), x, xbar, atz, z1, lo_halo, tau, alpha, nonneg, nullptr);
  return launch_ok("tv_primal_kernel");
}
int xct_tv_primal_step(const xct_tv_block* b, float* x, float* xbar, const float* atz, const float* z1,
                       const float* lo_halo, float tau, float alpha, int32_t nonneg, void* stream) {
  return tv_primal_impl(b, x, xbar, atz, z1, lo_halo, tau, alpha, nonneg, nullptr, stream);
}
int xct_tv_primal_step_stat(const xct_tv_block* b, float* x, float* xbar, const float* atz, const float* z1,
                            const float* lo_halo, float tau, float alpha, int32_t nonneg, double* sq_dx, void* stream) {
  if (!sq_dx) return fail(XCT_ERR_INVALID, "null statistics pointer");
  return tv_primal_impl(b, x, xbar, atz, z1, lo_halo, tau, alpha, nonneg, sq_dx, stream);
}

static int tv_dual_impl(const xct_tv_block* b, float* z1, const float* xbar, const float* hi_halo, float sigma,
                        float lam, double* stat, void* stream) {
  int rc = tv_check(b, z1, xbar);
  if (rc) return rc;
  if (!(sigma > 0.f)) return fail(XCT_ERR_INVALID, "sigma must be positive");
  const size_t n = (size_t)b->n0 * b->n1 * b->n2;
  if (stat)
    xct::tv_dual_kernel<true><<<tv_grid(n), 256, 0, (cudaStream_t)stream>>>(tv_dims(b), z1, xbar, hi_halo, sigma, lam, stat);
  else
    xct::tv_dual_kernel<false><<<tv_grid(n), 256, 0, (cudaStream_t)stream>>>(tv_dims(b), z1, xbar, hi_halo, sigma, lam, nullptr);
  return launch_ok("tv_dual_kernel");
}
int xct_tv_dual_step(const xct_tv_block* b, float* z1, const float* xbar, const float* hi_halo, float sigma,
                     float lam, void* stream) {
  return tv_dual_impl(b, z1, xbar, hi_halo, sigma, lam, nullptr, stream);
}
int xct_tv_dual_step_stat(const xct_tv_block* b, float* z1, const float* xbar, const float* hi_halo, float sigma,
                          float lam, double* sq_dz1, void* stream) {
  if (!sq_dz1) return fail(XCT_ERR_INVALID, "null statistics pointer");
  return tv_dual_impl(b, z1, xbar, hi_halo, sigma, lam, sq_dz1, stream);
}

int xct_l2_dual_step(int64_t n, float* z0, const float* ax, const float* y, float sigma, void* stream) {
  if (!z0 || !ax || !y || n < 1) return fail(XCT_ERR_INVALID, "null argument or empty array");
  if (!(sigma > 0.f)) return fail(XCT_ERR_INVALID, "sigma must be positive");
  xct::l2_dual_kernel<false><<<tv_grid((size_t)n), 256, 0, (cudaStream_t)stream>>>((size_t)n, z0, ax, y, sigma, nullptr, 0.f,
                                                                                   xct::SinoRows{1, 1, 0, 1}, nullptr);
  return launch_ok("l2_dual_kernel");
}
int xct_l2_dual_step_stat(int64_t n, float* z0, const float* ax, const float* y, float sigma, float* ax_x, float alpha,
                          int64_t inner, int32_t rows, int32_t row_lo, int32_t row_hi, double* stat2, void* stream) {
  if (!z0 || !ax || !y || !ax_x || !stat2 || n < 1) return fail(XCT_ERR_INVALID, "null argument or empty array");
  if (!(sigma > 0.f)) return fail(XCT_ERR_INVALID, "sigma must be positive");
  if (!(alpha > -1.f)) return fail(XCT_ERR_INVALID, "alpha must be above -1");
  if (inner < 1 || rows < 1 || n % (inner * rows) != 0) return fail(XCT_ERR_INVALID, "n is not a multiple of rows * inner");
  xct::l2_dual_kernel<true><<<tv_grid((size_t)n), 256, 0, (cudaStream_t)stream>>>(
      (size_t)n, z0, ax, y, sigma, ax_x, alpha, xct::SinoRows{(long long)inner, rows, row_lo, row_hi}, stat2);
  return launch_ok("l2_dual_kernel");
}

int xct_tv_norm(const xct_tv_block* b, const float* x, const float* hi_halo, double* sum, void* stream) {
  int rc = tv_check(b, x, sum);
  if (rc) return rc;
  const size_t n = (size_t)b->n0 * b->n1 * b->n2;
  xct::tv_norm_kernel<<<tv_grid(n), 256, 0, (cudaStream_t)stream>>>(tv_dims(b), x, hi_halo, sum);
  return launch_ok("tv_norm_kernel");
}

int xct_fd_forward(const xct_tv_block* b, const float* x, const float* hi_halo, float* out, void* stream) {
  int rc = tv_check(b, x, out);
  if (rc) return rc;
  const size_t n = (size_t)b->n0 * b->n1 * b->n2;
  xct::fd_forward_kernel<<<tv_grid(n), 256, 0, (cudaStream_t)stream>>>(tv_dims(b), x, hi_halo, out);
  return launch_ok("fd_forward_kernel");
}

int xct_fd_adjoint(const xct_tv_block* b, const float* z1, const float* lo_halo, float* out, void* stream) {
  int rc = tv_check(b, z1, out);
  if (rc) return rc;
  const size_t n = (size_t)b->n0 * b->n1 * b->n2;
  xct::fd_adjoint_kernel<<<tv_grid(n), 256, 0, (cudaStream_t)stream>>>(tv_dims(b), z1, lo_halo, out);
  return launch_ok("fd_adjoint_kernel");
}

// ------------------------------------------------- ADMM / LADMM / PADMM kernels (xct_solver.cuh)
}  // extern "C"
template <bool STAT>
static int grad_prox_impl(const xct_tv_block* b, const float* x, const float* hi_halo, float* z1, float* u1, float* w1,
                          float dscale, float thr, float inv_nu, int32_t mode, double* stat, void* stream) {
  int rc = tv_check(b, x, z1);
  if (rc) return rc;
  if (!u1 || (mode != XCT_SPLIT_ADMM && !w1)) return fail(XCT_ERR_INVALID, "null argument");
  if (STAT && !stat) return fail(XCT_ERR_INVALID, "null statistics pointer");
  if (!(thr >= 0.f)) return fail(XCT_ERR_INVALID, "threshold must be non-negative");
  const size_t n = (size_t)b->n0 * b->n1 * b->n2;
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case XCT_SPLIT_ADMM:
      xct::grad_prox_kernel<xct::kSplitAdmm, STAT><<<tv_grid(n), 256, 0, st>>>(tv_dims(b), x, hi_halo, z1, u1, w1, dscale, thr, inv_nu, stat);
      break;
    case XCT_SPLIT_LADMM:
      xct::grad_prox_kernel<xct::kSplitLadmm, STAT><<<tv_grid(n), 256, 0, st>>>(tv_dims(b), x, hi_halo, z1, u1, w1, dscale, thr, inv_nu, stat);
      break;
    case XCT_SPLIT_PADMM:
      xct::grad_prox_kernel<xct::kSplitPadmm, STAT><<<tv_grid(n), 256, 0, st>>>(tv_dims(b), x, hi_halo, z1, u1, w1, dscale, thr, inv_nu, stat);
      break;
    default:
      return fail(XCT_ERR_INVALID, "unknown split mode");
  }
  return launch_ok("grad_prox_kernel");
}
extern "C" {
int xct_grad_prox_step(const xct_tv_block* b, const float* x, const float* hi_halo, float* z1, float* u1, float* w1,
                       float dscale, float thr, float inv_nu, int32_t mode, void* stream) {
  return grad_prox_impl<false>(b, x, hi_halo, z1, u1, w1, dscale, thr, inv_nu, mode, nullptr, stream);
}
int xct_grad_prox_step_stat(const xct_tv_block* b, const float* x, const float* hi_halo, float* z1, float* u1, float* w1,
                            float dscale, float thr, float inv_nu, int32_t mode, double* stat3, void* stream) {
  return grad_prox_impl<true>(b, x, hi_halo, z1, u1, w1, dscale, thr, inv_nu, mode, stat3, stream);
}

}  // extern "C"
template <bool STAT>
static int sino_prox_impl(int64_t n, const float* ax, const float* y, float* z0, float* u0, float* w0, float c,
                          float inv_nu, int32_t mode, xct::SinoRows rows, double* stat, void* stream) {
  if (!ax || !y || !z0 || !u0 || !w0 || n < 1) return fail(XCT_ERR_INVALID, "null argument or empty array");
  if (STAT && !stat) return fail(XCT_ERR_INVALID, "null statistics pointer");
  if (rows.inner < 1 || rows.rows < 1 || n % (rows.inner * rows.rows) != 0)
    return fail(XCT_ERR_INVALID, "n is not a multiple of rows * inner");
  if (!(c > 0.f)) return fail(XCT_ERR_INVALID, "prox parameter must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == XCT_SPLIT_LADMM)
    xct::sino_prox_kernel<xct::kSplitLadmm, STAT><<<tv_grid((size_t)n), 256, 0, st>>>((size_t)n, ax, y, z0, u0, w0, c, inv_nu, rows, stat);
  else if (mode == XCT_SPLIT_PADMM)
    xct::sino_prox_kernel<xct::kSplitPadmm, STAT><<<tv_grid((size_t)n), 256, 0, st>>>((size_t)n, ax, y, z0, u0, w0, c, inv_nu, rows, stat);
  else
    return fail(XCT_ERR_INVALID, "unknown split mode");
  return launch_ok("sino_prox_kernel");
}
extern "C" {
int xct_sino_prox_step(int64_t n, const float* ax, const float* y, float* z0, float* u0, float* w0, float c,
                       float inv_nu, int32_t mode, void* stream) {
  return sino_prox_impl<false>(n, ax, y, z0, u0, w0, c, inv_nu, mode, xct::SinoRows{1, 1, 0, 1}, nullptr, stream);
}
int xct_sino_prox_step_stat(int64_t n, const float* ax, const float* y, float* z0, float* u0, float* w0, float c,
                            float inv_nu, int32_t mode, int64_t inner, int32_t rows, int32_t row_lo, int32_t row_hi,
                            double* stat3, void* stream) {
  return sino_prox_impl<true>(n, ax, y, z0, u0, w0, c, inv_nu, mode, xct::SinoRows{(long long)inner, rows, row_lo, row_hi},
                              stat3, stream);
}

int xct_grad_primal_step(const xct_tv_block* b, float* x, const float* atq, const float* w1, const float* lo_halo,
                         float step, float dscale, int32_t nonneg, void* stream) {
  int rc = tv_check(b, x, atq);
  if (rc) return rc;
  if (!w1) return fail(XCT_ERR_INVALID, "null argument");
  const size_t n = (size_t)b->n0 * b->n1 * b->n2;
  xct::grad_primal_kernel<<<tv_grid(n), 256, 0, (cudaStream_t)stream>>>(tv_dims(b), x, atq, w1, lo_halo, step, dscale, nonneg);
  return launch_ok("grad_primal_kernel");
}

int xct_admm_rhs(const xct_tv_block* b, const float* aty, const float* z1, const float* u1, const float* lo_halo,
                 float rho, float* rhs, double* sumsq, void* stream) {
  int rc = tv_check(b, aty, rhs);
  if (rc) return rc;
  if (!z1 || !u1 || !sumsq) return fail(XCT_ERR_INVALID, "null argument");
  const size_t n = (size_t)b->n0 * b->n1 * b->n2;
  xct::admm_rhs_kernel<<<tv_grid(n), 256, 0, (cudaStream_t)stream>>>(tv_dims(b), aty, z1, u1, lo_halo, rho, rhs, sumsq);
  return launch_ok("admm_rhs_kernel");
}

int xct_cg_init(const xct_tv_block* b, const float* x, const float* lo_halo, const float* hi_halo, const float* atax,
                const float* rhs, float rho, float* r, float* p, double* num, void* stream) {
  int rc = tv_check(b, x, atax);
  if (rc) return rc;
  if (!rhs || !r || !p || !num) return fail(XCT_ERR_INVALID, "null argument");
  const size_t n = (size_t)b->n0 * b->n1 * b->n2;
  xct::cg_init_kernel<<<tv_grid(n), 256, 0, (cudaStream_t)stream>>>(tv_dims(b), x, lo_halo, hi_halo, atax, rhs, rho, r, p, num);
  return launch_ok("cg_init_kernel");
}

int xct_cg_lhs(const xct_tv_block* b, const float* p, const float* lo_halo, const float* hi_halo, const float* atap,
               float rho, float* q, double* pq, double* zero_me, void* stream) {
  int rc = tv_check(b, p, atap);
  if (rc) return rc;
  if (!q || !pq) return fail(XCT_ERR_INVALID, "null argument");
  const size_t n = (size_t)b->n0 * b->n1 * b->n2;
  xct::cg_lhs_kernel<<<tv_grid(n), 256, 0, (cudaStream_t)stream>>>(tv_dims(b), p, lo_halo, hi_halo, atap, rho, q, pq, zero_me);
  return launch_ok("cg_lhs_kernel");
}

int xct_cg_update_xr(int64_t n, float* x, float* r, const float* p, const float* q, const double* num, const double* pq,
                     double* num_new, double* zero_me, void* stream) {
  if (!x || !r || !p || !q || !num || !pq || !num_new || n < 1) return fail(XCT_ERR_INVALID, "null argument or empty array");
  xct::cg_xr_kernel<<<tv_grid((size_t)n), 256, 0, (cudaStream_t)stream>>>((size_t)n, x, r, p, q, num, pq, num_new, zero_me);
  return launch_ok("cg_xr_kernel");
}

int xct_cg_update_p(int64_t n, float* p, const float* r, const double* num, const double* num_new, void* stream) {
  if (!p || !r || !num || !num_new || n < 1) return fail(XCT_ERR_INVALID, "null argument or empty array");
  xct::cg_p_kernel<<<tv_grid((size_t)n), 256, 0, (cudaStream_t)stream>>>((size_t)n, p, r, num, num_new);
  return launch_ok("cg_p_kernel");
}

int xct3d_debug_weights(const xct_plan* pl, int32_t view, int32_t* ul, float* w, void* stream) {
  if (!pl || !ul || !w) return fail(XCT_ERR_INVALID, "null argument");
  if (pl->ndim != 3) return fail(XCT_ERR_INVALID, "not a 3D plan");
  if (view < 0 || view >= pl->V) return fail(XCT_ERR_INVALID, "view out of range");
  DeviceGuard guard(pl->device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  if (pl->adj_plane || pl->fwd_plane) {
    xct::sep3d_weights_kernel<<<general_grid(in_elems(pl)), 256, 0, st>>>(
        pl->d_views, pl->d_rows, view, pl->n0, pl->n1, pl->n2, pl->d1, pl->row_off, ul, w);
    return launch_ok("sep3d_weights_kernel");
  }
  xct::gen3d_weights_kernel<<<general_grid(in_elems(pl)), 256, 0, st>>>(gen3_params(pl), view, ul, w);
  return launch_ok("gen3d_weights_kernel");
}

int xct2d_debug_weights(const xct_plan* pl, int32_t view, int32_t* inds, float* w, void* stream) {
  if (!pl || !inds || !w) return fail(XCT_ERR_INVALID, "null argument");
  if (pl->ndim != 2) return fail(XCT_ERR_INVALID, "not a 2D plan");
  if (view < 0 || view >= pl->V) return fail(XCT_ERR_INVALID, "view out of range");
  DeviceGuard guard(pl->device);
  if (!guard.ok) return fail(XCT_ERR_CUDA, "cudaSetDevice failed");
  xct::gen2d_weights_kernel<<<general_grid(in_elems(pl)), 256, 0, (cudaStream_t)stream>>>(
      gen2_params(pl, 1), view, inds, w);
  return launch_ok("gen2d_weights_kernel");
}

}  // extern "C"
