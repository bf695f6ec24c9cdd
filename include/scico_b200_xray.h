/*
 * scico_b200_xray.h -- C ABI of the B200-native X-ray projector pair.
 *
 * This is the drop-in boundary for the hot path of lanl/scico's native CT projectors.  Each
 * entry point replaces one method of the reference (paths relative to the reference checkout):
 *
 *   xct2d_plan_create  <- XRayTransform2D.__init__          scico/linop/xray/_xray2d.py:52-136
 *                         + per-view part of _calc_weights   scico/linop/xray/_xray2d.py:326-347
 *   xct3d_plan_create  <- XRayTransform3D.__init__          scico/linop/xray/_xray3d.py:54-93
 *   xct_forward        <- XRayTransform2D._project           scico/linop/xray/_xray2d.py:223-265
 *                         XRayTransform3D._project           scico/linop/xray/_xray3d.py:110-159
 *   xct_adjoint        <- XRayTransform2D._back_project      scico/linop/xray/_xray2d.py:267-306
 *                         XRayTransform3D._back_project      scico/linop/xray/_xray3d.py:161-204
 *   xct_forward_host / xct_adjoint_host : the same two calls for HOST buffers (what a
 *                         jax.pure_callback / NumPy caller binds; precedent for an external
 *                         projector behind LinearOperator: scico/linop/xray/astra/_astra_3d.py:498-511)
 *   xct3d_debug_weights / xct2d_debug_weights <- XRayTransform3D._calc_weights _xray3d.py:206-266
 *                         / XRayTransform2D._calc_weights _xray2d.py:308-351 (test hook: dumps the
 *                         index/weight arrays the reference materialises, for bit-level diffing)
 *
 * Conventions
 *   - plain C, no torch/JAX types; all arrays are C-contiguous float32, indices int32.
 *   - device entry points take DEVICE pointers, enqueue on `stream` (a cudaStream_t passed as
 *     void*), never synchronise the host and never allocate: they are CUDA-graph capturable and
 *     can be called from an XLA FFI handler (XLA owns the buffers; results are fully overwritten).
 *   - geometry is runtime data (a host table handed to plan_create), never baked at compile time.
 *   - every function returns 0 on success or a negative xct_status; xct_last_error() returns a
 *     thread-local message.  Nothing aborts.
 *   - plans are immutable after creation and may be used concurrently from several threads /
 *     streams of the device they were created on.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SCICO_B200_XRAY_H
#define SCICO_B200_XRAY_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define XCT_API
#else
#define XCT_API __attribute__((visibility("default")))
#endif

#define XCT_VERSION 100 /* 0.1.0 */

typedef enum xct_status {
  XCT_OK = 0,
  XCT_ERR_INVALID = -1,     /* bad argument (shape, null pointer, ...) */
  XCT_ERR_CUDA = -2,        /* CUDA runtime error (message in xct_last_error) */
  XCT_ERR_UNSUPPORTED = -3, /* geometry outside every kernel's envelope */
  XCT_ERR_NO_DEVICE = -4    /* no CUDA device: there is no CPU fallback */
} xct_status;

/* plan_create flags */
#define XCT_FLAG_FORCE_GENERAL 0x1u /* skip the separable fast path (testing / comparison) */
#define XCT_FLAG_NO_WALK 0x2u       /* keep the first-generation plane kernels (testing / comparison) */
#define XCT_FLAG_NO_HOST_PIPELINE 0x4u /* xct_*_host: one H2D, kernels, one D2H (testing / comparison) */
#define XCT_FLAG_NO_TMA 0x10u       /* walk adjoint: stage the sinogram window with cp.async, not TMA (testing / comparison) */
#define XCT_FLAG_NO_JOINT 0x8u      /* walk forward: one column per walk for every view (testing / comparison) */
#define XCT_FLAG_NO_TILE 0x40u      /* 3D joint forward: register-stationary voxels (TN = 8) instead of the CTA-shared tile (testing / comparison) */
#define XCT_FLAG_2D_PER_CLASS 0x80u  /* 2D joint forward: one launch per view class even for small problems (testing / comparison) */
#define XCT_FLAG_NO_ADJ_VEC 0x100u   /* walk adjoint: scalar taps from the (V, D0, D1) sinogram instead of the slice-interleaved copy (testing / comparison) */
#define XCT_FLAG_NO_BRICK 0x20u     /* general 3D matrices: thread-per-voxel kernels instead of the brick kernels (testing / comparison) */

/* kernel families a plan can resolve to (xct_plan_info.path) */
#define XCT_PATH_2D_PLANE 1   /* 2D, warp-autonomous plane kernels */
#define XCT_PATH_2D_GENERAL 2 /* 2D, thread-per-pixel fallback */
#define XCT_PATH_3D_SEP 3     /* 3D, separable geometry: rows <- axis 0 only, cols <- axes 1,2 only */
#define XCT_PATH_3D_GENERAL 4 /* 3D, arbitrary 2x4 matrices */

/* kernel generations (xct_plan_info.adj_kernel / fwd_kernel) */
#define XCT_KERNEL_GENERAL 0 /* thread-per-voxel, any geometry */
#define XCT_KERNEL_PLANE 1   /* warp-autonomous plane kernels (xct_plane.cuh) */
#define XCT_KERNEL_WALK 2    /* register-walk kernels with cp.async staging (xct_plane2.cuh) */
#define XCT_KERNEL_BRICK 3   /* general 3D matrices: brick kernels, TMA-staged detector windows (xct_brick.cuh) */

typedef struct xct_plan xct_plan; /* opaque */

/* 2D geometry: mirrors the constructor of XRayTransform2D after defaults are resolved. */
typedef struct xct2d_geom {
  int32_t n0, n1;          /* image shape (input_shape) */
  int32_t num_views;       /* len(angles) held by this plan (a view block when view-sharded) */
  int32_t det_count;       /* ny */
  const float *view_table; /* HOST (num_views,4) f32: Pxmin, Pdx0, Pdx1, width per view */
  int32_t device;          /* CUDA device ordinal */
  uint32_t flags;
} xct2d_geom;

/* 3D geometry: mirrors XRayTransform3D(input_shape, matrices, det_shape) plus the slab hooks. */
typedef struct xct3d_geom {
  int32_t n0, n1, n2;     /* LOCAL volume shape held in memory */
  int32_t d0, d1;         /* LOCAL detector rows x cols held in memory */
  int32_t num_views;      /* views held by this plan */
  const float *matrices;  /* HOST (num_views,2,4) f32 */
  int32_t slice_offset;   /* added to the axis-0 voxel coordinate (_xray3d.py:212) */
  int32_t det_row_offset; /* global detector row of local row 0 (z-slab sharding), else 0 */
  int32_t det_rows_total; /* global D0 used for the in-bounds test; 0 means d0 */
  int32_t device;
  uint32_t flags;
} xct3d_geom;

typedef struct xct_plan_info {
  int32_t ndim;           /* 2 or 3 */
  int32_t path;           /* XCT_PATH_* */
  int32_t num_views;
  int32_t fwd_lane_stride; /* conflict-free lane stride of the forward kernel (0 = atomics) */
  int32_t row_aligned;    /* 3D sep: every voxel row lands in exactly one detector row */
  int32_t device;
  int32_t adj_kernel;     /* XCT_KERNEL_*: what xct_adjoint launches (16-byte aligned input assumed) */
  int32_t fwd_kernel;     /* XCT_KERNEL_*: what xct_forward launches */
  int32_t fwd_joint;      /* walk forward: views that move by at most one bin per step use the joint-column kernel */
  int32_t adj_tma;        /* walk / brick adjoint: sinogram window staged by TMA (one box per view and tile) */
  int64_t in_elems;       /* elements of one forward input (per batch item) */
  int64_t out_elems;      /* elements of one forward output (per batch item) */
  int64_t updates;        /* voxel-view updates per application = in_elems * num_views */
} xct_plan_info;

/* View classes and row structure a plan derived from its geometry (xct_plan_get_classes, xct*_plan_analyse). */
typedef struct xct_plan_classes {
  int32_t joint_views[8];   /* joint-column forward: views per class [4*major_is_axis_b + 2*minor_up + major_positive] */
  int32_t two_bin_views[4]; /* one-column walk forward: views per class [2*major_is_axis_b + minor_up] */
  int32_t adj_jump_views;   /* views whose |ca| is within rounding distance of 1 (walk adjoint's jump-by-two variant) */
  int32_t rows_unit;        /* 3D sep: every (view, slice) lands in one detector row with weight 2 */
  int32_t rows_consecutive; /* 3D sep: local row = local slice + const per view (TMA box / krow flush) */
  int32_t fwd_cold;         /* some minor coefficient can move the bin by two per step */
  int32_t brick_views[6];   /* general 3D (brick forward): views per class [2*depth_axis + needs_shared_atomics] */
  int32_t fwd_tile;         /* 3D joint forward runs on a CTA-shared tile of 64 x 64 x 4 voxels (walk_forward_tile_kernel) */
  int32_t adj_interleaved;  /* walk adjoint reads a slice-interleaved (V, D0 / 4, D1, 4) copy of the sinogram (walk_adjoint_vec_kernel) */
} xct_plan_classes;

XCT_API int xct_version(void);
XCT_API const char *xct_last_error(void);
XCT_API int xct_device_count(void);

XCT_API int xct2d_plan_create(xct_plan **plan, const xct2d_geom *geom);
XCT_API int xct3d_plan_create(xct_plan **plan, const xct3d_geom *geom);
XCT_API void xct_plan_destroy(xct_plan *plan);
XCT_API int xct_plan_get_info(const xct_plan *plan, xct_plan_info *info);
XCT_API int xct_plan_get_classes(const xct_plan *plan, xct_plan_classes *classes);
/* Analysis only: the decisions xct*_plan_create would take for this geometry (kernel families, view
 * classes, joint / TMA eligibility), computed on the host without touching a CUDA device; `device` of the
 * geometry is ignored, either output may be NULL.  adj_tma reports eligibility (the driver's tensor-map
 * encoder is looked up only by a real plan). */
XCT_API int xct2d_plan_analyse(const xct2d_geom *geom, xct_plan_info *info, xct_plan_classes *classes);
XCT_API int xct3d_plan_analyse(const xct3d_geom *geom, xct_plan_info *info, xct_plan_classes *classes);

/* Forward projection.  in: (batch, *input_shape)  out: (batch, *output_shape), DEVICE pointers.
 * `batch` must be 1 for 3D plans.  `out` is fully overwritten (it may be uninitialised). */
XCT_API int xct_forward(const xct_plan *plan, const float *in_dev, float *out_dev, int32_t batch,
                        void *stream);
/* Back projection (exact adjoint).  in: (batch, *output_shape)  out: (batch, *input_shape).
 * 3D separable plans whose classes report adj_interleaved own a scratch of the sinogram's size (the detector rows
 * of four consecutive slices interleaved per bin), allocated by xct3d_plan_create and rewritten by every call:
 * xct_adjoint itself never allocates; calls on different streams are ordered on the scratch by an event (not inside
 * a stream capture, where the graph's own order applies). */
XCT_API int xct_adjoint(const xct_plan *plan, const float *in_dev, float *out_dev, int32_t batch,
                        void *stream);

/* ---- view-block sharding: back projection fused with its exchange step ----------------------------
 * With the views split over the GPUs of a node (BASELINE.json configs[2]; general 3D matrices) every GPU
 * back-projects its view block onto the WHOLE image / volume, and the partial results have to be summed
 * into the row block (axis 0) each GPU owns -- the reduce-scatter of SURVEY.md 8e.  xct_adjoint_scatter
 * does both in one kernel: its epilogue writes every value straight into the owner's memory (through a
 * CUDA-IPC mapping, i.e. over NVLink / NVSwitch for a peer's block), so no partial volume is written, sent
 * and summed by a separate collective.  Protocol (scico_b200/sharded.py::PeerBlocks): every rank calls
 * xct_adjoint_scatter on one of two alternating copies of the blocks, all ranks meet once (stream-ordered),
 * every owner sums the slots it received (store mode) or copies its block out and re-zeroes it (add mode).
 * Replaces, for this mode, XRayTransform*._back_project (_xray2d.py:267-306, _xray3d.py:161-204) followed
 * by the cross-device sum the reference leaves to jax.Array sharding. */
#define XCT_MAX_ROUTE_PARTS 16
typedef struct xct_out_route {
  int32_t nparts;                             /* row blocks = GPUs of the node */
  int32_t row_begin[XCT_MAX_ROUTE_PARTS + 1]; /* block k holds rows [row_begin[k], row_begin[k+1]) of axis 0 */
  float *ptr[XCT_MAX_ROUTE_PARTS];            /* DEVICE pointer (local or peer-mapped) of block k: (rows_k, *trailing dims) */
  int32_t store; /* 0: values are ADDED to the blocks (RED.ADD.F32, system scope; zero them first, every rank
                  *    targets the same block).  1: values are STORED (each element exactly once per call):
                  *    ptr[k] is then this rank's own slot in the owner's staging area and the owner sums the
                  *    slots afterwards with xct_sum_slots (posted NVLink writes, deterministic sum order). */
} xct_out_route;
/* in: (*output_shape) of the plan, DEVICE pointer.  One image / volume per call (no batch).  Asynchronous on
 * `stream`, no allocation, no host synchronisation. */
XCT_API int xct_adjoint_scatter(const xct_plan *plan, const float *in_dev, const xct_out_route *route, void *stream);

/* Device buffers the other processes of the node can map (cudaMalloc + CUDA IPC; one process per GPU). */
typedef struct xct_ipc_handle { unsigned char bytes[64]; } xct_ipc_handle;
XCT_API int xct_peer_alloc(int32_t device, size_t bytes, void **ptr, xct_ipc_handle *handle);
XCT_API int xct_peer_open(int32_t device, const xct_ipc_handle *handle, void **ptr); /* another process's buffer */
XCT_API int xct_peer_zero(int32_t device, void *ptr, size_t bytes, void *stream);                       /* async memset 0 */
XCT_API int xct_peer_copy_out(int32_t device, void *dst_dev, const void *src, size_t bytes, void *stream); /* async D2D copy */
/* dst[i] = ((slot_0[i] + slot_1[i]) + ...) for i < n; slot s starts at slots + s * pitch (elements). */
XCT_API int xct_sum_slots(int32_t device, float *dst_dev, const float *slots_dev, int32_t nslots, size_t n, size_t pitch,
                          void *stream);
/* Rendezvous of the fused exchange without a collective library call: flag words in peer-mapped memory.
 * xct_peer_signal: stream-ordered after the routed back projection, writes `epoch` into flag_ptrs[k] (this rank's
 * word in rank k's flag array, k < nranks; release, system scope).  xct_peer_wait: stream-ordered, returns once
 * flags[k] >= epoch for all k < nranks (this rank's own array: every peer has signalled), or sets *timed_out_dev = 1
 * after timeout_s seconds (a peer died) instead of hanging the device.  Epochs must grow from call to call. */
XCT_API int xct_peer_signal(int32_t device, int32_t *const *flag_ptrs, int32_t nranks, int32_t epoch, void *stream);
XCT_API int xct_peer_wait(int32_t device, const int32_t *flags_dev, int32_t nranks, int32_t epoch, double timeout_s,
                          int32_t *timed_out_dev, void *stream);
XCT_API int xct_peer_close(int32_t device, void *ptr); /* unmap (xct_peer_open) */
XCT_API int xct_peer_free(int32_t device, void *ptr);  /* release (xct_peer_alloc) */

/* ---- operator registry: for callers that can only carry an integer (XLA FFI custom-call attributes) ----
 * scico wires an external projector as a pair of callables with custom VJP / transpose rules
 * (scico/linop/xray/astra/_astra_3d.py:498-511); under jax.jit those become custom calls whose attributes are plain
 * integers, which may execute on any device of the process and may outlive the Python operator.  The registry
 * therefore owns a COPY of the geometry and one plan per device:
 *   xct_op_register_2d/3d : validate + copy the geometry (geom->device is ignored), returns an id (never reused);
 *   xct_op_plan           : the plan for a device, created on the first call for that device (allocates: call it
 *                           from an initialize-stage FFI handler or once per device before the first execution);
 *   xct_op_apply          : one application on a device's stream; the plan must exist (no allocation, no host
 *                           sync); the batch is in_count / operator input size (leading batch axes of any rank);
 *   xct_op_retain/release : reference count; the last release destroys the plans of every device.  Applying a
 *                           released id returns XCT_ERR_INVALID (no dangling pointer). */
XCT_API int xct_op_register_2d(const xct2d_geom *geom, int64_t *op_id);
XCT_API int xct_op_register_3d(const xct3d_geom *geom, int64_t *op_id);
XCT_API int xct_op_retain(int64_t op_id);
XCT_API int xct_op_release(int64_t op_id);
XCT_API int xct_op_plan(int64_t op_id, int32_t device, const xct_plan **plan);
XCT_API int xct_op_apply(int64_t op_id, int32_t device, int32_t forward, const float *in_dev, float *out_dev, int64_t in_count,
                         void *stream);

/* Same two operators for HOST buffers: H2D copy, kernel(s), D2H copy, synchronous on return.
 * Device staging buffers are cached inside the plan (these two calls are therefore NOT
 * re-entrant on one plan).  3D separable plans on the walk kernels cut the volume into chunks of
 * slices (at least 8 chunks, 16 for a slab of 128 slices or more) and overlap the H2D copy of chunk k+1 and
 * the D2H copy of chunk k-1 with the kernels of chunk k (page-locked host buffers are needed for the overlap;
 * pageable ones still work). */
XCT_API int xct_forward_host(xct_plan *plan, const float *in_host, float *out_host, int32_t batch);
XCT_API int xct_adjoint_host(xct_plan *plan, const float *in_host, float *out_host, int32_t batch);
/* The same two calls WITHOUT the final wait: they enqueue the copies and kernels on the plan's own streams
 * and return.  Calls on one plan execute in the order they were made; a forward and an adjoint enqueued back
 * to back share the copy engines and the SMs across the call boundary (each direction has its own staging
 * buffers).  The host buffers must be page-locked and stay untouched until xct_host_wait(plan) returns;
 * out_host is complete after it.  Not re-entrant on one plan (call from one thread at a time). */
XCT_API int xct_forward_host_async(xct_plan *plan, const float *in_host, float *out_host, int32_t batch);
XCT_API int xct_adjoint_host_async(xct_plan *plan, const float *in_host, float *out_host, int32_t batch);
XCT_API int xct_host_wait(xct_plan *plan);

/* Test hooks: dump what the reference's _calc_weights materialises, as computed by the code path
 * the plan resolved to.  DEVICE outputs.
 *   3D: ul (2, n0,n1,n2) int32 [row, col], w (4, n0,n1,n2) f32 [ul, ur, ll, lr], masked.
 *   2D: inds (n0,n1) int32, w (n0,n1) f32 (tap-0 weight, unmasked, as the reference returns it). */
XCT_API int xct3d_debug_weights(const xct_plan *plan, int32_t view, int32_t *ul_dev, float *w_dev,
                                void *stream);
XCT_API int xct2d_debug_weights(const xct_plan *plan, int32_t view, int32_t *inds_dev, float *w_dev,
                                void *stream);

/* ---- TV-regularised PDHG around the projector pair --------------------------
 * Fused device kernels for C = VerticalStack((A, D)), D = FiniteDifference(append=0)
 * (scico/linop/_diff.py:25-96), g = Separable(SquaredL2Loss(y), lam * L21Norm())
 * (scico/loss.py:220-226, scico/functional/_norm.py:254-263), f = ZeroFunctional or
 * NonNegativeIndicator; one PDHG iteration (scico/optimize/_primaldual.py:219-231) is
 *   atz = xct_adjoint(z0);  xct_tv_primal_step;  ax = xct_forward(xbar);  xct_tv_dual_step;  xct_l2_dual_step.
 * All pointers are DEVICE pointers, float32, C-contiguous; nothing allocates or synchronises. */
typedef struct xct_tv_block {
  int32_t n0, n1, n2; /* LOCAL volume block (a z-slab of the global volume when sharded) */
  int32_t is_first;   /* block holds the global first slice along axis 0 */
  int32_t is_last;    /* block holds the global last slice along axis 0 */
} xct_tv_block;

/* x <- prox_{tau f}(x - tau (atz + D^T z1)),  xbar <- (1 + alpha) x_new - alpha x_old.
 * z1: (3, n0, n1, n2).  lo_halo: plane z1[0][-1] of the previous slab (NULL when is_first). */
XCT_API int xct_tv_primal_step(const xct_tv_block *blk, float *x, float *xbar, const float *atz,
                               const float *z1, const float *lo_halo, float tau, float alpha,
                               int32_t nonneg, void *stream);
/* z1 <- conj_prox_{sigma, lam ||.||_{2,1}}(z1 + sigma D xbar).  hi_halo: plane xbar[n0] of the next
 * slab (NULL when is_last). */
XCT_API int xct_tv_dual_step(const xct_tv_block *blk, float *z1, const float *xbar, const float *hi_halo,
                             float sigma, float lam, void *stream);
/* z0 <- conj_prox_{sigma, 1/2 ||. - y||^2}(z0 + sigma ax), n elements. */
XCT_API int xct_l2_dual_step(int64_t n, float *z0, const float *ax, const float *y, float sigma, void *stream);
/* The same three steps with the iteration statistics of PDHG (scico/optimize/_primaldual.py:175-217:
 * Objective = g(C x), Prml_Rsdl = ||x - x_old|| / tau, Dual_Rsdl = ||z - z_old|| / sigma) accumulated in the
 * same pass into DEVICE doubles (added to: the caller zeroes them), so statistics cost neither copies of the old
 * iterates nor a second forward projection:
 *   primal: *sq_dx  += ||x_new - x_old||^2;      dual: *sq_dz1 += ||z1_new - z1_old||^2;
 *   l2:     stat2[0] += ||z0_new - z0_old||^2;   ax_x <- (ax + alpha ax_x) / (1 + alpha) (= A x_new when ax_x held
 *           A x_old: ax is A xbar and xbar = (1 + alpha) x_new - alpha x_old);   stat2[1] += ||ax_x - y||^2.
 * The sinogram sums run over detector rows [row_lo, row_hi) of the (n / (rows inner), rows, inner) block (rows a
 * z-slab shares with another slab are counted by their owner; pass inner = n, rows = 1, 0, 1 for everything).
 * xct_tv_norm: *sum += ||D x||_{2,1} (L21Norm, l2_axis = 0, scico/functional/_norm.py:225-252). */
XCT_API int xct_tv_primal_step_stat(const xct_tv_block *blk, float *x, float *xbar, const float *atz,
                                    const float *z1, const float *lo_halo, float tau, float alpha,
                                    int32_t nonneg, double *sq_dx, void *stream);
XCT_API int xct_tv_dual_step_stat(const xct_tv_block *blk, float *z1, const float *xbar, const float *hi_halo,
                                  float sigma, float lam, double *sq_dz1, void *stream);
XCT_API int xct_l2_dual_step_stat(int64_t n, float *z0, const float *ax, const float *y, float sigma, float *ax_x,
                                  float alpha, int64_t inner, int32_t rows, int32_t row_lo, int32_t row_hi,
                                  double *stat2, void *stream);
XCT_API int xct_tv_norm(const xct_tv_block *blk, const float *x, const float *hi_halo, double *sum, void *stream);
/* FiniteDifference(append=0): out (3, n0, n1, n2) = D x;  out (n0, n1, n2) = D^T z1. */
XCT_API int xct_fd_forward(const xct_tv_block *blk, const float *x, const float *hi_halo, float *out, void *stream);
XCT_API int xct_fd_adjoint(const xct_tv_block *blk, const float *z1, const float *lo_halo, float *out, void *stream);

/* ---- ADMM family around the projector pair ----------------------------------
 * Fused device kernels for the TV-regularised CT problems the reference's examples solve with
 *   ADMM + CG            scico/optimize/_admm.py:334-378, _admmaux.py:231-269, scico/solver.py:367-405
 *                        (examples/scripts/ct_tv_admm.py: f = SquaredL2Loss(y, A), g = lam L21Norm, C = D)
 *   LinearizedADMM       scico/optimize/_ladmm.py:253-277   (C = VerticalStack((A, D)))
 *   ProximalADMM         scico/optimize/_padmm.py:349-363   (examples/scripts/ct_3d_tv_padmm.py:
 *                        A = VerticalStack((C, alpha D)), B = -I, g = Separable(SquaredL2Loss(y), (lam/alpha) L21Norm))
 * D = FiniteDifference(append=0).  Gradient-shaped arrays are (3, n0, n1, n2); a 2D image is the block
 * (1, n0, n1) (its axis-0 component stays exactly zero).  DEVICE pointers, float32; reductions go to
 * DEVICE doubles; nothing allocates or synchronises. */
#define XCT_SPLIT_ADMM 0
#define XCT_SPLIT_LADMM 1
#define XCT_SPLIT_PADMM 2

/* Gradient block.  Cx = dscale * D x;
 *   ADMM/LADMM: z1 <- prox_{thr ||.||_{2,1}}(Cx + u1);                       u1 <- (u1 + Cx) - z1
 *   PADMM:      z1 <- prox_{thr ||.||_{2,1}}(z1 + inv_nu ((Cx - z1) + u1));  u1 <- (u1 + Cx) - z1
 *   w1 (LADMM) <- (Cx - z1) + u1;  w1 (PADMM) <- 2 u1_new - u1_old;  w1 unused (may be NULL) for ADMM.
 * hi_halo: plane x[n0] of the next slab (NULL when is_last). */
XCT_API int xct_grad_prox_step(const xct_tv_block *blk, const float *x, const float *hi_halo, float *z1, float *u1,
                               float *w1, float dscale, float thr, float inv_nu, int32_t mode, void *stream);
/* Sinogram block, g0 = 1/2 ||. - y||^2: prox_{c g0}(v) = (c y + v) / (c + 1) (scico/loss.py:220-226);
 * z0, u0, w0 updated as above with Cx = ax.  mode: XCT_SPLIT_LADMM or XCT_SPLIT_PADMM. */
XCT_API int xct_sino_prox_step(int64_t n, const float *ax, const float *y, float *z0, float *u0, float *w0, float c,
                               float inv_nu, int32_t mode, void *stream);
/* The two prox steps with the iteration statistics of ProximalADMM / LinearizedADMM (scico/optimize/_padmm.py:148-177,
 * 294-345 with the default fast_dual_residual; _ladmm.py:160-200) accumulated in the same pass into three DEVICE
 * doubles each (added to: the caller zeroes them):
 *   stat3[0] += ||Cx - z_new||^2        (primal residual ||A x + B z||, B = -I)
 *   stat3[1] += ||z_new - z_old||^2     (fast dual residual)
 *   stat3[2] += g(z_new)'s sum: ||z1_new||_{2,1} (gradient block), ||z0_new - y||^2 (sinogram block)
 * Sinogram sums run over detector rows [row_lo, row_hi) of the (n / (rows inner), rows, inner) block, see
 * xct_l2_dual_step_stat. */
XCT_API int xct_grad_prox_step_stat(const xct_tv_block *blk, const float *x, const float *hi_halo, float *z1, float *u1,
                                    float *w1, float dscale, float thr, float inv_nu, int32_t mode, double *stat3,
                                    void *stream);
XCT_API int xct_sino_prox_step_stat(int64_t n, const float *ax, const float *y, float *z0, float *u0, float *w0, float c,
                                    float inv_nu, int32_t mode, int64_t inner, int32_t rows, int32_t row_lo,
                                    int32_t row_hi, double *stat3, void *stream);
/* x <- prox_f(x - step (atq + dscale D^T w1)), f = 0 or the non-negativity indicator.
 * lo_halo: plane w1[0][-1] of the previous slab (NULL when is_first). */
XCT_API int xct_grad_primal_step(const xct_tv_block *blk, float *x, const float *atq, const float *w1,
                                 const float *lo_halo, float step, float dscale, int32_t nonneg, void *stream);
/* ADMM x-step right-hand side: rhs <- aty + rho D^T (z1 - u1);  *sumsq += ||rhs||^2.
 * lo_halo: plane (z1 - u1)[0][-1] of the previous slab. */
XCT_API int xct_admm_rhs(const xct_tv_block *blk, const float *aty, const float *z1, const float *u1,
                         const float *lo_halo, float rho, float *rhs, double *sumsq, void *stream);
/* CG on (A^T A + rho D^T D) x = rhs (scico/solver.py:367-405); the caller applies the projector pair.
 *   init:      r <- rhs - (rho D^T D x + atax);  p <- r;  *num += r.r
 *   lhs:       q <- rho D^T D p + atap;  *pq += p.q;  *zero_me <- 0 (may be NULL)
 *   update_xr: alpha = *num / *pq;  x += alpha p;  r -= alpha q;  *num_new += r.r;  *zero_me <- 0
 *   update_p:  beta = *num_new / *num;  p <- r + beta p
 * lo_halo / hi_halo: planes [-1] / [n0] of the neighbouring slabs of the array D^T D is applied to. */
XCT_API int xct_cg_init(const xct_tv_block *blk, const float *x, const float *lo_halo, const float *hi_halo,
                        const float *atax, const float *rhs, float rho, float *r, float *p, double *num, void *stream);
XCT_API int xct_cg_lhs(const xct_tv_block *blk, const float *p, const float *lo_halo, const float *hi_halo,
                       const float *atap, float rho, float *q, double *pq, double *zero_me, void *stream);
XCT_API int xct_cg_update_xr(int64_t n, float *x, float *r, const float *p, const float *q, const double *num,
                             const double *pq, double *num_new, double *zero_me, void *stream);
XCT_API int xct_cg_update_p(int64_t n, float *p, const float *r, const double *num, const double *num_new,
                            void *stream);

/* Number of this library's kernels launched by the calling thread since the last reset
 * (bench.py's gpu_launches claim). */
XCT_API int64_t xct_launch_count(void);
XCT_API void xct_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* SCICO_B200_XRAY_H */
