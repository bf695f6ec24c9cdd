/*
 * C restatement of SCICO's native X-ray projectors -- TEST INFRASTRUCTURE (oracle).
 *
 * Restates scico/linop/xray/_xray2d.py:223-351 and _xray3d.py:110-266 (reference paths are
 * relative to /root/reference).  It is bit-identical to oracle/xray_np.py (same fp32 expression
 * trees, same accumulation order when `fused == 0`) and is used (a) to check the CUDA kernels at
 * sizes NumPy cannot reach in seconds and (b) as the timed CPU baseline of bench.py
 * (cpu_baseline.kind = "port": the reference's XLA-CPU build cannot run here, JAX is absent).
 * Nothing under scico_b200/ links or loads this file.
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -ffp-contract=off; no -ffast-math: every
 * product and sum must round to fp32 exactly where the reference's expression tree rounds).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define XO_API __attribute__((visibility("default")))

XO_API int xo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

XO_API void xo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------ 2D */

/* _xray2d.py:331-349 for one pixel; table row = (Pxmin, Pdx0, Pdx1, width). */
static inline void weights_2d(const float *t, int i, int j, int *ind, float *w) {
  float Px = (t[0] + t[1] * (float)i) + t[2] * (float)j;
  float fl = floorf(Px);
  *ind = (int)fl;
  float dist = 1.0f - (Px - fl);
  *w = fminf(dist, t[3]) / t[3];
}

/* _xray2d.py:248-265.  fused==0: two scatter passes per view in the reference's order. */
XO_API void xo_project_2d(const float *im, const float *table, int V, int N0, int N1, int ny,
                          float *out, int fused) {
  memset(out, 0, sizeof(float) * (size_t)V * ny);
#pragma omp parallel for schedule(dynamic, 1)
  for (int a = 0; a < V; ++a) {
    const float *t = table + 4 * (size_t)a;
    float *y = out + (size_t)a * ny;
    if (fused) {
      for (int i = 0; i < N0; ++i)
        for (int j = 0; j < N1; ++j) {
          int ind;
          float w;
          weights_2d(t, i, j, &ind, &w);
          float v = im[(size_t)i * N1 + j];
          if (ind >= 0 && ind < ny) y[ind] += v * w;
          if (ind + 1 >= 0 && ind + 1 < ny) y[ind + 1] += v * (1.0f - w);
        }
    } else {
      for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < N0; ++i)
          for (int j = 0; j < N1; ++j) {
            int ind;
            float w;
            weights_2d(t, i, j, &ind, &w);
            int b = ind + pass;
            float wt = pass ? (1.0f - w) : w;
            if (b >= 0 && b < ny) y[b] += im[(size_t)i * N1 + j] * wt;
          }
    }
  }
}

/* _xray2d.py:292-306: two separate sums over views, added at the end. */
XO_API void xo_back_project_2d(const float *y, const float *table, int V, int N0, int N1, int ny,
                               float *out) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N0; ++i)
    for (int j = 0; j < N1; ++j) {
      float s0 = 0.0f, s1 = 0.0f;
      for (int a = 0; a < V; ++a) {
        int ind;
        float w;
        weights_2d(table + 4 * (size_t)a, i, j, &ind, &w);
        const float *row = y + (size_t)a * ny;
        if (ind >= 0 && ind < ny) s0 = s0 + row[ind] * w;
        if (ind + 1 >= 0 && ind + 1 < ny) s1 = s1 + row[ind + 1] * (1.0f - w);
      }
      out[(size_t)i * N1 + j] = s0 + s1;
    }
}

/* Dump of (ind, weight) for one view: used to diff the CUDA coordinate path bit-for-bit. */
XO_API void xo_weights_2d(const float *table_row, int N0, int N1, int32_t *inds, float *w) {
  for (int i = 0; i < N0; ++i)
    for (int j = 0; j < N1; ++j) {
      int ind;
      weights_2d(table_row, i, j, &ind, &w[(size_t)i * N1 + j]);
      inds[(size_t)i * N1 + j] = ind;
    }
}

/* ------------------------------------------------------------------ 3D */

typedef struct {
  int r0, c0;
  float w[4]; /* ul, ur, ll, lr  (ur = row+1, ll = col+1: _xray3d.py:155-158) */
} taps3d;

/* _xray3d.py:211-264 for one voxel; M = (2,4) row-major; xi already includes +0.5(+offset). */
static inline void weights_3d(const float *M, float xi, float xj, float xk, int D0, int D1,
                              taps3d *o) {
  float tn[2];
  int ul[2];
  for (int r = 0; r < 2; ++r) {
    const float *m = M + 4 * r;
    float Px = ((m[0] * xi + m[1] * xj) + m[2] * xk) + m[3];
    float left = Px - 0.25f;
    tn[r] = fminf(ceilf(left) - left, 0.5f);
    ul[r] = (int)floorf(left);
  }
  o->r0 = ul[0];
  o->c0 = ul[1];
  o->w[0] = (tn[0] * tn[1]) * 4.0f;
  o->w[1] = ((0.5f - tn[0]) * tn[1]) * 4.0f;
  o->w[2] = (tn[0] * (0.5f - tn[1])) * 4.0f;
  o->w[3] = ((0.5f - tn[0]) * (0.5f - tn[1])) * 4.0f;
  int rin0 = ul[0] >= 0 && ul[0] < D0, rin1 = ul[0] + 1 >= 0 && ul[0] + 1 < D0;
  int cin0 = ul[1] >= 0 && ul[1] < D1, cin1 = ul[1] + 1 >= 0 && ul[1] + 1 < D1;
  if (!(rin0 && cin0)) o->w[0] = 0.0f;
  if (!(rin1 && cin0)) o->w[1] = 0.0f;
  if (!(rin0 && cin1)) o->w[2] = 0.0f;
  if (!(rin1 && cin1)) o->w[3] = 0.0f;
}

static inline float coord(int idx, int offset) {
  float x = (float)idx + 0.5f; /* jnp.mgrid + 0.5 */
  if (offset) x = x + (float)offset; /* _xray3d.py:212 */
  return x;
}

/* _xray3d.py:110-159.  fused==0: four scatter passes per view (ul, ur, ll, lr). */
XO_API void xo_project_3d(const float *im, const float *matrices, int V, int N0, int N1, int N2,
                          int D0, int D1, int slice_offset, float *out, int fused) {
  memset(out, 0, sizeof(float) * (size_t)V * D0 * D1);
  static const int dr[4] = {0, 1, 0, 1}, dc[4] = {0, 0, 1, 1};
#pragma omp parallel for schedule(dynamic, 1)
  for (int v = 0; v < V; ++v) {
    const float *M = matrices + 8 * (size_t)v;
    float *p = out + (size_t)v * D0 * D1;
    int npass = fused ? 1 : 4;
    for (int pass = 0; pass < npass; ++pass)
      for (int i = 0; i < N0; ++i) {
        float xi = coord(i, slice_offset);
        for (int j = 0; j < N1; ++j) {
          float xj = coord(j, 0);
          const float *row = im + ((size_t)i * N1 + j) * N2;
          for (int k = 0; k < N2; ++k) {
            taps3d t;
            weights_3d(M, xi, xj, coord(k, 0), D0, D1, &t);
            float val = row[k];
            if (fused) {
              for (int q = 0; q < 4; ++q)
                if (t.w[q] != 0.0f) p[(size_t)(t.r0 + dr[q]) * D1 + t.c0 + dc[q]] += t.w[q] * val;
            } else {
              int r = t.r0 + dr[pass], c = t.c0 + dc[pass];
              if (r >= 0 && r < D0 && c >= 0 && c < D1) p[(size_t)r * D1 + c] += t.w[pass] * val;
            }
          }
        }
      }
  }
}

/* _xray3d.py:161-204: views in ascending order; taps in order ul, ur, ll, lr. */
XO_API void xo_back_project_3d(const float *proj, const float *matrices, int V, int N0, int N1,
                               int N2, int D0, int D1, int slice_offset, float *out) {
  static const int dr[4] = {0, 1, 0, 1}, dc[4] = {0, 0, 1, 1};
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < N0; ++i)
    for (int j = 0; j < N1; ++j) {
      float xi = coord(i, slice_offset), xj = coord(j, 0);
      float *o = out + ((size_t)i * N1 + j) * N2;
      for (int k = 0; k < N2; ++k) {
        float xk = coord(k, 0);
        float acc = 0.0f;
        for (int v = 0; v < V; ++v) {
          taps3d t;
          weights_3d(matrices + 8 * (size_t)v, xi, xj, xk, D0, D1, &t);
          const float *y = proj + (size_t)v * D0 * D1;
          for (int q = 0; q < 4; ++q) {
            int r = t.r0 + dr[q], c = t.c0 + dc[q];
            /* reference clamps the gather index and multiplies by a zero weight */
            r = r < 0 ? 0 : (r >= D0 ? D0 - 1 : r);
            c = c < 0 ? 0 : (c >= D1 ? D1 - 1 : c);
            acc = acc + y[(size_t)r * D1 + c] * t.w[q];
          }
        }
        o[k] = acc;
      }
    }
}

/* Back projection evaluated at selected voxels only (same arithmetic and order as above): lets
 * tests check full-size volumes (1024^3) at a random sample of voxels in seconds. */
XO_API void xo_back_project_3d_points(const float *proj, const float *matrices, int V, int D0, int D1,
                                      int slice_offset, const int32_t *ijk, int npts, float *out) {
  static const int dr[4] = {0, 1, 0, 1}, dc[4] = {0, 0, 1, 1};
#pragma omp parallel for schedule(static)
  for (int p = 0; p < npts; ++p) {
    float xi = coord(ijk[3 * p], slice_offset), xj = coord(ijk[3 * p + 1], 0), xk = coord(ijk[3 * p + 2], 0);
    float acc = 0.0f;
    for (int v = 0; v < V; ++v) {
      taps3d t;
      weights_3d(matrices + 8 * (size_t)v, xi, xj, xk, D0, D1, &t);
      const float *y = proj + (size_t)v * D0 * D1;
      for (int q = 0; q < 4; ++q) {
        int r = t.r0 + dr[q], c = t.c0 + dc[q];
        r = r < 0 ? 0 : (r >= D0 ? D0 - 1 : r);
        c = c < 0 ? 0 : (c >= D1 ? D1 - 1 : c);
        acc = acc + y[(size_t)r * D1 + c] * t.w[q];
      }
    }
    out[p] = acc;
  }
}

/* Dump (r0, c0, 4 weights) for one view: used to diff the CUDA coordinate path bit-for-bit. */
XO_API void xo_weights_3d(const float *M, int N0, int N1, int N2, int D0, int D1, int slice_offset,
                          int32_t *ul, float *w) {
  size_t n = (size_t)N0 * N1 * N2;
  for (int i = 0; i < N0; ++i)
    for (int j = 0; j < N1; ++j)
      for (int k = 0; k < N2; ++k) {
        taps3d t;
        weights_3d(M, coord(i, slice_offset), coord(j, 0), coord(k, 0), D0, D1, &t);
        size_t p = ((size_t)i * N1 + j) * N2 + k;
        ul[p] = t.r0;
        ul[n + p] = t.c0;
        for (int q = 0; q < 4; ++q) w[q * n + p] = t.w[q];
      }
}
