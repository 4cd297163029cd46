/*
 * pcfm_oracle.c — CPU restatement of the reference's hot-path algorithms.  TEST INFRASTRUCTURE ONLY:
 * nothing under oracle/ may be imported, linked or executed by the product (point-cloud-flow-matching_b200/);
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Every function follows one reference kernel (file:line under /root/reference given at each function) loop
 * for loop, in plain scalar C.  Floating point mirrors what nvcc emits for the reference on sm_100a: the sums
 * of three squares are FMUL+FFMA+FFMA (checked in the SASS of oracle/_ref), written here with
 * fmaf() and compiled with -ffp-contract=off so the C compiler adds no contraction of its own.
 *
 * Pinning: indices / integer outputs are compared bit-exactly with golden vectors produced by the compiled
 * reference kernels on a B200 (tests/golden/*.npz, generator tests/golden/make_golden.py) and with the
 * reference's own known answers (unit_test.py, test_emd_loss.py, README CD(x,x)=0).
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -o oracle/_build/libpcfm_oracle.so oracle/pcfm_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

static inline float sq3(float dx, float dy, float dz) { return fmaf(dz, dz, fmaf(dx, dx, dy * dy)); }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* ---- avg_voxelize: vox.cu:18-72 (forward), vox.cu:86-110 (backward) ---------------------------------- */
API void oracle_avg_voxelize_forward(const float* feat, const int* coords, int b, int c, int n, int r, float* out,
                                     int* ind, int* cnt) {
  const int r2 = r * r, s = r2 * r;
  memset(out, 0, sizeof(float) * (size_t)b * c * s);
  memset(cnt, 0, sizeof(int) * (size_t)b * s);
#pragma omp parallel for
  for (int bi = 0; bi < b; ++bi) {
    const int* co = coords + (size_t)bi * 3 * n;
    int* in = ind + (size_t)bi * n;
    int* cn = cnt + (size_t)bi * s;
    for (int i = 0; i < n; ++i) { /* grid_stats_kernel vox.cu:28-33 */
      in[i] = co[i] * r2 + co[i + n] * r + co[i + n + n];
      cn[in[i]] += 1;
    }
    for (int j = 0; j < c; ++j) { /* avg_voxelize_kernel vox.cu:60-71, atomics applied in ascending i */
      const float* f = feat + ((size_t)bi * c + j) * n;
      float* o = out + ((size_t)bi * c + j) * s;
      for (int i = 0; i < n; ++i) {
        const int pos = in[i], cur = cn[pos];
        if (cur > 0) {
          const float div = (float)(1.0 / (double)(float)cur); /* vox.cu:65 `1.0 / static_cast<float>(cur_cnt)` */
          o[pos] += f[i] * div;
        }
      }
    }
  }
}

API void oracle_avg_voxelize_backward(const float* grad_y, const int* ind, const int* cnt, int b, int c, int n, int s,
                                      float* grad_x) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < c; ++j) {
      const int* in = ind + (size_t)bi * n;
      const int* cn = cnt + (size_t)bi * s;
      const float* gy = grad_y + ((size_t)bi * c + j) * s;
      float* gx = grad_x + ((size_t)bi * c + j) * n;
      for (int i = 0; i < n; ++i) {
        const int pos = in[i], cur = cn[pos];
        gx[i] = 0.f;
        if (cur > 0) gx[i] += gy[pos] * (float)(1.0 / (double)(float)cur); /* vox.cu:103-107 */
      }
    }
}

/* ---- trilinear_devoxelize: trilinear_devox.cu:21-105 (forward), :119-162 (backward) ------------------ */
API void oracle_trilinear_devoxelize_forward(const float* coords, const float* feat, int b, int c, int n, int r,
                                             int training, float* outs, int* inds, float* wgts) {
  const int r2 = r * r, r3 = r2 * r;
#pragma omp parallel for
  for (int bi = 0; bi < b; ++bi) {
    const float* co = coords + (size_t)bi * 3 * n;
    for (int i = 0; i < n; ++i) {
      const float x = co[i], y = co[i + n], z = co[i + n + n];
      const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
      const float x1 = x - xl, y1 = y - yl, z1 = z - zl;
      const float x0 = 1.0f - x1, y0 = 1.0f - y1, z0 = 1.0f - z1;
      float w[8];
      w[0] = x0 * y0 * z0; w[1] = x0 * y0 * z1; w[2] = x0 * y1 * z0; w[3] = x0 * y1 * z1;
      w[4] = x1 * y0 * z0; w[5] = x1 * y0 * z1; w[6] = x1 * y1 * z0; w[7] = x1 * y1 * z1;
      const int xh = (x1 > 0) ? -1 : 0, yh = (y1 > 0) ? -1 : 0, zh = (z1 > 0) ? 1 : 0; /* :64-66 */
      int id[8];
      id[0] = (int)xl * r2 + (int)yl * r + (int)zl;
      id[1] = id[0] + zh;
      id[2] = id[0] + (yh & r);
      id[3] = id[2] + zh;
      id[4] = id[0] + (xh & r2);
      id[5] = id[4] + zh;
      id[6] = id[4] + (yh & r);
      id[7] = id[6] + zh;
      if (training)
        for (int k = 0; k < 8; ++k) {
          wgts[((size_t)bi * 8 + k) * n + i] = w[k];
          inds[((size_t)bi * 8 + k) * n + i] = id[k];
        }
      for (int j = 0; j < c; ++j) {
        const float* f = feat + ((size_t)bi * c + j) * r3;
        float acc = 0.f;
        for (int k = 0; k < 8; ++k) acc += w[k] * f[id[k]]; /* :98-103; association differs from SASS (<=1e-6) */
        outs[((size_t)bi * c + j) * n + i] = acc;
      }
    }
  }
}

API void oracle_trilinear_devoxelize_backward(const float* grad_y, const int* inds, const float* wgts, int b, int c,
                                              int n, int r3, float* grad_x) {
  memset(grad_x, 0, sizeof(float) * (size_t)b * c * r3);
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < c; ++j) {
      float* gx = grad_x + ((size_t)bi * c + j) * r3;
      const float* gy = grad_y + ((size_t)bi * c + j) * n;
      for (int i = 0; i < n; ++i)
        for (int k = 0; k < 8; ++k) /* :150-159 */
          gx[inds[((size_t)bi * 8 + k) * n + i]] += wgts[((size_t)bi * 8 + k) * n + i] * gy[i];
    }
}

/* ---- ball_query: ball_query.cu:19-50 -------------------------------------------------------------- */
API void oracle_ball_query(const float* centers, const float* points, int b, int n, int m, float radius, int u,
                           int* out) {
  const float r2 = radius * radius; /* ball_query.cpp:24 */
  memset(out, 0, sizeof(int) * (size_t)b * m * u);
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < m; ++j) {
      const float* ce = centers + (size_t)bi * 3 * m;
      const float* pt = points + (size_t)bi * 3 * n;
      int* o = out + ((size_t)bi * m + j) * u;
      const float cx = ce[j], cy = ce[j + m], cz = ce[j + m + m];
      for (int k = 0, cnt = 0; k < n && cnt < u; ++k) {
        const float d2 = sq3(cx - pt[k], cy - pt[k + n], cz - pt[k + n + n]);
        if (d2 < r2) {
          if (cnt == 0)
            for (int v = 0; v < u; ++v) o[v] = k;
          o[cnt] = k;
          ++cnt;
        }
      }
    }
}

/* ---- grouping: grouping.cu:18-36, 58-77; gather: sampling.cu:17-31, 52-66 ----------------------------- */
API void oracle_grouping_forward(const float* feat, const int* idx, int b, int c, int n, int m, int u, float* out) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j)
        for (int k = 0; k < u; ++k)
          out[(((size_t)bi * c + l) * m + j) * u + k] =
              feat[((size_t)bi * c + l) * n + idx[((size_t)bi * m + j) * u + k]];
}
API void oracle_grouping_backward(const float* grad_y, const int* idx, int b, int c, int n, int m, int u,
                                  float* grad_x) {
  memset(grad_x, 0, sizeof(float) * (size_t)b * c * n);
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j)
        for (int k = 0; k < u; ++k)
          grad_x[((size_t)bi * c + l) * n + idx[((size_t)bi * m + j) * u + k]] +=
              grad_y[(((size_t)bi * c + l) * m + j) * u + k];
}
API void oracle_gather_forward(const float* feat, const int* idx, int b, int c, int n, int m, float* out) {
  oracle_grouping_forward(feat, idx, b, c, n, m, 1, out);
}
API void oracle_gather_backward(const float* grad_y, const int* idx, int b, int c, int n, int m, float* grad_x) {
  oracle_grouping_backward(grad_y, idx, b, c, n, m, 1, grad_x);
}

/* ---- furthest point sampling: sampling.cu:86-167, literal 512-thread emulation ------------------------ */
API void oracle_furthest_point_sampling(const float* coords, int b, int n, int m, int* indices) {
  const int BS = 512;
#pragma omp parallel for
  for (int bi = 0; bi < b; ++bi) {
    if (m <= 0) continue;
    const float* co = coords + (size_t)bi * 3 * n;
    int* out = indices + (size_t)bi * m;
    float* dist = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    float dists[512];
    int dists_i[512];
    for (int k = 0; k < n; ++k) dist[k] = 1e38f; /* sampling.cpp:53 */
    int old = 0;
    out[0] = old;
    for (int j = 1; j < m; ++j) {
      const float x1 = co[old], y1 = co[old + n], z1 = co[old + n + n];
      for (int t = 0; t < BS; ++t) { /* each emulated thread: sampling.cu:120-146 */
        int besti = 0;
        float best = -1;
        for (int k = t; k < n; k += BS) {
          const float td = dist[k];
          const float d = sq3(co[k] - x1, co[k + n] - y1, co[k + n + n] - z1);
          const float d2 = fminf(d, td);
          if (d2 != td) dist[k] = d2;
          if (d2 > best) { best = d2; besti = k; }
        }
        dists[t] = best;
        dists_i[t] = besti;
      }
      for (int u = 0; (1 << u) < BS; ++u) /* tree reduce: sampling.cu:150-160 */
        for (int t = 0; t < (BS >> (u + 1)); ++t) {
          const int i1 = (t * 2) << u, i2 = (t * 2 + 1) << u;
          if (dists[i1] < dists[i2]) { dists[i1] = dists[i2]; dists_i[i1] = dists_i[i2]; }
        }
      old = dists_i[0];
      out[j] = old;
    }
    free(dist);
  }
}

/* ---- three nearest neighbours + interpolate: neighbor_interpolate.cu:20-116, 145-170 ------------------- */
API void oracle_three_nn(const float* points, const float* centers, int b, int n, int m, float* weights,
                         int* indices) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < n; ++j) {
      const float* pt = points + (size_t)bi * 3 * n;
      const float* ce = centers + (size_t)bi * 3 * m;
      const float ux = pt[j], uy = pt[j + n], uz = pt[j + n + n];
      double best0 = 1e40, best1 = 1e40, best2 = 1e40;
      int i0 = 0, i1 = 0, i2 = 0;
      for (int k = 0; k < m; ++k) {
        const float d = sq3(ux - ce[k], uy - ce[k + m], uz - ce[k + m + m]);
        if (d < best2) {
          best2 = d; i2 = k;
          if (d < best1) {
            best2 = best1; i2 = i1; best1 = d; i1 = k;
            if (d < best0) { best1 = best0; i1 = i0; best0 = d; i0 = k; }
          }
        }
      }
      best0 = fmax(fmin((double)1e10f, best0), (double)1e-10f);
      best1 = fmax(fmin((double)1e10f, best1), (double)1e-10f);
      best2 = fmax(fmin((double)1e10f, best2), (double)1e-10f);
      const float d0d1 = (float)(best0 * best1), d0d2 = (float)(best0 * best2), d1d2 = (float)(best1 * best2);
      const float inv = 1.0f / (d0d1 + d0d2 + d1d2);
      float* w = weights + (size_t)bi * 3 * n;
      int* ix = indices + (size_t)bi * 3 * n;
      w[j] = d1d2 * inv; ix[j] = i0;
      w[j + n] = d0d2 * inv; ix[j + n] = i1;
      w[j + n + n] = d0d1 * inv; ix[j + n + n] = i2;
    }
}
API void oracle_three_nn_interpolate(const float* feat, const int* idx, const float* w, int b, int c, int m, int n,
                                     float* out) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float* f = feat + ((size_t)bi * c + l) * m;
      const int* ix = idx + (size_t)bi * 3 * n;
      const float* ww = w + (size_t)bi * 3 * n;
      for (int j = 0; j < n; ++j)
        out[((size_t)bi * c + l) * n + j] =
            f[ix[j]] * ww[j] + f[ix[j + n]] * ww[j + n] + f[ix[j + n + n]] * ww[j + n + n];
    }
}
API void oracle_three_nn_interpolate_backward(const float* grad_y, const int* idx, const float* w, int b, int c,
                                              int n, int m, float* grad_x) {
  memset(grad_x, 0, sizeof(float) * (size_t)b * c * m);
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      float* gx = grad_x + ((size_t)bi * c + l) * m;
      const float* gy = grad_y + ((size_t)bi * c + l) * n;
      const int* ix = idx + (size_t)bi * 3 * n;
      const float* ww = w + (size_t)bi * 3 * n;
      for (int j = 0; j < n; ++j) {
        gx[ix[j]] += gy[j] * ww[j];
        gx[ix[j + n]] += gy[j] * ww[j + n];
        gx[ix[j + n + n]] += gy[j] * ww[j + n + n];
      }
    }
}

/* ---- chamfer: chamfer3D.cu:12-134 (one direction, chunks of 512 kept literally), :155-174 (gradient) --- */
static void nm_distance(int b, int n, const float* xyz, int m, const float* xyz2, float* result, int* result_i) {
  const int batch = 512;
#pragma omp parallel for collapse(2)
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const float x1 = xyz[((size_t)i * n + j) * 3 + 0], y1 = xyz[((size_t)i * n + j) * 3 + 1],
                  z1 = xyz[((size_t)i * n + j) * 3 + 2];
      for (int k2 = 0; k2 < m; k2 += batch) {
        const int end_k = imin(m, k2 + batch) - k2;
        const float* buf = xyz2 + ((size_t)i * m + k2) * 3;
        int best_i = 0;
        float best = 0;
        for (int k = 0; k < end_k; ++k) { /* unrolled x4 in the reference; same comparisons in the same order */
          const float d = sq3(buf[k * 3 + 0] - x1, buf[k * 3 + 1] - y1, buf[k * 3 + 2] - z1);
          if (k == 0 || d < best) { best = d; best_i = k + k2; }
        }
        if (k2 == 0 || result[(size_t)i * n + j] > best) { /* chamfer3D.cu:126-129 */
          result[(size_t)i * n + j] = best;
          result_i[(size_t)i * n + j] = best_i;
        }
      }
    }
}
API void oracle_chamfer_forward(const float* xyz1, const float* xyz2, int b, int n, int m, float* dist1, float* dist2,
                                int* idx1, int* idx2) {
  nm_distance(b, n, xyz1, m, xyz2, dist1, idx1); /* chamfer3D.cu:142-143 */
  nm_distance(b, m, xyz2, n, xyz1, dist2, idx2);
}
static void nm_grad(int b, int n, const float* xyz1, int m, const float* xyz2, const float* gd, const int* idx,
                    float* g1, float* g2) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const size_t a = ((size_t)i * n + j) * 3;
      const int j2 = idx[(size_t)i * n + j];
      const size_t c2 = ((size_t)i * m + j2) * 3;
      const float g = gd[(size_t)i * n + j] * 2;
      for (int d = 0; d < 3; ++d) {
        const float t = g * (xyz1[a + d] - xyz2[c2 + d]);
        g1[a + d] += t;
        g2[c2 + d] += -t;
      }
    }
}
API void oracle_chamfer_backward(const float* xyz1, const float* xyz2, const float* gd1, const float* gd2,
                                 const int* idx1, const int* idx2, int b, int n, int m, float* g1, float* g2) {
  nm_grad(b, n, xyz1, m, xyz2, gd1, idx1, g1, g2); /* accumulates, like the reference (caller zeroes) */
  nm_grad(b, m, xyz2, n, xyz1, gd2, idx2, g2, g1);
}

/* ---- chamfer for D-dimensional points: chamfer6D.cu:11-183 (chunks of 2048; 2D: 512, 5D: 2048), :185-216 ------
 * d = x*x + y*y + z*z + ... is contracted by nvcc into FMUL(y,y); FFMA(x,x,.); FFMA(z,z,.); ... (SASS of the compiled
 * reference).  The reference's unrolled loop bound `end_k-(end_k&6)` (chamfer6D.cu:28) runs past the end of a
 * chunk whose length is 1 or 3 mod 8 (it can return index m); this restatement (like the CUDA kernel) takes the true
 * minimum. */
static void nm_distance_nd(int dim, int b, int n, const float* xyz, int m, const float* xyz2, float* result,
                           int* result_i) {
  const int batch = dim == 2 ? 512 : 2048;
#pragma omp parallel for collapse(2)
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const float* a = xyz + ((size_t)i * n + j) * dim;
      for (int k2 = 0; k2 < m; k2 += batch) {
        const int end_k = imin(m, k2 + batch) - k2;
        const float* buf = xyz2 + ((size_t)i * m + k2) * dim;
        int best_i = 0;
        float best = 0;
        for (int k = 0; k < end_k; ++k) {
          const float* o = buf + (size_t)k * dim;
          float d = (o[1] - a[1]) * (o[1] - a[1]);
          d = fmaf(o[0] - a[0], o[0] - a[0], d);
          for (int c = 2; c < dim; ++c) d = fmaf(o[c] - a[c], o[c] - a[c], d);
          if (k == 0 || d < best) { best = d; best_i = k + k2; }
        }
        if (k2 == 0 || result[(size_t)i * n + j] > best) {
          result[(size_t)i * n + j] = best;
          result_i[(size_t)i * n + j] = best_i;
        }
      }
    }
}
API void oracle_chamfer_nd_forward(const float* xyz1, const float* xyz2, int b, int n, int m, int dim, float* dist1,
                                   float* dist2, int* idx1, int* idx2) {
  nm_distance_nd(dim, b, n, xyz1, m, xyz2, dist1, idx1);
  nm_distance_nd(dim, b, m, xyz2, n, xyz1, dist2, idx2);
}
static void nm_grad_nd(int dim, int b, int n, const float* xyz1, int m, const float* xyz2, const float* gd,
                       const int* idx, float* g1, float* g2) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const size_t a = ((size_t)i * n + j) * dim;
      const int j2 = idx[(size_t)i * n + j];
      const size_t c2 = ((size_t)i * m + j2) * dim;
      const float g = gd[(size_t)i * n + j] * 2;
      for (int d = 0; d < dim; ++d) {
        const float t = g * (xyz1[a + d] - xyz2[c2 + d]);
        g1[a + d] += t;
        g2[c2 + d] += -t;
      }
    }
}
API void oracle_chamfer_nd_backward(const float* xyz1, const float* xyz2, const float* gd1, const float* gd2,
                                    const int* idx1, const int* idx2, int b, int n, int m, int dim, float* g1,
                                    float* g2) {
  nm_grad_nd(dim, b, n, xyz1, m, xyz2, gd1, idx1, g1, g2);
  nm_grad_nd(dim, b, m, xyz2, n, xyz1, gd2, idx2, g2, g1);
}

/* ---- EMD: emd_kernel.cu:24-156 (approxmatch), :199-241 (matchcost), :285-353 (gradients) -------------- */
/* --use_fast_math (PyTorchEMD/backend.py:20): __expf(x) = ex2.approx.ftz(x * log2(e)) with the product rounded to
 * fp32 first.  That rounding is the dominant deviation from libm expf for large |x|, so it is restated; the
 * <= 2 ulp error of the MUFU unit itself cannot be. */
static inline float fast_expf(float x) { return exp2f(x * 1.4426950408889634f); }
static inline float d2f(const float* a, const float* c) {
  const float dx = c[0] - a[0], dy = c[1] - a[1], dz = c[2] - a[2];
  return dx * dx + dy * dy + dz * dz;
}
API void oracle_emd_approxmatch(const float* xyz1, const float* xyz2, int b, int n, int m, float* match) {
#pragma omp parallel for
  for (int i = 0; i < b; ++i) {
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * m * 3;
    float* mt = match + (size_t)i * n * m;
    float* remainL = (float*)malloc(sizeof(float) * (size_t)(2 * (n + m) + 4));
    float* remainR = remainL + n;
    float* ratioL = remainR + m;
    float* ratioR = ratioL + n;
    float multiL, multiR;
    if (n >= m) { multiL = 1; multiR = (float)(n / m); } else { multiL = (float)(m / n); multiR = 1; }
    memset(mt, 0, sizeof(float) * (size_t)n * m);
    for (int k = 0; k < n; ++k) remainL[k] = multiL;
    for (int l = 0; l < m; ++l) remainR[l] = multiR;
    for (int j = 7; j >= -2; --j) {
      float level = -powf(4.0f, (float)j);
      if (j == -2) level = 0;
      for (int k = 0; k < n; ++k) {
        float suml = 1e-9f;
        for (int l = 0; l < m; ++l) suml += fast_expf(level * d2f(p1 + k * 3, p2 + l * 3)) * remainR[l];
        ratioL[k] = remainL[k] / suml;
      }
      for (int l = 0; l < m; ++l) {
        float sumr = 0;
        for (int k = 0; k < n; ++k) sumr += fast_expf(level * d2f(p1 + k * 3, p2 + l * 3)) * ratioL[k];
        sumr *= remainR[l];
        const float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
        ratioR[l] = consumption * remainR[l];
        remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
      }
      for (int k = 0; k < n; ++k) {
        float suml = 0;
        const float rl = ratioL[k];
        for (int l = 0; l < m; ++l) {
          const float w = fast_expf(level * d2f(p1 + k * 3, p2 + l * 3)) * rl * ratioR[l];
          mt[(size_t)l * n + k] += w;
          suml += w;
        }
        remainL[k] = fmaxf(0.0f, remainL[k] - suml);
      }
    }
    free(remainL);
  }
}
API void oracle_emd_matchcost(const float* xyz1, const float* xyz2, const float* match, int b, int n, int m,
                              float* cost) {
#pragma omp parallel for
  for (int i = 0; i < b; ++i) {
    double acc = 0; /* the reference reduces 512 fp32 partials in a tree; double keeps the oracle order-free */
    for (int k = 0; k < n; ++k)
      for (int l = 0; l < m; ++l)
        acc += (double)d2f(xyz1 + ((size_t)i * n + k) * 3, xyz2 + ((size_t)i * m + l) * 3) *
               match[(size_t)i * n * m + (size_t)l * n + k];
    cost[i] = (float)acc;
  }
}
API void oracle_emd_matchcost_backward(const float* grad_cost, const float* xyz1, const float* xyz2,
                                       const float* match, int b, int n, int m, float* grad1, float* grad2) {
#pragma omp parallel for
  for (int i = 0; i < b; ++i) {
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * m * 3;
    const float* mt = match + (size_t)i * n * m;
    for (int k = 0; k < n; ++k) { /* matchcostgrad1: emd_kernel.cu:332-353 */
      double d[3] = {0, 0, 0};
      for (int l = 0; l < m; ++l)
        for (int a = 0; a < 3; ++a) d[a] += (double)(p1[k * 3 + a] - p2[l * 3 + a]) * (mt[(size_t)l * n + k] * 2);
      for (int a = 0; a < 3; ++a) grad1[((size_t)i * n + k) * 3 + a] = (float)d[a] * grad_cost[i];
    }
    for (int l = 0; l < m; ++l) { /* matchcostgrad2: emd_kernel.cu:285-325 */
      double d[3] = {0, 0, 0};
      for (int k = 0; k < n; ++k)
        for (int a = 0; a < 3; ++a) d[a] += (double)(p2[l * 3 + a] - p1[k * 3 + a]) * (mt[(size_t)l * n + k] * 2);
      for (int a = 0; a < 3; ++a) grad2[((size_t)i * m + l) * 3 + a] = (float)d[a] * grad_cost[i];
    }
  }
}
