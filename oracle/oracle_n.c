/*
 * oracle_n.c — Oracle-N: CPU restatement of the projective point-to-plane
 * pyramid ICP that the sm_100a kernels implement.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under realsensetracker_b200/ may include,
 * link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * PARITY UNPINNED w.r.t. the reference: the reference (yycho0108/RealsenseTracker)
 * holds no tests, golden vectors or fixtures for this path and cannot be built
 * here (Eigen, nanoflann, ChoUtil, fmt absent), and none of this arithmetic
 * exists in it — its ICP is KD-tree point-to-point (align_icp.cpp:73-161; see
 * oracle_r.cpp for that restatement).  Oracle-N is the *specification* of the
 * north-star pipeline; it borrows only the reference's
 * conventions, each cited where used:
 *   - pose maps src -> dst, read as initial guess and overwritten
 *       rs_tracker/align/src/align_icp.cpp:82,107,156
 *   - invalid depth -> point at the origin   rs_tracker/driver/src/rs_driver.cpp:83-88
 *   - normals oriented so that n.(p - viewpoint) <= 0, viewpoint = camera origin
 *       rs_tracker/common/src/point_cloud_utils.cpp:206-216
 *   - Geman-McClure weight (mu/(r^2+mu))^2    align_icp.cpp:116-118
 *   - Huber loss                               rs_tracker/align/src/align_gicp.cpp:67
 *   - fp32 per-point products, fp64 accumulation  align_icp.cpp:125-130
 *   - too few points -> failure                align_icp.cpp:77-79
 *   - statistics of the last iterate are those of the correspondences BEFORE
 *     the final update                          align_icp.cpp:104-113,157
 *
 * Build: gcc -O2 -ffp-contract=off -mfma (see oracle/Makefile). Every fused
 * multiply-add is an explicit fmaf(); nothing else may be contracted, so the
 * per-pixel fp32 arithmetic is bit-identical to the device code, which uses
 * __fmul_rn/__fadd_rn/__fmaf_rn/__frcp_rn/__fsqrt_rn in the same order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rst_align.h"

typedef struct on_level {
  int32_t w, h;
  float fx, fy, cx, cy, ifx, ify;
} on_level;

/* Pyramid level geometry. Pixel-centre convention for the principal point
 * (same +-0.5 convention as rs_tracker/align/include/rs_tracker/align/sample.hpp:63-64). */
void on_level_info(const rst_intrinsics* K0, int32_t w0, int32_t h0, int32_t level, on_level* L) {
  L->w = w0; L->h = h0;
  L->fx = K0->fx; L->fy = K0->fy; L->cx = K0->cx; L->cy = K0->cy;
  for (int l = 0; l < level; ++l) {
    L->w /= 2; L->h /= 2;
    L->fx = L->fx * 0.5f; L->fy = L->fy * 0.5f;
    L->cx = (L->cx + 0.5f) * 0.5f - 0.5f;
    L->cy = (L->cy + 0.5f) * 0.5f - 0.5f;
  }
  L->ifx = 1.0f / L->fx;
  L->ify = 1.0f / L->fy;
}

/* K6: 2x2 integer depth pooling. out is (w/2) x (h/2), dense. */
void on_pyr_down(const uint16_t* in, int32_t w, int32_t h, int32_t tol, uint16_t* out) {
  const int32_t w2 = w / 2, h2 = h / 2;
  for (int v = 0; v < h2; ++v)
    for (int u = 0; u < w2; ++u) {
      uint32_t d[4] = {in[(2 * v) * w + 2 * u], in[(2 * v) * w + 2 * u + 1],
                       in[(2 * v + 1) * w + 2 * u], in[(2 * v + 1) * w + 2 * u + 1]};
      uint32_t m = 0xFFFFFFFFu;
      for (int k = 0; k < 4; ++k) if (d[k] != 0 && d[k] < m) m = d[k];
      uint32_t sum = 0, n = 0;
      if (m != 0xFFFFFFFFu)
        for (int k = 0; k < 4; ++k) if (d[k] != 0 && d[k] - m <= (uint32_t)tol) { sum += d[k]; ++n; }
      out[v * w2 + u] = n ? (uint16_t)((sum + n / 2) / n) : 0;
    }
}

static inline int z_ok(uint16_t d, const rst_params* P, float* z) {
  *z = (float)d * P->depth_scale;
  return d != 0 && *z >= P->z_min && *z <= P->z_max;
}

/* K1+K2: geometry map G[v][u] = {nx, ny, nz, z}; all zero where the vertex or the
 * normal is invalid. Vertex = (kx*z, ky*z, z), kx = (u - cx) * (1/fx). */
void on_geometry(const uint16_t* D, const on_level* L, const rst_params* P, float* G) {
  const int w = L->w, h = L->h;
  memset(G, 0, sizeof(float) * 4 * (size_t)w * h);
  for (int v = 1; v < h - 1; ++v)
    for (int u = 1; u < w - 1; ++u) {
      float z, zl, zr, zu, zd;
      if (!z_ok(D[v * w + u], P, &z)) continue;
      if (!z_ok(D[v * w + u - 1], P, &zl) || !z_ok(D[v * w + u + 1], P, &zr) ||
          !z_ok(D[(v - 1) * w + u], P, &zu) || !z_ok(D[(v + 1) * w + u], P, &zd))
        continue;
      const float tol = P->normal_depth_tol * z;
      if (!(fabsf(zl - z) <= tol && fabsf(zr - z) <= tol && fabsf(zu - z) <= tol && fabsf(zd - z) <= tol))
        continue;
      const float kxl = ((float)(u - 1) - L->cx) * L->ifx, kxr = ((float)(u + 1) - L->cx) * L->ifx;
      const float kx = ((float)u - L->cx) * L->ifx;
      const float kyu = ((float)(v - 1) - L->cy) * L->ify, kyd = ((float)(v + 1) - L->cy) * L->ify;
      const float ky = ((float)v - L->cy) * L->ify;
      /* a = V(u+1,v) - V(u-1,v), b = V(u,v+1) - V(u,v-1) */
      const float ax = kxr * zr - kxl * zl, ay = ky * zr - ky * zl, az = zr - zl;
      const float bx = kx * zd - kx * zu, by = kyd * zd - kyu * zu, bz = zd - zu;
      /* n = a x b, each component fmaf(p, q, -(r*s)) */
      float nx = fmaf(ay, bz, -(az * by));
      float ny = fmaf(az, bx, -(ax * bz));
      float nz = fmaf(ax, by, -(ay * bx));
      const float len2 = fmaf(nz, nz, fmaf(ny, ny, nx * nx));
      if (!(len2 >= 1e-30f) || !(len2 < INFINITY)) continue; /* degenerate normal */
      float inv = 1.0f / sqrtf(len2);
      /* orient toward the camera: flip if n . V > 0 (point_cloud_utils.cpp:210-214) */
      const float dotv = fmaf(nz, z, fmaf(ny, ky * z, nx * (kx * z)));
      if (dotv > 0.0f) inv = -inv;
      float* g = G + 4 * ((size_t)v * w + u);
      g[0] = nx * inv; g[1] = ny * inv; g[2] = nz * inv; g[3] = z;
    }
}

/* f2: grey level-0 intensity from RGB (CV_8UC3, rs_driver.cpp:212): I = (0.299 R + 0.587 G + 0.114 B) / 255 */
void on_intensity(const uint8_t* rgb, int32_t w, int32_t h, float* I) {
  for (size_t i = 0; i < (size_t)w * h; ++i)
    I[i] = fmaf(0.114f, (float)rgb[3 * i + 2], fmaf(0.587f, (float)rgb[3 * i + 1], 0.299f * (float)rgb[3 * i])) * (1.0f / 255.0f);
}

/* intensity pyramid: plain 2x2 mean, ((a + b) + (c + d)) * 0.25 */
void on_intensity_down(const float* in, int32_t w, int32_t h, float* out) {
  const int32_t w2 = w / 2, h2 = h / 2;
  for (int v = 0; v < h2; ++v)
    for (int u = 0; u < w2; ++u)
      out[v * w2 + u] = ((in[(2 * v) * w + 2 * u] + in[(2 * v) * w + 2 * u + 1]) +
                         (in[(2 * v + 1) * w + 2 * u] + in[(2 * v + 1) * w + 2 * u + 1])) * 0.25f;
}

static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

static inline float robust_w(const rst_params* P, float r) {
  if (P->robust_kind == RST_ROBUST_HUBER) {
    const float a = fabsf(r);
    return a <= P->robust_scale ? 1.0f : P->robust_scale / a;
  }
  if (P->robust_kind == RST_ROBUST_GEMAN_MCCLURE) {
    const float t = P->robust_scale / fmaf(r, r, P->robust_scale);
    return t * t;
  }
  return 1.0f;
}

/* K3+K4: association + normal equations at one level under `pose` (col-major 4x4).
 * src_G may be NULL when the normal gate is disabled. idx may be NULL. */
static void on_evaluate_impl(const uint16_t* src_D, const float* src_G, const float* dst_G, const float* src_I,
                             const float* dst_I, const on_level* L, const rst_params* P, const float* pose,
                             int32_t* idx, rst_stats* st) {
  const int w = L->w, h = L->h;
  const float R00 = pose[0], R10 = pose[1], R20 = pose[2];
  const float R01 = pose[4], R11 = pose[5], R21 = pose[6];
  const float R02 = pose[8], R12 = pose[9], R22 = pose[10];
  const float tx = pose[12], ty = pose[13], tz = pose[14];
  const float dmax2 = P->dist_max * P->dist_max;
  const int use_ngate = P->normal_cos_min > -1.0f;
  const float fw = (float)w, fh = (float)h;
  double A[21], b[6], swr2 = 0.0;
  int64_t count = 0;
  memset(A, 0, sizeof(A));
  memset(b, 0, sizeof(b));
  for (int v = 0; v < h; ++v)
    for (int u = 0; u < w; ++u) {
      const size_t i = (size_t)v * w + u;
      if (idx) idx[i] = -1;
      float z;
      if (!z_ok(src_D[i], P, &z)) continue;
      if (use_ngate && !(src_G[4 * i + 3] > 0.0f)) continue;
      const float kx = ((float)u - L->cx) * L->ifx, ky = ((float)v - L->cy) * L->ify;
      const float px = kx * z, py = ky * z, pz = z;
      const float qx_ = fmaf(R00, px, fmaf(R01, py, fmaf(R02, pz, tx)));
      const float qy_ = fmaf(R10, px, fmaf(R11, py, fmaf(R12, pz, ty)));
      const float qz_ = fmaf(R20, px, fmaf(R21, py, fmaf(R22, pz, tz)));
      if (!(qz_ >= 1e-6f)) continue; /* also keeps 1/qz in the normal range */
      const float iz = 1.0f / qz_;
      const float uf = fmaf(L->fx, qx_ * iz, L->cx);
      const float vf = fmaf(L->fy, qy_ * iz, L->cy);
      /* accept iff the half-to-even rounded pixel lies in the image */
      const float ur = rintf(uf), vr = rintf(vf);
      if (!(ur >= 0.0f && ur <= fw - 1.0f && vr >= 0.0f && vr <= fh - 1.0f)) continue;
      const int ui = (int)ur, vi = (int)vr;
      const float* g = dst_G + 4 * ((size_t)vi * w + ui);
      const float gz = g[3];
      if (!(gz > 0.0f)) continue;
      const float nx = g[0], ny = g[1], nz = g[2];
      const float kxq = ((float)ui - L->cx) * L->ifx, kyq = ((float)vi - L->cy) * L->ify;
      const float dx = fmaf(-kxq, gz, qx_), dy = fmaf(-kyq, gz, qy_), dz = qz_ - gz;
      const float dist2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      if (!(dist2 <= dmax2)) continue;
      if (use_ngate) {
        const float sx = src_G[4 * i], sy = src_G[4 * i + 1], sz = src_G[4 * i + 2];
        const float rx = fmaf(R00, sx, fmaf(R01, sy, R02 * sz));
        const float ry = fmaf(R10, sx, fmaf(R11, sy, R12 * sz));
        const float rz = fmaf(R20, sx, fmaf(R21, sy, R22 * sz));
        const float c = fmaf(rz, nz, fmaf(ry, ny, rx * nx));
        if (!(c >= P->normal_cos_min)) continue;
      }
      if (idx) idx[i] = vi * w + ui;
      /* residual and Jacobian row: r = n.(p' - q), J = [p' x n ; n] */
      const float r = fmaf(nz, dz, fmaf(ny, dy, nx * dx));
      float J[6];
      J[0] = fmaf(qy_, nz, -(qz_ * ny));
      J[1] = fmaf(qz_, nx, -(qx_ * nz));
      J[2] = fmaf(qx_, ny, -(qy_ * nx));
      J[3] = nx; J[4] = ny; J[5] = nz;
      const float wgt = robust_w(P, r);
      int k = 0;
      for (int a = 0; a < 6; ++a) {
        const float wj = wgt * J[a];
        for (int c = a; c < 6; ++c) A[k++] += (double)(wj * J[c]);
        b[a] += (double)(wj * r);
      }
      swr2 += (double)(wgt * r * r);
      ++count;
      if (src_I && dst_I && P->photo_weight > 0.0f) {
        /* photometric row (f2): r_I = I_dst(pi(p')) - I_src(u,v), bilinear sample clamped at the borders
         * (sample.hpp:32-47), residual sign of photometric_cost.hpp:63; J_I = [p' x d ; d] with
         * d = (dI/du fx/z, dI/dv fy/z, -(d_x x + d_y y)/z), scaled by sqrt(lambda). */
        const float x0f = floorf(uf), y0f = floorf(vf);
        const float axf = uf - x0f, ayf = vf - y0f;
        const int x0 = clampi((int)x0f, 0, w - 1), x1 = clampi((int)x0f + 1, 0, w - 1);
        const int y0 = clampi((int)y0f, 0, h - 1), y1 = clampi((int)y0f + 1, 0, h - 1);
        const float I00 = dst_I[(size_t)y0 * w + x0], I10 = dst_I[(size_t)y0 * w + x1];
        const float I01 = dst_I[(size_t)y1 * w + x0], I11 = dst_I[(size_t)y1 * w + x1];
        const float dt = I10 - I00, db = I11 - I01;
        const float top = fmaf(axf, dt, I00), bot = fmaf(axf, db, I01);
        const float gv = bot - top;
        const float val = fmaf(ayf, gv, top);
        const float gu = fmaf(ayf, db - dt, dt);
        const float sl = sqrtf(P->photo_weight);
        const float rI = sl * (val - src_I[i]);
        const float da = sl * ((gu * L->fx) * iz), dbb = sl * ((gv * L->fy) * iz);
        const float dc = -(fmaf(da, qx_, dbb * qy_) * iz);
        float JI[6];
        JI[0] = fmaf(qy_, dc, -(qz_ * dbb));
        JI[1] = fmaf(qz_, da, -(qx_ * dc));
        JI[2] = fmaf(qx_, dbb, -(qy_ * da));
        JI[3] = da; JI[4] = dbb; JI[5] = dc;
        k = 0;
        for (int a = 0; a < 6; ++a) {
          for (int c = a; c < 6; ++c) A[k++] += (double)(JI[a] * JI[c]);
          b[a] += (double)(JI[a] * rI);
        }
        swr2 += (double)(rI * rI);
      }
    }
  memcpy(st->A, A, sizeof(A));
  memcpy(st->b, b, sizeof(b));
  st->sum_wr2 = swr2;
  st->count = (int32_t)count;
  st->rmse = count > 0 ? (float)sqrt(swr2 / (double)count) : 0.0f;
}

void on_evaluate(const uint16_t* src_D, const float* src_G, const float* dst_G, const on_level* L,
                 const rst_params* P, const float* pose, int32_t* idx, rst_stats* st) {
  on_evaluate_impl(src_D, src_G, dst_G, NULL, NULL, L, P, pose, idx, st);
}

/* K3+K4 with the photometric term: src_I / dst_I are the dense intensity maps of this level */
void on_evaluate_photo(const uint16_t* src_D, const float* src_G, const float* dst_G, const float* src_I,
                       const float* dst_I, const on_level* L, const rst_params* P, const float* pose,
                       int32_t* idx, rst_stats* st) {
  on_evaluate_impl(src_D, src_G, dst_G, src_I, dst_I, L, P, pose, idx, st);
}

/* K5: solve (A + damping*I) xi = -b by Cholesky in fp64. Returns a status bit. */
int32_t on_solve(const double* Aut, const double* b, int32_t count, const rst_params* P, double* xi) {
  double M[6][6], Lm[6][6];
  int k = 0;
  double maxdiag = 0.0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) { M[i][j] = M[j][i] = Aut[k++]; }
  for (int i = 0; i < 6; ++i) {
    if (!isfinite(b[i])) return RST_STATUS_NON_FINITE;
    for (int j = 0; j < 6; ++j) if (!isfinite(M[i][j])) return RST_STATUS_NON_FINITE;
    M[i][i] += (double)P->damping;
    if (M[i][i] > maxdiag) maxdiag = M[i][i];
  }
  if (count < P->min_count) return RST_STATUS_TOO_FEW;
  memset(Lm, 0, sizeof(Lm));
  for (int j = 0; j < 6; ++j) {
    double d = M[j][j];
    for (int p = 0; p < j; ++p) d -= Lm[j][p] * Lm[j][p];
    if (!(d > 1e-12 * maxdiag)) return RST_STATUS_DEGENERATE;
    const double l = sqrt(d);
    Lm[j][j] = l;
    for (int i = j + 1; i < 6; ++i) {
      double s = M[i][j];
      for (int p = 0; p < j; ++p) s -= Lm[i][p] * Lm[j][p];
      Lm[i][j] = s / l;
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double s = -b[i];
    for (int p = 0; p < i; ++p) s -= Lm[i][p] * y[p];
    y[i] = s / Lm[i][i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
    for (int p = i + 1; p < 6; ++p) s -= Lm[p][i] * xi[p];
    xi[i] = s / Lm[i][i];
  }
  for (int i = 0; i < 6; ++i) if (!isfinite(xi[i])) return RST_STATUS_NON_FINITE;
  return RST_STATUS_OK;
}

/* T <- Exp(xi) * T, xi = (omega, v), left perturbation, closed-form SE(3)
 * exponential in fp64. Rt = row-major R (9) followed by t (3). */
void on_pose_update(const double* xi, double* Rt) {
  const double wx = xi[0], wy = xi[1], wz = xi[2];
  const double th2 = wx * wx + wy * wy + wz * wz;
  double a, bb, c;
  if (th2 < 1e-8) {
    a = 1.0 - th2 / 6.0; bb = 0.5 - th2 / 24.0; c = 1.0 / 6.0 - th2 / 120.0;
  } else {
    const double th = sqrt(th2);
    a = sin(th) / th; bb = (1.0 - cos(th)) / th2; c = (1.0 - a) / th2;
  }
  const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double W2[9], Rd[9], V[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += W[3 * i + k] * W[3 * k + j];
      W2[3 * i + j] = s;
    }
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    Rd[i] = I + a * W[i] + bb * W2[i];
    V[i] = I + bb * W[i] + c * W2[i];
  }
  double Rn[9], tn[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += Rd[3 * i + k] * Rt[3 * k + j];
      Rn[3 * i + j] = s;
    }
    tn[i] = Rd[3 * i] * Rt[9] + Rd[3 * i + 1] * Rt[10] + Rd[3 * i + 2] * Rt[11] +
            V[3 * i] * xi[3] + V[3 * i + 1] * xi[4] + V[3 * i + 2] * xi[5];
  }
  memcpy(Rt, Rn, sizeof(Rn));
  memcpy(Rt + 9, tn, sizeof(tn));
}

static void rt_to_pose(const double* Rt, float* pose) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) pose[i + 4 * j] = (float)Rt[3 * i + j];
    pose[12 + i] = (float)Rt[9 + i];
    pose[4 * i + 3] = 0.0f;
  }
  pose[15] = 1.0f;
}

/* Full coarse-to-fine alignment of one pair; dense w*h uint16 frames.
 * pose: column-major 4x4 fp32, initial guess in, result out (align_icp.cpp:82,156).
 * Returns the status word (0 = OK). */
static int32_t on_align_pair_impl(const uint16_t* src, const uint16_t* dst, const uint8_t* src_rgb, const uint8_t* dst_rgb,
                                  int32_t w, int32_t h, const rst_intrinsics* K, const rst_params* P, float* pose, rst_stats* st) {
  const int nl = P->num_levels;
  const int photo = src_rgb && dst_rgb && P->photo_weight > 0.0f;
  float* sI[RST_MAX_LEVELS] = {0}; float* dI[RST_MAX_LEVELS] = {0};
  uint16_t* sD[RST_MAX_LEVELS]; uint16_t* dD[RST_MAX_LEVELS];
  float* sG[RST_MAX_LEVELS]; float* dG[RST_MAX_LEVELS];
  on_level L[RST_MAX_LEVELS];
  const int use_ngate = P->normal_cos_min > -1.0f;
  for (int l = 0; l < nl; ++l) {
    on_level_info(K, w, h, l, &L[l]);
    const size_t n = (size_t)L[l].w * L[l].h;
    sD[l] = (uint16_t*)malloc(n * 2); dD[l] = (uint16_t*)malloc(n * 2);
    if (l == 0) { memcpy(sD[0], src, n * 2); memcpy(dD[0], dst, n * 2); }
    else {
      on_pyr_down(sD[l - 1], L[l - 1].w, L[l - 1].h, P->pyr_depth_tol, sD[l]);
      on_pyr_down(dD[l - 1], L[l - 1].w, L[l - 1].h, P->pyr_depth_tol, dD[l]);
    }
    dG[l] = (float*)malloc(n * 16);
    on_geometry(dD[l], &L[l], P, dG[l]);
    sG[l] = NULL;
    if (use_ngate) { sG[l] = (float*)malloc(n * 16); on_geometry(sD[l], &L[l], P, sG[l]); }
    if (photo) {
      sI[l] = (float*)malloc(n * 4); dI[l] = (float*)malloc(n * 4);
      if (l == 0) { on_intensity(src_rgb, w, h, sI[0]); on_intensity(dst_rgb, w, h, dI[0]); }
      else { on_intensity_down(sI[l - 1], L[l - 1].w, L[l - 1].h, sI[l]); on_intensity_down(dI[l - 1], L[l - 1].w, L[l - 1].h, dI[l]); }
    }
  }
  double Rt[12];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Rt[3 * i + j] = (double)pose[i + 4 * j];
    Rt[9 + i] = (double)pose[12 + i];
  }
  rst_stats s;
  memset(&s, 0, sizeof(s));
  int32_t status = 0, any_status = 0, failed = 0, iters = 0;   /* status: last evaluated iteration; any_status: sticky OR */
  float cur[16];
  for (int l = nl - 1; l >= 0; --l) {
    int converged = 0;  /* per level: set once an update is smaller than converge_eps (0 = never) */
    for (int it = 0; it < P->iters[l] && !converged; ++it) {
      rt_to_pose(Rt, cur);
      on_evaluate_impl(sD[l], sG[l], dG[l], sI[l], dI[l], &L[l], P, cur, NULL, &s);
      double xi[6];
      const int32_t rc = on_solve(s.A, s.b, s.count, P, xi);
      status = rc;
      if (rc == RST_STATUS_OK) {
        on_pose_update(xi, Rt);
        const double wn = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
        const double vn = sqrt(xi[3] * xi[3] + xi[4] * xi[4] + xi[5] * xi[5]);
        if (P->converge_eps > 0.0f && wn < (double)P->converge_eps && vn < (double)P->converge_eps) converged = 1;
      } else {
        any_status |= rc; ++failed;
      }
      ++iters;
    }
  }
  rt_to_pose(Rt, pose);
  for (int i = 0; i < 16; ++i) if (!isfinite(pose[i])) { status |= RST_STATUS_NON_FINITE; any_status |= RST_STATUS_NON_FINITE; }
  s.status = status; s.any_status = any_status; s.failed_iterations = failed; s.iterations = iters;
  if (st) *st = s;
  for (int l = 0; l < nl; ++l) { free(sD[l]); free(dD[l]); free(dG[l]); free(sG[l]); free(sI[l]); free(dI[l]); }
  return status;
}

int32_t on_align_pair(const uint16_t* src, const uint16_t* dst, int32_t w, int32_t h,
                      const rst_intrinsics* K, const rst_params* P, float* pose, rst_stats* st) {
  return on_align_pair_impl(src, dst, NULL, NULL, w, h, K, P, pose, st);
}

/* RGB-D alignment: geometric + lambda * photometric (BASELINE config 4). rgb: dense w*h*3 uint8. */
int32_t on_align_pair_rgbd(const uint16_t* src, const uint16_t* dst, const uint8_t* src_rgb, const uint8_t* dst_rgb,
                           int32_t w, int32_t h, const rst_intrinsics* K, const rst_params* P, float* pose, rst_stats* st) {
  return on_align_pair_impl(src, dst, src_rgb, dst_rgb, w, h, K, P, pose, st);
}
