/* rst_synth.c — analytic-scene ray caster; see rst_synth.h. */
#include "rst_synth.h"

#include <math.h>
#include <string.h>

/* splitmix64: counter-based, so pixel noise is independent of traversal order */
static uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
static double u01(uint64_t h) { return (double)((h >> 11) + 1) * (1.0 / 9007199254740994.0); }

void rst_synth_scene_default(uint64_t seed, rst_synth_scene* s) {
  memset(s, 0, sizeof(*s));
  s->room_lo[0] = -3.0; s->room_lo[1] = -1.5; s->room_lo[2] = -1.0;
  s->room_hi[0] = 3.0;  s->room_hi[1] = 1.5;  s->room_hi[2] = 4.0;
  const double sc[3][4] = {{-1.0, 0.6, 2.5, 0.60}, {1.2, 0.2, 3.0, 0.50}, {0.2, -0.6, 2.0, 0.35}};
  s->n_spheres = 3;
  for (int i = 0; i < 3; ++i) {
    for (int k = 0; k < 3; ++k) {
      double j = (u01(mix64(seed * 1315423911ull + 17u * i + k)) - 0.5) * 0.3;
      s->sphere_c[i][k] = sc[i][k] + (seed ? j : 0.0);
    }
    s->sphere_r[i] = sc[i][3];
  }
  s->n_boxes = 1;
  s->box_c[0][0] = -1.6; s->box_c[0][1] = 1.0; s->box_c[0][2] = 3.2;
  s->box_h[0][0] = 0.4;  s->box_h[0][1] = 0.5; s->box_h[0][2] = 0.4;
  s->box_yaw[0] = 0.5 + (seed ? (u01(mix64(seed + 99)) - 0.5) * 0.4 : 0.0);
}

/* exit distance of a ray that starts inside an axis-aligned box */
static double exit_aabb(const double* o, const double* d, const double* lo, const double* hi) {
  double s = INFINITY;
  for (int k = 0; k < 3; ++k) {
    if (d[k] > 0) { double t = (hi[k] - o[k]) / d[k]; if (t < s) s = t; }
    else if (d[k] < 0) { double t = (lo[k] - o[k]) / d[k]; if (t < s) s = t; }
  }
  return s;
}

/* entry distance of a ray into a box given in its own frame (slab test) */
static double enter_box(const double* o, const double* d, const double* h) {
  double t0 = 0.0, t1 = INFINITY;
  for (int k = 0; k < 3; ++k) {
    if (d[k] != 0.0) {
      double a = (-h[k] - o[k]) / d[k], b = (h[k] - o[k]) / d[k];
      if (a > b) { double t = a; a = b; b = t; }
      if (a > t0) t0 = a;
      if (b < t1) t1 = b;
    } else if (o[k] < -h[k] || o[k] > h[k]) {
      return INFINITY;
    }
  }
  return (t0 <= t1 && t0 > 0.0) ? t0 : INFINITY;
}

static double hit_sphere(const double* o, const double* d, const double* c, double r) {
  double oc[3] = {o[0] - c[0], o[1] - c[1], o[2] - c[2]};
  double a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  double b = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2];
  double cc = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - r * r;
  double disc = b * b - a * cc;
  if (disc <= 0) return INFINITY;
  double sq = sqrt(disc);
  double t = (-b - sq) / a;
  if (t > 1e-9) return t;
  return INFINITY;
}

int64_t rst_synth_render(const rst_synth_scene* sc, const double* T, double fx, double fy,
                         double cx, double cy, int32_t w, int32_t h, double depth_scale,
                         const rst_synth_noise* noise, uint64_t frame_seed, uint16_t* depth,
                         int32_t stride, uint8_t* rgb) {
  /* column-major 4x4: R(i,j) = T[i + 4*j], t = T[12..14] */
  const double o[3] = {T[12], T[13], T[14]};
  int64_t n_valid = 0;
  double cyaw[RST_SYNTH_MAX_BOXES], syaw[RST_SYNTH_MAX_BOXES];
  for (int b = 0; b < sc->n_boxes; ++b) { cyaw[b] = cos(sc->box_yaw[b]); syaw[b] = sin(sc->box_yaw[b]); }

#pragma omp parallel for schedule(static) reduction(+ : n_valid)
  for (int v = 0; v < h; ++v) {
    for (int u = 0; u < w; ++u) {
      const double xn = (u - cx) / fx, yn = (v - cy) / fy;
      double d[3];
      for (int i = 0; i < 3; ++i) d[i] = T[i] * xn + T[i + 4] * yn + T[i + 8];
      double s = exit_aabb(o, d, sc->room_lo, sc->room_hi);
      for (int k = 0; k < sc->n_spheres; ++k) {
        double t = hit_sphere(o, d, sc->sphere_c[k], sc->sphere_r[k]);
        if (t < s) s = t;
      }
      for (int b = 0; b < sc->n_boxes; ++b) {
        /* world -> box frame: rotate by -yaw about y */
        double ol[3] = {o[0] - sc->box_c[b][0], o[1] - sc->box_c[b][1], o[2] - sc->box_c[b][2]};
        double o2[3] = {cyaw[b] * ol[0] - syaw[b] * ol[2], ol[1], syaw[b] * ol[0] + cyaw[b] * ol[2]};
        double d2[3] = {cyaw[b] * d[0] - syaw[b] * d[2], d[1], syaw[b] * d[0] + cyaw[b] * d[2]};
        double t = enter_box(o2, d2, sc->box_h[b]);
        if (t < s) s = t;
      }
      double z = s; /* camera-frame direction is (xn, yn, 1): depth == ray parameter */
      const uint64_t pix = frame_seed * 0x100000001B3ull + (uint64_t)v * 65536ull + (uint64_t)u;
      int invalid = !(z > 0.0) || !isfinite(z);
      if (noise && !invalid) {
        if (noise->sigma_lsb_at_1m > 0) {
          double a = u01(mix64(pix * 3 + 1)), b = u01(mix64(pix * 3 + 2));
          double g = sqrt(-2.0 * log(a)) * cos(6.283185307179586 * b);
          z += g * noise->sigma_lsb_at_1m * depth_scale * z * z;
        }
        if (noise->p_invalid_pixel > 0 && u01(mix64(pix * 3)) < noise->p_invalid_pixel) invalid = 1;
        if (noise->p_invalid_block > 0) {
          uint64_t blk = frame_seed * 0x9E3779B1ull + (uint64_t)(v / 16) * 4096ull + (uint64_t)(u / 16) + 0x5bd1e995ull;
          if (u01(mix64(blk)) < noise->p_invalid_block) invalid = 1;
        }
      }
      double q = floor(z / depth_scale + 0.5);
      uint16_t dq = (!invalid && q >= 1.0 && q <= 65535.0) ? (uint16_t)q : 0;
      depth[(int64_t)v * stride + u] = dq;
      n_valid += dq != 0;
      if (rgb) {
        const double px = o[0] + s * d[0], py = o[1] + s * d[1], pz = o[2] + s * d[2];
        double r = 0.5 + 0.35 * sin(5.0 * px + 1.3 * pz) * cos(4.0 * py);
        double g = 0.5 + 0.35 * sin(3.0 * py - 2.0 * px) * cos(3.5 * pz);
        double bl = 0.5 + 0.35 * cos(4.5 * pz + 2.0 * py) * sin(2.5 * px + 0.7);
        uint8_t* p = rgb + ((int64_t)v * w + u) * 3;
        p[0] = (uint8_t)floor(r * 255.0 + 0.5);
        p[1] = (uint8_t)floor(g * 255.0 + 0.5);
        p[2] = (uint8_t)floor(bl * 255.0 + 0.5);
      }
    }
  }
  return n_valid;
}
