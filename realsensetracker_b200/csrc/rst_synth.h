/*
 * rst_synth.h — deterministic synthetic RGB-D frame source (CPU, C ABI).
 *
 * Stands in for the camera: the reference's only camera-free source is
 * `RandomSource` (rs_tracker/driver/include/rs_tracker/driver/data_source.hpp:22-41),
 * which emits uniform random clouds and cannot exercise a depth-image pipeline,
 * and its recorded `data/*.pb` clouds are not in the tree. This renderer
 * ray-casts an analytic room (box interior + spheres + oriented boxes, so all
 * six degrees of freedom are constrained) to CV_16UC1-style depth and CV_8UC3
 * colour at a given camera pose, so every test has a known rigid motion.
 *
 * Camera convention: x right, y down, z forward; depth = z of the hit point;
 * depth[v][u] = round(z / depth_scale), 0 = invalid / out of range.
 */
#ifndef RST_SYNTH_H_
#define RST_SYNTH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RST_SYNTH_MAX_SPHERES 8
#define RST_SYNTH_MAX_BOXES 4

typedef struct rst_synth_scene {
  double room_lo[3], room_hi[3];                 /* camera sits inside          */
  int32_t n_spheres;
  double sphere_c[RST_SYNTH_MAX_SPHERES][3];
  double sphere_r[RST_SYNTH_MAX_SPHERES];
  int32_t n_boxes;
  double box_c[RST_SYNTH_MAX_BOXES][3];          /* centre                      */
  double box_h[RST_SYNTH_MAX_BOXES][3];          /* half extents                */
  double box_yaw[RST_SYNTH_MAX_BOXES];           /* rotation about world y      */
} rst_synth_scene;

typedef struct rst_synth_noise {
  double sigma_lsb_at_1m;   /* gaussian depth noise (LSB) at 1 m, grows with z^2; 0 = off */
  double p_invalid_pixel;   /* Bernoulli probability of zeroing a pixel          */
  double p_invalid_block;   /* probability of zeroing a whole 16x16 block        */
} rst_synth_noise;

/* Room 6 x 3 x 5 m with 3 spheres + 1 box; `seed` jitters the object layout. */
void rst_synth_scene_default(uint64_t seed, rst_synth_scene* scene);

/* Renders one frame. T_wc: column-major 4x4 camera->world. depth: h rows of
 * depth_stride_px uint16. rgb: NULL or h*w*3 bytes. noise may be NULL.
 * Returns the number of valid depth pixels. Multi-threaded (OpenMP) over rows;
 * the result does not depend on the thread count. */
int64_t rst_synth_render(const rst_synth_scene* scene, const double* T_wc,
                         double fx, double fy, double cx, double cy,
                         int32_t w, int32_t h, double depth_scale,
                         const rst_synth_noise* noise, uint64_t frame_seed,
                         uint16_t* depth, int32_t depth_stride_px, uint8_t* rgb);

#ifdef __cplusplus
}
#endif
#endif
