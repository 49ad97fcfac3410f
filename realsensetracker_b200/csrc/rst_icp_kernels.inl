/*
 * rst_icp_kernels.inl — K3 + K4 + K5: projective association, point-to-plane residual / Jacobian, the 29-sum
 * reduction and the on-device 6x6 solve, as two kernels over ONE shared per-lane pixel pipeline (PixelPipe).
 * Compiled once per slice of the kernel variants (rst_icp_part*.cu, see the end of this file).
 *
 *   k_icp_iter  : the default schedule. One launch = one iteration of one level over a grid of
 *                 (blocks per pair) x (pairs): the block's source depth arrives by bulk async copies (UBLKCP +
 *                 mbarrier), block partials go to global memory, the last block of a pair (ticket) reduces them in
 *                 fp64 in a fixed order, solves and updates the pose. Also the single-evaluation entry point
 *                 (rst_evaluate: association index dump, no pose update). Small batches chain their launches with
 *                 programmatic dependent launch.
 *   k_icp_fused : one thread-block CLUSTER owns one frame pair and runs EVERY iteration of the chosen pyramid
 *                 levels inside a single launch: each CTA walks its share of the level's pixels (source depth
 *                 streamed through a double-buffered shared-memory tile), the CTAs' 29 partial sums are combined
 *                 in fixed rank order through distributed shared memory by the leader CTA, which solves (fp64
 *                 Cholesky), updates the fp64 master pose and hands the fp32 pose of the next iteration back
 *                 through DSMEM; two cluster barriers per iteration, no global partials, no tickets, no launch
 *                 per iteration. Selectable (rst_set_schedule: fused / hybrid); measured slower than one launch
 *                 per iteration at 128 pairs, so it is not the default.
 *
 * Nearest reference counterpart: the correspondence + weight loop, covariance accumulation and closed-form
 * solve of AlignIcp3d, align_icp.cpp:101-151. Arithmetic specification: DESIGN.md §3.
 */
#include <cooperative_groups.h>

#include <type_traits>

#include "rst_device.cuh"
#include "rst_kernels.cuh"

namespace cg = cooperative_groups;

namespace rst {

constexpr float kRintMagic = 12582912.0f;  // 1.5 * 2^23: x + magic rounds x to nearest-even integer

// transposed butterfly: after the 5 steps lane L holds the warp total of acc[L]; every total is
// formed by the same (xor 16, 8, 4, 2, 1) addition tree as a plain shuffle all-reduce.
template <int OFF>
__device__ __forceinline__ void butterfly_step(float (&acc)[kAccPad], bool upper) {
#pragma unroll
  for (int i = 0; i < OFF; ++i) {
    const float send = upper ? acc[i] : acc[i + OFF];
    const float keep = upper ? acc[i + OFF] : acc[i];
    acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}

// position of a warp-chunk in the image, walked incrementally (no division in the loop)
struct ChunkPos {
  int v, u;  // row, first column of this lane in the chunk (its two pixels are columns u and u + 32)
};

// shared-memory landing zone of one pipeline stage: [pixel of the stage][thread], filled by cp.async in K3, read in K4.
// Photometric variant: also the four bilinear taps of the destination intensity and the source intensity.
template <bool PHOTO>
struct StageBuf {
  float4 g[kPxPerStage][kIcpThreads];
};
template <>
struct StageBuf<true> {
  float4 g[kPxPerStage][kIcpThreads];
  float ph[5][kPxPerStage][kIcpThreads];   // I00, I10, I01, I11, I_src
};

// per-thread state of one pipeline stage: what K4 needs besides the gathered texel. A lane's two pixels of a
// chunk (columns u and u + 32) share one packed register pair (.x, .y).
template <bool NGATE, bool WRITE_IDX>
struct StageRegs {
  f2 qx[kChunksPerWarp], qy[kChunksPerWarp], qz[kChunksPerWarp];   // transformed source point p'
  float kxq[kPxPerStage], kyq[kPxPerStage];                        // (u'-cx)/fx, (v'-cy)/fy of the target pixel
  float rnx[NGATE ? kPxPerStage : 1], rny[NGATE ? kPxPerStage : 1], rnz[NGATE ? kPxPerStage : 1];  // R * n_src
  int tgt[WRITE_IDX ? kPxPerStage : 1], src[WRITE_IDX ? kPxPerStage : 1];                            // idx_out bookkeeping
};

// Gather offset of one source pixel: the target texel's byte offset when every K3 gate holds, else `guard` (an
// all-zero texel). The tests chain through the predicate input of ISETP/FSETP (one predicate, one select) instead
// of one select per test.
template <bool EXTRA>
__device__ __forceinline__ uint32_t gather_offset(uint32_t d_rel, uint32_t d_span, float qz, uint32_t ui, uint32_t w,
                                                  uint32_t vi, uint32_t h, uint32_t off, uint32_t guard, uint32_t extra_ok) {
  uint32_t r;
  if (EXTRA) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.le.u32 p, %1, %2;\n\t"
        "setp.ne.and.u32 p, %11, 0, p;\n\t"
        "setp.ge.and.f32 p, %3, %4, p;\n\t"
        "setp.lt.and.u32 p, %5, %6, p;\n\t"
        "setp.lt.and.u32 p, %7, %8, p;\n\t"
        "selp.u32 %0, %9, %10, p;\n\t}"
        : "=r"(r)
        : "r"(d_rel), "r"(d_span), "f"(qz), "f"(kMinProjZ), "r"(ui), "r"(w), "r"(vi), "r"(h), "r"(off), "r"(guard), "r"(extra_ok));
  } else {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.le.u32 p, %1, %2;\n\t"
        "setp.ge.and.f32 p, %3, %4, p;\n\t"
        "setp.lt.and.u32 p, %5, %6, p;\n\t"
        "setp.lt.and.u32 p, %7, %8, p;\n\t"
        "selp.u32 %0, %9, %10, p;\n\t}"
        : "=r"(r)
        : "r"(d_rel), "r"(d_span), "f"(qz), "f"(kMinProjZ), "r"(ui), "r"(w), "r"(vi), "r"(h), "r"(off), "r"(guard));
  }
  return r;
}

// The 29 sums of one thread as 12 packed pairs + 3 scalars + (sum w r^2, count): with x = (J0..J5) = (E0, E1, E2) in
// pairs, every product x_i x_c, x_i r is one half of E_a * E_b, E_a * swap(E_b) or E_a * (r, r); only the three
// in-pair cross terms stay scalar: 13 FFMA2 + 3 FFMA per pixel instead of 28 FFMA + the count add. Every half is the
// same IEEE fma as the scalar form, so the sums are bit-identical to a scalar accumulation in pixel order.
struct Accum29 {
  f2 P00, P11, P22, P01, P01s, P02, P02s, P12, P12s, B0, B1, B2, RC;  // RC = (sum w r^2, accepted pixels: exact in fp32)
  float a01, a23, a45;
  __device__ __forceinline__ void clear() {
    const f2 z = mk2(0.f, 0.f);
    P00 = P11 = P22 = P01 = P01s = P02 = P02s = P12 = P12s = B0 = B1 = B2 = RC = z;
    a01 = a23 = a45 = 0.f;
  }
  __device__ __forceinline__ void add(float j0, float j1, float j2, float j3, float j4, float j5, float r, float okf) {
    const f2 E0 = mk2(j0, j1), E1 = mk2(j2, j3), E2 = mk2(j4, j5), rr = bc2(r);
    P00 = fma2(E0, E0, P00); P11 = fma2(E1, E1, P11); P22 = fma2(E2, E2, P22);
    P01 = fma2(E0, E1, P01); P01s = fma2(E0, swp2(E1), P01s);
    P02 = fma2(E0, E2, P02); P02s = fma2(E0, swp2(E2), P02s);
    P12 = fma2(E1, E2, P12); P12s = fma2(E1, swp2(E2), P12s);
    B0 = fma2(E0, rr, B0); B1 = fma2(E1, rr, B1); B2 = fma2(E2, rr, B2);
    a01 = ffma(j0, j1, a01); a23 = ffma(j2, j3, a23); a45 = ffma(j4, j5, a45);
    const f2 ro = mk2(r, okf);
    RC = fma2(ro, ro, RC);   // okf is 0 or 1, so okf * okf counts the pixel
  }
  // canonical order: A upper triangle row-major (0,0)..(5,5), b, sum w r^2, count
  __device__ __forceinline__ void store(float (&acc)[kAccPad]) const {
    const float v[kAccPad] = {P00.x, a01, P01.x, P01s.x, P02.x, P02s.x, P00.y, P01s.y, P01.y, P02s.y, P02.y,
                              P11.x, a23, P12.x, P12s.x, P11.y, P12s.y, P12.y, P22.x, a45, P22.y,
                              B0.x, B0.y, B1.x, B1.y, B2.x, B2.y, RC.x, RC.y, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kAccPad; ++k) acc[k] = v[k];
  }
};

// constants of one (level, pair) every lane needs
struct PipeLevel {
  int W, H, row_span, group_dv, group_du;
  float fx, fy, cx, cy, ifx, ify;
  uint32_t guard, w16;   // byte offsets: the all-zero texel behind the frame (w * h * 16), one texel row (w * 16)
  const float4* Gs;      // source geometry map (normal gate only)
  const float4* Gd;      // destination geometry map
  const float* Is;       // intensity maps (photometric term only)
  const float* Id;
};

struct PipeParams {
  float depth_scale, dmax2, ncos_min, robust_scale, sqrt_lambda;
  uint32_t d_lo, d_span;   // valid raw depth: (d - d_lo) <= d_span
};

// ----------------------------------------------------------------------------------
// The per-lane pixel pipeline. A warp covers kChunksPerWarp chunks of 64 px per group (each lane 2 pixels per
// chunk, columns u and u + 32: the 32 lanes of one gather instruction cover 32 CONSECUTIVE source pixels, whose
// target texels are nearly always 512 contiguous bytes = 4 cache lines) and walks `n` consecutive groups:
//   K3  transform + project, start the 16-byte cp.async gathers of the target texels into shared memory
//       (fp32 arithmetic packed over the lane's two pixels),
//   K4  one stage later: gates, residual, Jacobian, branch-free accumulation into the 29 sums.
// Two stages are always in flight (2 * kPxPerStage gathers per thread), at no register cost.
// ----------------------------------------------------------------------------------
template <int ROBUST, bool NGATE, bool WRITE_IDX, bool PHOTO>
struct PixelPipe {
  float R00, R01, R02, R10, R11, R12, R20, R21, R22, tx, ty, tz;  // pose of this iteration, src -> dst
  PipeLevel L;
  PipeParams P;
  int32_t* idx_out;           // WRITE_IDX: this pair's w*h index map
  Accum29 acc;
  ChunkPos pos;               // chunk of this warp in the NEXT K3 group
  const uint32_t (*sd)[32];   // staged source depth: chunk-major, 64 uint16 per chunk (current tile)
  int cl;                     // tile-local chunk index of this warp in the next K3 group
  int tid, lane;

  __device__ __forceinline__ void set_pose(const float* p) {
    R00 = p[0]; R01 = p[1]; R02 = p[2]; R10 = p[3]; R11 = p[4]; R12 = p[5];
    R20 = p[6]; R21 = p[7]; R22 = p[8]; tx = p[9]; ty = p[10]; tz = p[11];
  }
  __device__ __forceinline__ void next_chunk(ChunkPos& p) const {   // +1 chunk
    p.u += kChunkPx;
    if (p.u >= L.row_span) { p.u -= L.row_span; p.v += 1; }
  }
  __device__ __forceinline__ void next_group(ChunkPos& p) const {   // +kChunksPerBlock chunks (host-precomputed strides)
    p.v += L.group_dv; p.u += L.group_du;
    if (p.u >= L.row_span) { p.u -= L.row_span; p.v += 1; }
  }

  // ---- K3: transform + project the group at `pos`, start the gathers into stage buffer `sbuf`
  __device__ __forceinline__ void k3(StageRegs<NGATE, WRITE_IDX>& st, StageBuf<PHOTO>& sb) {
    const int W = L.W, H = L.H;
    const float cx = L.cx, cy = L.cy, ifx = L.ifx, ify = L.ify;
    ChunkPos p = pos;
#pragma unroll
    for (int k = 0; k < kChunksPerWarp; ++k) {
      const float ky = fmul(fsub((float)p.v, cy), ify);
      const float fu0 = (float)p.u;
      const f2 kx = mul2(add2(mk2(fu0, fu0 + 32.0f), bc2(-cx)), bc2(ifx));   // (float(u) - cx) * ifx, columns u and u + 32
      const uint16_t* sd16 = reinterpret_cast<const uint16_t*>(&sd[cl + k][0]);
      const uint32_t d0 = sd16[lane], d1 = sd16[32 + lane];
      // exact uint16 -> float without the conversion unit: 2^23 + d, minus 2^23
      const f2 z = mul2(add2(mk2(__int_as_float(0x4B000000u | d0), __int_as_float(0x4B000000u | d1)), bc2(-8388608.0f)),
                        bc2(P.depth_scale));
      uint32_t extra[2] = {1u, 1u};
      if (NGATE) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int e = 2 * k + j;
          const bool okd = ((j ? d1 : d0) - P.d_lo) <= P.d_span;
          float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
          if (okd) gs = __ldg(L.Gs + (uint32_t)(p.v * W + p.u + 32 * j));
          extra[j] = gs.w > 0.0f ? 1u : 0u;
          st.rnx[e] = ffma(R00, gs.x, ffma(R01, gs.y, fmul(R02, gs.z)));
          st.rny[e] = ffma(R10, gs.x, ffma(R11, gs.y, fmul(R12, gs.z)));
          st.rnz[e] = ffma(R20, gs.x, ffma(R21, gs.y, fmul(R22, gs.z)));
        }
      }
      const f2 px = mul2(kx, z), py = mul2(bc2(ky), z);
      const f2 qx = fma2(bc2(R00), px, fma2(bc2(R01), py, fma2(bc2(R02), z, bc2(tx))));
      const f2 qy = fma2(bc2(R10), px, fma2(bc2(R11), py, fma2(bc2(R12), z, bc2(ty))));
      const f2 qz = fma2(bc2(R20), px, fma2(bc2(R21), py, fma2(bc2(R22), z, bc2(tz))));
      // the clamp only matters for rejected pixels: it keeps 1/qz, u_f, v_f finite so that they can
      // flow through the branch-free arithmetic below. 1/x = rcp_rn_normal(x), packed.
      const f2 qzc = mk2(fmaxf(qz.x, kMinProjZ), fmaxf(qz.y, kMinProjZ));
      f2 y0;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0.x) : "f"(qzc.x));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0.y) : "f"(qzc.y));
      const f2 iz = fma2(y0, fma2(neg2(qzc), y0, bc2(1.0f)), y0);
      const f2 uf = fma2(bc2(L.fx), mul2(qx, iz), bc2(cx));
      const f2 vf = fma2(bc2(L.fy), mul2(qy, iz), bc2(cy));
      // round-half-even without F2I/I2F: x + 1.5*2^23 holds rint(x) in its low mantissa bits for
      // |x| < 2^22; anything else (including huge values) maps outside [0, w) as an unsigned integer,
      // so one unsigned compare per axis is the complete "rint(u_f) in [0, w-1]" test.
      const f2 um = add2(uf, bc2(kRintMagic)), vm = add2(vf, bc2(kRintMagic));
      st.qx[k] = qx; st.qy[k] = qy; st.qz[k] = qz;
      {  // kx(u') = (float(u') - cx) * ifx with float(u') = um - magic (exact)
        const f2 kxq = mul2(add2(add2(um, bc2(-kRintMagic)), bc2(-cx)), bc2(ifx));
        const f2 kyq = mul2(add2(add2(vm, bc2(-kRintMagic)), bc2(-cy)), bc2(ify));
        st.kxq[2 * k] = kxq.x; st.kxq[2 * k + 1] = kxq.y; st.kyq[2 * k] = kyq.x; st.kyq[2 * k + 1] = kyq.y;
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int e = 2 * k + j;
        const uint32_t ui = (uint32_t)(__float_as_int(j ? um.y : um.x) - 0x4B400000);
        const uint32_t vi = (uint32_t)(__float_as_int(j ? vm.y : vm.x) - 0x4B400000);
        // d != 0 && z_min <= z <= z_max (bounds precomputed on the host) && qz >= kMinProjZ && u', v' inside the
        // image; rejected pixels gather the all-zero guard texel behind the frame, which fails the gz > 0 gate of
        // K4. Byte offset inside the frame (32-bit: a frame holds < 2^27 texels) added to the 64-bit base.
        const uint32_t off = gather_offset<NGATE>((j ? d1 : d0) - P.d_lo, P.d_span, j ? qz.y : qz.x, ui, (uint32_t)W, vi,
                                                  (uint32_t)H, vi * L.w16 + (ui << 4), L.guard, extra[j]);
        if (WRITE_IDX) {
          st.tgt[e] = off == L.guard ? -1 : (int)(off >> 4);
          st.src[e] = (p.v < H && p.u + 32 * j < W) ? p.v * W + p.u + 32 * j : -1;
        }
        cp_async_16(&sb.g[e][tid], reinterpret_cast<const char*>(L.Gd) + off);
        if constexpr (PHOTO) {
          // the five intensity values the photometric row of K4 reads ride in the same commit group: bilinear taps of
          // the destination at floor(u_f), floor(v_f) clamped to the image, and the source pixel's own intensity
          const float ufj = j ? uf.y : uf.x, vfj = j ? vf.y : vf.x;
          const int xi = (int)floorf(ufj), yi = (int)floorf(vfj);
          const int x0 = min(max(xi, 0), W - 1), x1 = min(max(xi + 1, 0), W - 1);
          const int y0 = min(max(yi, 0), H - 1), y1 = min(max(yi + 1, 0), H - 1);
          cp_async_4(&sb.ph[0][e][tid], L.Id + y0 * W + x0);
          cp_async_4(&sb.ph[1][e][tid], L.Id + y0 * W + x1);
          cp_async_4(&sb.ph[2][e][tid], L.Id + y1 * W + x0);
          cp_async_4(&sb.ph[3][e][tid], L.Id + y1 * W + x1);
          cp_async_4(&sb.ph[4][e][tid], L.Is + min(p.v * W + p.u + 32 * j, W * H - 1));
        }
      }
      if (k + 1 < kChunksPerWarp) next_chunk(p);
    }
    cp_async_commit();
    next_group(pos);
    cl += kChunksPerBlock;
  }

  // ---- K4: gates, residual, Jacobian, branch-free accumulation of one landed stage
  __device__ __forceinline__ void k4(const StageRegs<NGATE, WRITE_IDX>& st, const StageBuf<PHOTO>& sb) {
#pragma unroll
    for (int e = 0; e < kPxPerStage; ++e) {
      const int k = e >> 1;
      const float qx = (e & 1) ? st.qx[k].y : st.qx[k].x, qy = (e & 1) ? st.qy[k].y : st.qy[k].x;
      const float qz = (e & 1) ? st.qz[k].y : st.qz[k].x;
      const float4 g = sb.g[e][tid];
      const float gz = g.w;
      const float dx = ffma(-st.kxq[e], gz, qx);   // p' - q, q = (kx(u') gz, ky(v') gz, gz) recomputed from the texel's z
      const float dy = ffma(-st.kyq[e], gz, qy);
      const float dz = fsub(qz, gz);
      const float dist2 = ffma(dz, dz, ffma(dy, dy, fmul(dx, dx)));
      bool ok = (gz > 0.0f) && (dist2 <= P.dmax2);   // gz == 0: invalid texel or rejected in K3 (guard texel)
      if (NGATE) {
        const float cs = ffma(st.rnz[e], g.z, ffma(st.rny[e], g.y, fmul(st.rnx[e], g.x)));
        ok = ok && (cs >= P.ncos_min);
      }
      if (WRITE_IDX) {
        if (st.src[e] >= 0) idx_out[st.src[e]] = ok ? st.tgt[e] : -1;
      }
      // rejected pixels contribute exact zeros: their normal is masked to 0 (so r = 0 and J = 0) and
      // every other operand is finite (q from finite inputs, texel = real map data or the zero guard)
      float nx = ok ? g.x : 0.0f, ny = ok ? g.y : 0.0f, nz = ok ? g.z : 0.0f;
      float r = ffma(nz, dz, ffma(ny, dy, fmul(nx, dx)));
      if (ROBUST != RST_ROBUST_NONE) {
        // A = sum (sqrt(w) J)(sqrt(w) J)^T: scale the normal (hence J and r) by sqrt(w) once
        float wgt;
        // IEEE division and square root by their branch-free normal-range sequences: the operands are bounded
        // (|r| <= dist_max, scale > 0), the weight lies in (0, 1]
        if (ROBUST == RST_ROBUST_HUBER) {
          const float ar = fabsf(r);
          wgt = ar <= P.robust_scale ? 1.0f : div_rn_normal(P.robust_scale, ar);
        } else {
          const float t = div_rn_normal(P.robust_scale, ffma(r, r, P.robust_scale));
          wgt = fmul(t, t);
        }
        const float sw = sqrt_rn_normal(wgt);
        nx = fmul(sw, nx); ny = fmul(sw, ny); nz = fmul(sw, nz); r = fmul(sw, r);
      }
      acc.add(ffma(qy, nz, -fmul(qz, ny)), ffma(qz, nx, -fmul(qx, nz)), ffma(qx, ny, -fmul(qy, nx)), nx, ny, nz, r,
              ok ? 1.0f : 0.0f);
      if constexpr (PHOTO) {
        // photometric row (f2): r_I = I_dst(pi(p')) - I_src(u,v) by bilinear sampling clamped at the borders,
        // J_I = [p' x d ; d], d = (dI/du fx/z, dI/dv fy/z, -(d_x x + d_y y)/z), all scaled by sqrt(lambda).
        // The taps were fetched by K3 at the same floor(u_f), floor(v_f) (same expressions, bit for bit).
        const float fx = L.fx, fy = L.fy;
        const float iz = rcp_rn_normal(fmaxf(qz, kMinProjZ));
        const float uf = ffma(fx, fmul(qx, iz), L.cx), vf = ffma(fy, fmul(qy, iz), L.cy);
        const float axf = ok ? fsub(uf, floorf(uf)) : 0.0f, ayf = ok ? fsub(vf, floorf(vf)) : 0.0f;
        const float I00 = sb.ph[0][e][tid], I10 = sb.ph[1][e][tid], I01 = sb.ph[2][e][tid], I11 = sb.ph[3][e][tid];
        const float Isrc = sb.ph[4][e][tid];
        const float dt = fsub(I10, I00), db = fsub(I11, I01);
        const float top = ffma(axf, dt, I00), bot = ffma(axf, db, I01);
        const float gv = fsub(bot, top);
        const float val = ffma(ayf, gv, top);
        const float gu = ffma(ayf, fsub(db, dt), dt);
        const float sl = ok ? P.sqrt_lambda : 0.0f;    // rejected pixels contribute exact zeros
        const float rI = fmul(sl, fsub(val, Isrc));
        const float da = fmul(sl, fmul(fmul(gu, fx), iz)), dbv = fmul(sl, fmul(fmul(gv, fy), iz));
        const float dc = -fmul(ffma(da, qx, fmul(dbv, qy)), iz);
        acc.add(ffma(qy, dc, -fmul(qz, dbv)), ffma(qz, da, -fmul(qx, dc)), ffma(qx, dbv, -fmul(qy, da)), da, dbv, dc, rI, 0.0f);
      }
    }
  }

  // ---- two-stage software pipeline over `n` consecutive groups: K3(g+1) is issued before K4(g) consumes its
  //      texels. `tile(g)` is called in front of K3 of every group g > 0 with g % TILE_GROUPS == 0 (the caller
  //      switches `sd` / `cl` to the next staged depth tile there); the gather pipeline is NOT drained at a tile
  //      boundary.
  template <int TILE_GROUPS, class TileFn>   // TILE_GROUPS: even, or 0 = one tile holds everything
  __device__ __forceinline__ void run(int n, StageBuf<PHOTO>* s_g, TileFn&& tile) {
    static_assert(TILE_GROUPS % 2 == 0, "tile boundaries must fall on even groups");
    if (n <= 0) return;
    StageRegs<NGATE, WRITE_IDX> st0, st1;
    k3(st0, s_g[0]);
#pragma unroll 1
    for (int gi = 0; gi < n; gi += 2) {
      if (gi + 1 < n) {
        k3(st1, s_g[1]);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      k4(st0, s_g[0]);
      if (gi + 1 < n) {
        if (gi + 2 < n) {
          if (TILE_GROUPS > 0 && (gi + 2) % (TILE_GROUPS > 0 ? TILE_GROUPS : 2) == 0) tile(gi + 2);
          k3(st0, s_g[0]);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        k4(st1, s_g[1]);
      }
    }
  }
};

// Stages `n_chunks` chunks starting at level chunk index `c_base` into `dst` ([chunk][32 words]): every chunk is
// 64 px = 8 pieces of 16 B; pieces beyond the row end / image end are zero-filled (= invalid depth) by the
// src-size form of cp.async. One commit group.
__device__ __forceinline__ void stage_depth(uint32_t (*dst)[32], const uint16_t* __restrict__ Ds, int depth_pitch, int W, int H,
                                            int chunks_per_row, uint32_t cpr_magic, int c_base, int n_chunks, int tid) {
  for (int q = tid; q < n_chunks * 8; q += kIcpThreads) {
    const int cl = q >> 3, piece = q & 7;
    const int c = c_base + cl;
    const int v = chunks_per_row == 1 ? c : (int)__umulhi((uint32_t)c, cpr_magic);  // c / chunks_per_row
    const int u = (c - v * chunks_per_row) * kChunkPx + piece * 8;
    int npx = v < H ? W - u : 0;
    npx = npx < 0 ? 0 : (npx > 8 ? 8 : npx);
    const uint16_t* src = npx > 0 ? Ds + (uint32_t)(v * depth_pitch + u) : Ds;
    cp_async_16_zfill(&dst[cl][piece * 4], src, npx * 2);
  }
  cp_async_commit();
}

// Bulk-copy form of stage_depth (k_icp_iter): the tile is a run of consecutive chunks, i.e. a run of row segments, and
// a row segment is contiguous both in global memory (up to the row pitch) and in the chunk-major tile. Warp 0 issues
// ONE cp.async.bulk per row segment (one per lane, a single copy for the whole tile when rows are dense: w a multiple
// of 64) that completes on `bar`; no per-16-byte address arithmetic, no LDGSTS issue slots. What a copy cannot deliver
// is zero-filled with plain stores: the columns between the pitch and the end of a row's last chunk, and chunks below
// the image. Every thread of the block calls this; after mbar_wait(bar, parity) + stage_depth_bulk_fix + a block
// barrier the tile is complete.
struct BulkTile {
  int v0, n_rows, c_base, c_end;
};
__device__ __forceinline__ BulkTile stage_depth_bulk(uint32_t (*dst)[32], const uint16_t* __restrict__ Ds, int pitch, int H,
                                                     int cpr, uint32_t cpr_magic, int c_base, int n_chunks, uint64_t* bar, int tid) {
  BulkTile t;
  t.c_base = c_base; t.c_end = c_base + n_chunks;
  t.v0 = cpr == 1 ? c_base : (int)__umulhi((uint32_t)c_base, cpr_magic);
  const int v_last = cpr == 1 ? t.c_end - 1 : (int)__umulhi((uint32_t)(t.c_end - 1), cpr_magic);
  t.n_rows = v_last - t.v0 + 1;
  uint16_t* d16 = reinterpret_cast<uint16_t*>(&dst[0][0]);
  const int row_px = cpr * kChunkPx;
  if (tid < 32) {
    if (pitch == row_px) {   // dense rows: the whole tile is one contiguous run, clipped at the image end
      if (tid == 0) {
        const int c_img = H * cpr;
        const int nc = min(t.c_end, c_img) - c_base;
        const uint32_t bytes = nc > 0 ? (uint32_t)nc * (kChunkPx * 2) : 0u;
        mbar_arrive_expect_tx(bar, bytes);
        if (bytes) bulk_g2s(d16, Ds + (uint32_t)c_base * kChunkPx, bytes, bar);
      }
    } else {
      uint32_t total = 0;
      for (int r = tid; r < t.n_rows; r += 32) {
        const int v = t.v0 + r, cr = v * cpr;
        const int ca = max(c_base, cr) - cr, cb = min(t.c_end, cr + cpr) - cr;
        const int px = min(cb * kChunkPx, pitch) - ca * kChunkPx;
        if (v < H) total += (uint32_t)px * 2;
      }
      total = __reduce_add_sync(0xffffffffu, total);
      if (tid == 0) mbar_arrive_expect_tx(bar, total);
      __syncwarp();
      for (int r = tid; r < t.n_rows; r += 32) {
        const int v = t.v0 + r, cr = v * cpr;
        const int ca = max(c_base, cr) - cr, cb = min(t.c_end, cr + cpr) - cr;
        const int px = min(cb * kChunkPx, pitch) - ca * kChunkPx;
        if (v < H) bulk_g2s(d16 + (cr + ca - c_base) * kChunkPx, Ds + (uint32_t)(v * pitch + ca * kChunkPx), (uint32_t)px * 2, bar);
      }
    }
  }
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  if (pitch < row_px) {   // columns [pitch, row_px) of every row that ends inside the tile (block-uniform branch)
    const int ppr = (row_px - pitch) >> 3;   // 16-byte pieces, 1..7
    for (int r = tid >> 3; r < t.n_rows; r += kIcpThreads >> 3) {
      const int v = t.v0 + r, cr = v * cpr, piece = tid & 7;
      if (piece < ppr && v < H && cr + cpr <= t.c_end)
        *reinterpret_cast<uint4*>(d16 + (cr - c_base) * kChunkPx + pitch + piece * 8) = z4;
    }
  }
  {  // chunks below the image (last block of a pair only)
    const int c_zero = max(c_base, H * cpr);
    for (int q = (c_zero - c_base) * 8 + tid; q < n_chunks * 8; q += kIcpThreads)
      *reinterpret_cast<uint4*>(d16 + q * 8) = z4;
  }
  return t;
}
// After the copies have landed: columns [w, pitch) hold whatever the caller's row padding holds (frames bound in place
// with rst_set_frames_device); they count as invalid depth.
__device__ __forceinline__ void stage_depth_bulk_fix(uint32_t (*dst)[32], const BulkTile& t, int pitch, int W, int H, int cpr, int tid) {
  if (W >= pitch) return;
  uint16_t* d16 = reinterpret_cast<uint16_t*>(&dst[0][0]);
  for (int r = tid; r < t.n_rows; r += kIcpThreads) {
    const int v = t.v0 + r, cr = v * cpr;
    if (v < H && cr + cpr <= t.c_end)
      for (int u = W; u < pitch; ++u) d16[(cr - t.c_base) * kChunkPx + u] = 0;
  }
}

// K5 stage 1 inside a block: fixed-shape warp tree (xor 16,8,4,2,1, transposed), then the warps in index order.
// Returns, for tid < kAcc, the block's sum of column tid (other threads: 0). Contains one __syncthreads.
__device__ __forceinline__ float block_reduce29(const Accum29& a29, float (*s_warp)[kAccPad], int tid, int lane, int warp) {
  float acc[kAccPad];
  a29.store(acc);
  butterfly_step<16>(acc, (lane & 16) != 0);
  butterfly_step<8>(acc, (lane & 8) != 0);
  butterfly_step<4>(acc, (lane & 4) != 0);
  butterfly_step<2>(acc, (lane & 2) != 0);
  butterfly_step<1>(acc, (lane & 1) != 0);
  s_warp[warp][lane] = acc[0];
  __syncthreads();
  float s = 0.f;
  if (tid < kAcc) {
    s = s_warp[0][tid];
#pragma unroll
    for (int w = 1; w < kIcpThreads / 32; ++w) s += s_warp[w][tid];
  }
  return s;
}

// K5 stage 2 on one thread: solve, update the fp64 master pose, report. `Rt` = master pose (row-major R, t).
// Returns true when the pose was updated; *converged is set when the update is below converge_eps.
struct SolveReport {
  int status;        // of this iteration
  bool updated, converged;
};
__device__ inline SolveReport solve_and_update(const double* tot, int min_count, float damping, float converge_eps, bool update_pose,
                                               double* Rt, rst_stats* st) {
  double A[21], b[6], xi[6];
#pragma unroll
  for (int k = 0; k < 21; ++k) A[k] = tot[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) b[k] = tot[21 + k];
  const double swr2 = tot[27];
  const int count = (int)tot[28];
  SolveReport rep{0, false, false};
  const int rc = solve6(A, b, count, min_count, (double)damping, xi);
  if (update_pose) {
    if (rc == RST_STATUS_OK) {
      double Rn[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) Rn[k] = Rt[k];
      se3_update(xi, Rn);
      bool fin = true;
#pragma unroll
      for (int k = 0; k < 12; ++k) fin &= isfinite(Rn[k]);
      if (fin) {
#pragma unroll
        for (int k = 0; k < 12; ++k) Rt[k] = Rn[k];
        rep.updated = true;
        if (converge_eps > 0.f) {
          const double wn = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
          const double vn = sqrt(xi[3] * xi[3] + xi[4] * xi[4] + xi[5] * xi[5]);
          rep.converged = wn < (double)converge_eps && vn < (double)converge_eps;
        }
      } else {
        rep.status |= RST_STATUS_NON_FINITE;
      }
    } else {
      rep.status |= rc;
    }
    st->iterations += 1;
  } else {
    rep.status |= rc;
  }
  st->status = rep.status;
  if (rep.status != RST_STATUS_OK) { st->any_status |= rep.status; if (update_pose) st->failed_iterations += 1; }
  st->count = count;
  st->sum_wr2 = swr2;
  st->rmse = count > 0 ? (float)sqrt(swr2 / (double)count) : 0.f;
#pragma unroll
  for (int k = 0; k < 21; ++k) st->A[k] = A[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) st->b[k] = b[k];
  return rep;
}

__device__ __forceinline__ void write_pose_outputs(const double* Rt, float* f12, float* cm16) {
#pragma unroll
  for (int k = 0; k < 12; ++k) f12[k] = (float)Rt[k];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) cm16[r + 4 * c] = (float)Rt[3 * r + c];
    cm16[12 + r] = (float)Rt[9 + r];
    cm16[4 * r + 3] = 0.f;
  }
  cm16[15] = 1.f;
}

// ----------------------------------------------------------------------------------
// k_icp_iter: one iteration of one level. grid (blocks_per_pair, n_pairs), kIcpThreads threads.
// ----------------------------------------------------------------------------------
template <int ROBUST, bool NGATE, bool WRITE_IDX, bool PHOTO, bool EARLY>
__global__ void __launch_bounds__(kIcpThreads, (PHOTO || ROBUST != RST_ROBUST_NONE || NGATE || WRITE_IDX) ? 4 : RST_ICP_MINB)
k_icp_iter(const __grid_constant__ IcpArgs a) {
  // gathered destination texels land here through cp.async: [stage][pixel][thread], 16 B each, so the
  // two-deep gather pipeline costs no registers and every LDS.128 is conflict-free
  // the two stage buffers live in dynamic shared memory: the photometric variants exceed the 48 KB static limit
  extern __shared__ __align__(16) unsigned char s_dyn[];
  StageBuf<PHOTO>* s_g = reinterpret_cast<StageBuf<PHOTO>*>(s_dyn);
  // source depth of the whole block (<= 8192 px), staged once with 16-byte zero-filling cp.async
  __shared__ __align__(16) uint32_t s_d[kMaxGroups * kChunksPerBlock][32];
  __shared__ float s_warp[kIcpThreads / 32][kAccPad];
  __shared__ double s_tot[kAccPad];
  __shared__ __align__(8) uint64_t s_bar;   // completion of the bulk copies of the depth tile
  __shared__ int s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = a.pair_offset + blockIdx.y;
  if (tid == 0) mbar_init(&s_bar, 1);       // warp 0 arms it below (same warp: program order), the others wait after the barrier
  if (EARLY && !a.pdl) {  // convergence test on: this pair may have left the level already (block-uniform)
    if (a.done[pair]) return;
  }
  const int2 slots = a.pairs[pair];
  const int W = a.g.w, H = a.g.h;
  const uint16_t* __restrict__ Ds = a.lv.depth + (int64_t)slots.x * a.lv.depth_frame;

  PixelPipe<ROBUST, NGATE, WRITE_IDX, PHOTO> pipe;
  pipe.tid = tid; pipe.lane = lane;
  pipe.L.W = W; pipe.L.H = H; pipe.L.row_span = a.chunks_per_row * kChunkPx;
  pipe.L.group_dv = a.group_dv; pipe.L.group_du = a.group_du;
  pipe.L.fx = a.g.fx; pipe.L.fy = a.g.fy; pipe.L.cx = a.g.cx; pipe.L.cy = a.g.cy; pipe.L.ifx = a.g.ifx; pipe.L.ify = a.g.ify;
  pipe.L.guard = a.guard_texel << 4; pipe.L.w16 = (uint32_t)W << 4;
  pipe.L.Gs = a.lv.geom + (int64_t)slots.x * a.lv.geom_frame;
  pipe.L.Gd = a.lv.geom + (int64_t)slots.y * a.lv.geom_frame;  // [w*h] = all-zero guard texel
  pipe.L.Is = PHOTO ? a.lv.intensity + (int64_t)slots.x * a.lv.int_frame : nullptr;
  pipe.L.Id = PHOTO ? a.lv.intensity + (int64_t)slots.y * a.lv.int_frame : nullptr;
  pipe.P.depth_scale = a.depth_scale; pipe.P.dmax2 = a.dmax2; pipe.P.ncos_min = a.ncos_min;
  pipe.P.robust_scale = a.robust_scale; pipe.P.sqrt_lambda = a.sqrt_lambda; pipe.P.d_lo = a.d_lo; pipe.P.d_span = a.d_span;
  pipe.idx_out = WRITE_IDX ? a.idx_out + (int64_t)blockIdx.y * W * H : nullptr;
  // Programmatic dependent launch (small batches, the latency path): the next iteration's launch may start while this
  // one is still in its reduction / solve tail, and runs its pose-independent head (parameter setup, depth staging)
  // under it; everything the previous launch writes (pose, done flag, tickets) is read after griddepcontrol.wait.
  if (a.pdl == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (!a.pdl) pipe.set_pose(a.pose_f32 + 12 * pair);
  pipe.acc.clear();
  {  // first chunk of this warp in group 0; later chunks/groups are reached by adding strides
    const int c = blockIdx.x * a.groups * kChunksPerBlock + warp * kChunksPerWarp;
    const int v = a.chunks_per_row == 1 ? c : (int)__umulhi((uint32_t)c, a.cpr_magic);  // c / chunks_per_row
    pipe.pos.v = v; pipe.pos.u = (c - v * a.chunks_per_row) * kChunkPx + lane;
  }
  // Depth tile: bulk copies (one per row segment, a single one when rows are dense) unless the rows are short
  // (coarse levels of ragged sizes), where 16-byte cp.async spread over all threads has less latency than a train of
  // small bulk copies. Block-uniform choice.
  const int c_base = blockIdx.x * a.groups * kChunksPerBlock;
  if (a.lv.depth_pitch == a.chunks_per_row * kChunkPx || a.lv.depth_pitch >= 256) {
    __syncwarp();
    const BulkTile bt = stage_depth_bulk(s_d, Ds, a.lv.depth_pitch, H, a.chunks_per_row, a.cpr_magic, c_base,
                                         a.groups * kChunksPerBlock, &s_bar, tid);
    if (a.pdl) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      pipe.set_pose(a.pose_f32 + 12 * pair);
    }
    __syncthreads();          // s_bar initialised for everyone; zero-fill stores visible
    mbar_wait(&s_bar, 0);     // the tile has landed (nobody may leave before: the copies target this block's shared memory)
    if (W < a.lv.depth_pitch) {
      stage_depth_bulk_fix(s_d, bt, a.lv.depth_pitch, W, H, a.chunks_per_row, tid);
      __syncthreads();
    }
  } else {
    stage_depth(s_d, Ds, a.lv.depth_pitch, W, H, a.chunks_per_row, a.cpr_magic, c_base, a.groups * kChunksPerBlock, tid);
    if (a.pdl) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      pipe.set_pose(a.pose_f32 + 12 * pair);
    }
    cp_async_wait<0>();
    __syncthreads();
  }
  if (EARLY && a.pdl) {
    if (a.done[pair]) return;
  }
  pipe.sd = s_d; pipe.cl = warp * kChunksPerWarp;
  pipe.template run<0>(a.groups, s_g, [](int) {});
  // pdl == 2 (large batches): the next launch may start only now, when this block has left its pixel loop — its
  // blocks then come up while the last blocks of this launch reduce and solve, instead of competing for SM slots
  // with blocks that still have pixels to process
  if (a.pdl == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // ---- K5 stage 1
  const float bs = block_reduce29(pipe.acc, s_warp, tid, lane, warp);
  float* __restrict__ part = a.partials + ((int64_t)pair * a.max_blocks + blockIdx.x) * kAccPad;
  if (tid < kAcc) __stcg(part + tid, bs);

  // ---- K5 stage 2: the last block of this pair reduces the partials and solves
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const uint32_t t = atomicAdd(a.tickets + pair, 1u);
    s_last = (t == (uint32_t)(a.blocks_per_pair - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // Column c of the block partials is summed by 8 lanes (lane `sub` takes blocks sub, sub + 8, ... in index order),
  // then a 4/2/1 tree over the lanes: that order is part of the specification. The first 64 threads each own four
  // adjacent columns and fetch them as ONE float4 per block, five blocks in flight: the 150 partial rows of a pair in
  // the latency tiling cost 4 L2 round trips instead of 10 x 2.
  if (tid < 64) {   // two whole warps: the shuffles below are warp-uniform
    const int sub = tid & 7, cg4 = (tid >> 3) * 4;
    const float4* __restrict__ base = reinterpret_cast<const float4*>(a.partials + (int64_t)pair * a.max_blocks * kAccPad + cg4);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int b0 = sub; b0 < a.blocks_per_pair; b0 += 40) {
      const int b1 = b0 + 8, b2 = b0 + 16, b3 = b0 + 24, b4 = b0 + 32, nb = a.blocks_per_pair;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 v0 = __ldcg(base + (int64_t)b0 * (kAccPad / 4));
      const float4 v1 = b1 < nb ? __ldcg(base + (int64_t)b1 * (kAccPad / 4)) : z;
      const float4 v2 = b2 < nb ? __ldcg(base + (int64_t)b2 * (kAccPad / 4)) : z;
      const float4 v3 = b3 < nb ? __ldcg(base + (int64_t)b3 * (kAccPad / 4)) : z;
      const float4 v4 = b4 < nb ? __ldcg(base + (int64_t)b4 * (kAccPad / 4)) : z;
      s0 += (double)v0.x; s1 += (double)v0.y; s2 += (double)v0.z; s3 += (double)v0.w;
      if (b1 < nb) { s0 += (double)v1.x; s1 += (double)v1.y; s2 += (double)v1.z; s3 += (double)v1.w; }
      if (b2 < nb) { s0 += (double)v2.x; s1 += (double)v2.y; s2 += (double)v2.z; s3 += (double)v2.w; }
      if (b3 < nb) { s0 += (double)v3.x; s1 += (double)v3.y; s2 += (double)v3.z; s3 += (double)v3.w; }
      if (b4 < nb) { s0 += (double)v4.x; s1 += (double)v4.y; s2 += (double)v4.z; s3 += (double)v4.w; }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    if (sub == 0) { s_tot[cg4] = s0; s_tot[cg4 + 1] = s1; s_tot[cg4 + 2] = s2; s_tot[cg4 + 3] = s3; }   // columns >= kAcc: zeros, unused
  }
  __syncthreads();
  if (tid == 0) {
    double Rt[12];
    double* m = a.pose_master + 12 * pair;
#pragma unroll
    for (int k = 0; k < 12; ++k) Rt[k] = m[k];
    const SolveReport rep = solve_and_update(s_tot, a.min_count, a.damping, EARLY ? a.converge_eps : 0.f, a.update_pose != 0, Rt,
                                             a.stats + pair);
    if (rep.updated) {
#pragma unroll
      for (int k = 0; k < 12; ++k) m[k] = Rt[k];
      write_pose_outputs(Rt, a.pose_f32_out + 12 * pair, a.poses_cm + 16 * pair);
      if (EARLY && rep.converged) a.done[pair] = 1;  // read by the NEXT launch
    }
    a.tickets[pair] = 0u;  // ready for the next iteration / graph replay
  }
}

// ----------------------------------------------------------------------------------
// k_icp_fused: grid (C, n_pairs) in clusters of (C, 1, 1); cluster = one pair, all iterations of levels
// level_hi .. level_lo in one launch.
// ----------------------------------------------------------------------------------
constexpr int kTileGroups = 8;                               // groups per staged depth tile
constexpr int kTileChunks = kTileGroups * kChunksPerBlock;   // 64 chunks = 4096 px = 8 KB per buffer

template <int ROBUST, bool NGATE, bool PHOTO>
__global__ void __launch_bounds__(kIcpThreads, PHOTO ? 3 : (ROBUST != RST_ROBUST_NONE || NGATE) ? 4 : RST_FUSED_MINB)
k_icp_fused(const __grid_constant__ FusedArgs a) {
  // the two stage buffers live in dynamic shared memory: the photometric variants exceed the 48 KB static limit
  extern __shared__ __align__(16) unsigned char s_dyn[];
  StageBuf<PHOTO>* s_g = reinterpret_cast<StageBuf<PHOTO>*>(s_dyn);
  __shared__ __align__(16) uint32_t s_d[2][kTileChunks][32];   // double-buffered source-depth tiles
  __shared__ float s_warp[kIcpThreads / 32][kAccPad];
  __shared__ float s_part[kAccPad];     // this CTA's 29 sums of the current iteration (the leader reads them through DSMEM)
  __shared__ double s_tot[kAccPad];     // leader: cluster totals
  __shared__ float s_pose[16];          // leader: fp32 pose of the next iteration [0..11], [12] = flags (bit 0: level converged)
  __shared__ float s_cur[16];           // every CTA: its copy of the current pose (kept out of the registers of the pixel loop)
  __shared__ double s_master[12];       // leader: fp64 master pose
  __shared__ rst_stats s_stat;          // leader: statistics of the last evaluated iterate
  struct TileCtx { const uint16_t* Ds; int c_begin, n_my, n_tiles; };
  __shared__ TileCtx s_tc;              // this CTA's share of the current level (read at tile boundaries only)

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const unsigned C = cluster.num_blocks();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = a.pair_offset + blockIdx.y;
  const int2 slots = a.pairs[pair];
  const bool leader = rank == 0;

  if (tid < 12) s_cur[tid] = a.pose_f32[12 * pair + tid];
  if (leader) {
    if (tid < 12) s_master[tid] = a.pose_master[12 * pair + tid];
    if (tid == 0) s_stat = a.stats[pair];   // reset by k_init_pairs
  }
  __syncthreads();

  // One instantiation of the level body per pyramid level: a.lvl[l] then sits at a fixed offset of the kernel
  // parameters (constant bank operands) instead of in registers indexed by a run-time level.
  auto level_body = [&](auto level_tag) {
    constexpr int l = decltype(level_tag)::value;
    const FusedLevel& FL = a.lvl[l];
    const int W = FL.g.w, H = FL.g.h;
    const uint16_t* __restrict__ Ds = FL.lv.depth + (int64_t)slots.x * FL.lv.depth_frame;
    PixelPipe<ROBUST, NGATE, false, PHOTO> pipe;   // local to the level: its constants fold into parameter-bank operands
    pipe.tid = tid; pipe.lane = lane;
    pipe.P.depth_scale = a.depth_scale; pipe.P.dmax2 = a.dmax2; pipe.P.ncos_min = a.ncos_min;
    pipe.P.robust_scale = a.robust_scale; pipe.P.sqrt_lambda = a.sqrt_lambda; pipe.P.d_lo = a.d_lo; pipe.P.d_span = a.d_span;
    pipe.idx_out = nullptr;
    pipe.L.W = W; pipe.L.H = H; pipe.L.row_span = FL.chunks_per_row * kChunkPx;
    pipe.L.group_dv = FL.group_dv; pipe.L.group_du = FL.group_du;
    pipe.L.fx = FL.g.fx; pipe.L.fy = FL.g.fy; pipe.L.cx = FL.g.cx; pipe.L.cy = FL.g.cy; pipe.L.ifx = FL.g.ifx; pipe.L.ify = FL.g.ify;
    pipe.L.guard = FL.guard_texel << 4; pipe.L.w16 = (uint32_t)W << 4;
    pipe.L.Gs = FL.lv.geom + (int64_t)slots.x * FL.lv.geom_frame;
    pipe.L.Gd = FL.lv.geom + (int64_t)slots.y * FL.lv.geom_frame;
    pipe.L.Is = PHOTO ? FL.lv.intensity + (int64_t)slots.x * FL.lv.int_frame : nullptr;
    pipe.L.Id = PHOTO ? FL.lv.intensity + (int64_t)slots.y * FL.lv.int_frame : nullptr;
    // this CTA's contiguous share of the level's groups (a function of the image size and C only)
    const int g_begin = min((int)rank * FL.groups_per_cta, FL.n_groups);
    const int n_my = min(g_begin + FL.groups_per_cta, FL.n_groups) - g_begin;
    const int c_begin = g_begin * kChunksPerBlock;
    const int n_tiles = (n_my + kTileGroups - 1) / kTileGroups;
    __syncthreads();   // s_tc of the previous level is no longer read
    if (tid == 0) { s_tc.Ds = Ds; s_tc.c_begin = c_begin; s_tc.n_my = n_my; s_tc.n_tiles = n_tiles; }
    __syncthreads();
    auto stage_tile = [&](int t) {   // tile t of this CTA -> buffer t & 1 (its description comes from shared memory:
                                     // nothing of it stays in registers across the pixel loop)
      const int ng = min(kTileGroups, s_tc.n_my - t * kTileGroups);
      stage_depth(s_d[t & 1], s_tc.Ds, FL.lv.depth_pitch, W, H, FL.chunks_per_row, FL.cpr_magic, s_tc.c_begin + t * kTileChunks,
                  ng * kChunksPerBlock, tid);
    };

    for (int it = 0; it < FL.iters; ++it) {
      pipe.set_pose(s_cur);
      pipe.acc.clear();
      if (n_my > 0) {
        {
          const int c = c_begin + warp * kChunksPerWarp;
          const int v = FL.chunks_per_row == 1 ? c : (int)__umulhi((uint32_t)c, FL.cpr_magic);
          pipe.pos.v = v; pipe.pos.u = (c - v * FL.chunks_per_row) * kChunkPx + lane;
        }
        stage_tile(0);
        cp_async_wait<0>();
        __syncthreads();
        if (n_tiles > 1) stage_tile(1);
        pipe.sd = s_d[0]; pipe.cl = warp * kChunksPerWarp;
        pipe.template run<kTileGroups>(n_my, s_g, [&](int g) {
          // Tile T = g / kTileGroups: its cp.async group was committed a whole tile ago and has been forced complete
          // in every thread by the pipeline's wait_group calls since; the barrier makes it visible block-wide and
          // guarantees that nobody still reads tile T - 1, whose buffer tile T + 1 is staged into.
          const int T = g / kTileGroups;
          __syncthreads();
          if (T + 1 < s_tc.n_tiles) stage_tile(T + 1);
          pipe.sd = s_d[T & 1]; pipe.cl = warp * kChunksPerWarp;
        });
      }
      // ---- K5: block tree -> s_part; cluster totals in fixed rank order (fp64) on the leader; solve; new pose
      const float bs = block_reduce29(pipe.acc, s_warp, tid, lane, warp);
      if (tid < kAccPad) s_part[tid] = bs;
      cluster.sync();
      if (leader && warp == 0) {
        double s = 0.0;
        if (lane < kAcc)
          for (unsigned r = 0; r < C; ++r) s += (double)cluster.map_shared_rank(s_part, r)[lane];
        s_tot[lane] = s;
        __syncwarp();
        if (lane == 0) {
          const SolveReport rep = solve_and_update(s_tot, a.min_count, a.damping, a.converge_eps, true, s_master, &s_stat);
#pragma unroll
          for (int k = 0; k < 12; ++k) s_pose[k] = (float)s_master[k];
          s_pose[12] = __int_as_float(rep.updated && rep.converged ? 1 : 0);
        }
      }
      cluster.sync();
      const float* lp = cluster.map_shared_rank(s_pose, 0);
      if (tid < 13) s_cur[tid] = lp[tid];
      __syncthreads();
      if (__float_as_int(s_cur[12]) & 1) break;   // converged on this level (uniform across the cluster)
    }
  };
  static_assert(RST_MAX_LEVELS == 4, "level dispatch below is written for 4 levels");
  if (a.level_hi >= 3 && a.level_lo <= 3) level_body(std::integral_constant<int, 3>{});
  if (a.level_hi >= 2 && a.level_lo <= 2) level_body(std::integral_constant<int, 2>{});
  if (a.level_hi >= 1 && a.level_lo <= 1) level_body(std::integral_constant<int, 1>{});
  if (a.level_hi >= 0 && a.level_lo <= 0) level_body(std::integral_constant<int, 0>{});
  cluster.sync();   // nobody leaves while its shared memory may still be read by the cluster
  if (leader && tid == 0) {
    double* m = a.pose_master + 12 * pair;
#pragma unroll
    for (int k = 0; k < 12; ++k) m[k] = s_master[k];
    write_pose_outputs(s_master, a.pose_f32 + 12 * pair, a.poses_cm + 16 * pair);
    a.stats[pair] = s_stat;
  }
}

// ----------------------------------------------------------------------------------
// launchers
// ----------------------------------------------------------------------------------
// Function attributes are per device: `done` is a bit mask over device ordinals (a process may hold contexts on
// several GPUs; ordinals >= 64 simply set the attribute on every launch).
static bool attr_done(uint64_t* done, int* dev_out) {
  int dev = 0;
  cudaGetDevice(&dev);
  *dev_out = dev;
  return dev < 64 && ((*done >> dev) & 1ull);
}
static void attr_mark(uint64_t* done, int dev) { if (dev < 64) *done |= 1ull << dev; }

// dynamic shared memory of one block (the stage buffers); beyond 48 KB in total the kernel has to opt in once
template <bool PHOTO, class Kern>
static cudaError_t stage_smem_opt_in(Kern kern, uint64_t* done) {
  int dev;
  if (PHOTO && !attr_done(done, &dev)) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(StageBuf<PHOTO>)));
    if (e != cudaSuccess) return e;
    attr_mark(done, dev);
  }
  return cudaSuccess;
}

template <int ROBUST, bool NGATE, bool WRITE_IDX, bool PHOTO, bool EARLY>
static cudaError_t launch_icp_t(const IcpArgs& a, int n_pairs, cudaStream_t s) {
  dim3 grid(a.blocks_per_pair, n_pairs);
  auto kern = k_icp_iter<ROBUST, NGATE, WRITE_IDX, PHOTO, EARLY>;
  static uint64_t opted = 0;
  if (cudaError_t e = stage_smem_opt_in<PHOTO>(kern, &opted)) return e;
  const size_t smem = 2 * sizeof(StageBuf<PHOTO>);
  if (a.pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kIcpThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
  }
  kern<<<grid, kIcpThreads, smem, s>>>(a);
  return cudaGetLastError();
}

template <int ROBUST, bool PHOTO>
cudaError_t launch_icp_r(const IcpArgs& a, int n_pairs, bool ngate, bool widx, cudaStream_t s) {
  const bool early = a.done != nullptr;  // the convergence test never runs together with the index dump (rst_evaluate)
  if (widx) return ngate ? launch_icp_t<ROBUST, true, true, PHOTO, false>(a, n_pairs, s) : launch_icp_t<ROBUST, false, true, PHOTO, false>(a, n_pairs, s);
  if (early) return ngate ? launch_icp_t<ROBUST, true, false, PHOTO, true>(a, n_pairs, s) : launch_icp_t<ROBUST, false, false, PHOTO, true>(a, n_pairs, s);
  return ngate ? launch_icp_t<ROBUST, true, false, PHOTO, false>(a, n_pairs, s) : launch_icp_t<ROBUST, false, false, PHOTO, false>(a, n_pairs, s);
}

template <int ROBUST, bool NGATE, bool PHOTO>
static cudaError_t launch_fused_t(const FusedArgs& a, int n_pairs, int cluster, cudaStream_t s) {
  auto kern = k_icp_fused<ROBUST, NGATE, PHOTO>;
  if (cluster > 8) {   // beyond the portable cluster size: opt in once per instantiation and device
    static uint64_t allowed = 0;
    int dev;
    if (!attr_done(&allowed, &dev)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return e;
      attr_mark(&allowed, dev);
    }
  }
  static uint64_t opted = 0;
  if (cudaError_t e = stage_smem_opt_in<PHOTO>(kern, &opted)) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cluster, n_pairs, 1);
  cfg.blockDim = dim3(kIcpThreads, 1, 1);
  cfg.dynamicSmemBytes = 2 * sizeof(StageBuf<PHOTO>);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

template <int ROBUST>
cudaError_t launch_fused_r(const FusedArgs& a, int n_pairs, int cluster, bool ngate, bool photo, cudaStream_t s) {
  if (photo) return ngate ? launch_fused_t<ROBUST, true, true>(a, n_pairs, cluster, s) : launch_fused_t<ROBUST, false, true>(a, n_pairs, cluster, s);
  return ngate ? launch_fused_t<ROBUST, true, false>(a, n_pairs, cluster, s) : launch_fused_t<ROBUST, false, false>(a, n_pairs, cluster, s);
}

// ----------------------------------------------------------------------------------
// This file is compiled once per RST_ICP_PART (rst_icp_part*.cu): every part instantiates one slice of the kernel
// variants, so the slices build concurrently. Part 0 also holds the run-time dispatch.
// ----------------------------------------------------------------------------------
#define RST_DECL_ICP(R, P) template cudaError_t launch_icp_r<R, P>(const IcpArgs&, int, bool, bool, cudaStream_t)
#define RST_DECL_FUSED(R) template cudaError_t launch_fused_r<R>(const FusedArgs&, int, int, bool, bool, cudaStream_t)
#if RST_ICP_PART == 0
RST_DECL_ICP(RST_ROBUST_NONE, false);
#else
extern RST_DECL_ICP(RST_ROBUST_NONE, false);
#endif
#if RST_ICP_PART == 1
RST_DECL_ICP(RST_ROBUST_HUBER, false);
#else
extern RST_DECL_ICP(RST_ROBUST_HUBER, false);
#endif
#if RST_ICP_PART == 2
RST_DECL_ICP(RST_ROBUST_GEMAN_MCCLURE, false);
#else
extern RST_DECL_ICP(RST_ROBUST_GEMAN_MCCLURE, false);
#endif
#if RST_ICP_PART == 3
RST_DECL_ICP(RST_ROBUST_NONE, true);
#else
extern RST_DECL_ICP(RST_ROBUST_NONE, true);
#endif
#if RST_ICP_PART == 4
RST_DECL_ICP(RST_ROBUST_HUBER, true);
#else
extern RST_DECL_ICP(RST_ROBUST_HUBER, true);
#endif
#if RST_ICP_PART == 5
RST_DECL_ICP(RST_ROBUST_GEMAN_MCCLURE, true);
#else
extern RST_DECL_ICP(RST_ROBUST_GEMAN_MCCLURE, true);
#endif
#if RST_ICP_PART == 6
RST_DECL_FUSED(RST_ROBUST_NONE);
#else
extern RST_DECL_FUSED(RST_ROBUST_NONE);
#endif
#if RST_ICP_PART == 7
RST_DECL_FUSED(RST_ROBUST_HUBER);
#else
extern RST_DECL_FUSED(RST_ROBUST_HUBER);
#endif
#if RST_ICP_PART == 8
RST_DECL_FUSED(RST_ROBUST_GEMAN_MCCLURE);
#else
extern RST_DECL_FUSED(RST_ROBUST_GEMAN_MCCLURE);
#endif

#if RST_ICP_PART == 0
template <bool PHOTO>
static cudaError_t launch_icp_p(const IcpArgs& a, int n_pairs, int robust_kind, bool ngate, bool widx, cudaStream_t s) {
  switch (robust_kind) {
    case RST_ROBUST_HUBER: return launch_icp_r<RST_ROBUST_HUBER, PHOTO>(a, n_pairs, ngate, widx, s);
    case RST_ROBUST_GEMAN_MCCLURE: return launch_icp_r<RST_ROBUST_GEMAN_MCCLURE, PHOTO>(a, n_pairs, ngate, widx, s);
    default: return launch_icp_r<RST_ROBUST_NONE, PHOTO>(a, n_pairs, ngate, widx, s);
  }
}

cudaError_t launch_icp_iter(const IcpArgs& a, int n_pairs, int robust_kind, bool normal_gate, bool write_idx, bool photo,
                            cudaStream_t s) {
  if (n_pairs <= 0) return cudaSuccess;
  return photo ? launch_icp_p<true>(a, n_pairs, robust_kind, normal_gate, write_idx, s)
               : launch_icp_p<false>(a, n_pairs, robust_kind, normal_gate, write_idx, s);
}

cudaError_t launch_icp_fused(const FusedArgs& a, int n_pairs, int cluster, int robust_kind, bool normal_gate, bool photo,
                             cudaStream_t s) {
  if (n_pairs <= 0) return cudaSuccess;
  switch (robust_kind) {
    case RST_ROBUST_HUBER: return launch_fused_r<RST_ROBUST_HUBER>(a, n_pairs, cluster, normal_gate, photo, s);
    case RST_ROBUST_GEMAN_MCCLURE: return launch_fused_r<RST_ROBUST_GEMAN_MCCLURE>(a, n_pairs, cluster, normal_gate, photo, s);
    default: return launch_fused_r<RST_ROBUST_NONE>(a, n_pairs, cluster, normal_gate, photo, s);
  }
}

#endif  // part 0

#if RST_ICP_PART == 6
// how many clusters of `cluster` CTAs of the plain fused kernel the device can hold at once (0 on error)
int fused_max_active_clusters(int cluster) {
  auto kern = k_icp_fused<RST_ROBUST_NONE, false, false>;
  if (cluster > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cluster, 1024, 1);
  cfg.blockDim = dim3(kIcpThreads, 1, 1);
  cfg.dynamicSmemBytes = 2 * sizeof(StageBuf<false>);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

#endif  // part 6

}  // namespace rst
