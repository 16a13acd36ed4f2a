/*
 * vfi_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 kernels.  It is imported only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product
 * (video-frame-interpolation_b200/) never links, imports or calls it.
 *
 * What it restates (all citations relative to /root/reference unless noted):
 *   - flow-guided backward warp: src/models/ema_vfi.py:149-171 (grid build :153-157, add flow :162,
 *     normalise :165-166, permute :168, F.grid_sample bilinear / zeros / align_corners=True :169).
 *     grid_sample's own arithmetic lives in PyTorch (aten::grid_sampler_2d, torch 2.11.0, not vendored
 *     in the reference); the published algorithm is restated from SURVEY.md Appendix A.
 *   - modulated deformable convolution v2: call site src/models/ema_vfi.py:60 (3x via :136-138),
 *     geometry fixed by :45-51 (3x3, stride 1, pad 1, dilation 1, groups 1, one offset group,
 *     mask on).  The arithmetic lives in torchvision 0.26.0 (torchvision::deform_conv2d and
 *     torchvision::_deform_conv2d_backward, only a compiled _C.so is present); the published
 *     algorithm is restated from SURVEY.md Appendix B.
 *
 * Pinning: the reference repository has no tests and no golden vectors, so the
 * oracle is pinned against outputs of the reference itself: the .npz files under tests/golden/ are produced by
 * tests/golden/make_golden.py, which imports /root/reference/src/models/ema_vfi.py unmodified and
 * runs EMA_VFI.warp, torchvision.ops.deform_conv2d and their autograd backward on CPU.
 * tests/test_oracle.py checks this file against every one of those fixtures.
 *
 * Numerics: the per-sample arithmetic (coordinates, bilinear weights, mask product) is IEEE fp32
 * in the reference's operation order with no FMA contraction (build with -ffp-contract=off).
 * Long reductions (the 603-term channel/tap contraction and the sums over pixels) accumulate in
 * double and round once, so the oracle sits at the centre of the rounding cloud of any fp32
 * summation order (torchvision hands that contraction to a BLAS whose order is unspecified).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(_OPENMP)
#include <omp.h>
#endif

#define VFI_ORACLE_VERSION 1

int vfi_oracle_version(void) { return VFI_ORACLE_VERSION; }

int vfi_oracle_threads(void) {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void vfi_oracle_set_threads(int n) {
#if defined(_OPENMP)
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------- */
/* Warp                                                                                        */
/* ------------------------------------------------------------------------------------------- */

/* Sampling position for output pixel (x, y) displaced by (fx, fy).
 * ema_vfi.py:162      v = grid + flow
 * ema_vfi.py:165-166  g = 2.0 * v / max(size-1, 1) - 1.0        (mul, div, sub: three roundings)
 * grid_sampler        i = ((g + 1) / 2) * (size - 1)            (align_corners=True un-normalise)
 * The round trip is not the identity in fp32, so it is replayed literally. */
static inline float sample_coord(int pix, float disp, int size) {
  float v = (float)pix + disp;
  float denom = (float)(size - 1 > 1 ? size - 1 : 1);
  float g = 2.0f * v;
  g = g / denom;
  g = g - 1.0f;
  float t = g + 1.0f;
  t = t / 2.0f;
  return t * (float)(size - 1);
}

typedef struct {
  int x0, y0;          /* north-west corner */
  float wx0, wx1;      /* weight of column x0 / x0+1 */
  float wy0, wy1;      /* weight of row    y0 / y0+1 */
} warp_tap;

static inline warp_tap warp_locate(int x, int y, float fx, float fy, int H, int W) {
  warp_tap t;
  float ix = sample_coord(x, fx, W);
  float iy = sample_coord(y, fy, H);
  /* Positions further than a few pixels outside the frame (or NaN) touch no valid corner; pin them to a
   * fixed out-of-bounds spot so the float->int cast below is always defined. */
  if (!(ix >= -4.0f)) ix = -4.0f; else if (ix > (float)W + 4.0f) ix = (float)W + 4.0f;
  if (!(iy >= -4.0f)) iy = -4.0f; else if (iy > (float)H + 4.0f) iy = (float)H + 4.0f;
  float fx0 = floorf(ix), fy0 = floorf(iy);
  t.x0 = (int)fx0;
  t.y0 = (int)fy0;
  t.wx1 = ix - fx0;            /* ix - x0 */
  t.wx0 = (fx0 + 1.0f) - ix;   /* x1 - ix */
  t.wy1 = iy - fy0;
  t.wy0 = (fy0 + 1.0f) - iy;
  return t;
}

static inline int inside(int x, int y, int H, int W) { return x >= 0 && x < W && y >= 0 && y < H; }

/* src [B,C,H,W], flow [B,2,H,W] (channel 0 = x displacement, 1 = y), out [B,C,H,W]; all NCHW fp32. */
void vfi_oracle_warp_fwd(const float* src, const float* flow, float* out, int B, int C, int H, int W) {
  const size_t plane = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int y = 0; y < H; ++y) {
      const float* fl = flow + (size_t)b * 2 * plane;
      for (int x = 0; x < W; ++x) {
        warp_tap t = warp_locate(x, y, fl[(size_t)y * W + x], fl[plane + (size_t)y * W + x], H, W);
        float w_nw = t.wx0 * t.wy0, w_ne = t.wx1 * t.wy0, w_sw = t.wx0 * t.wy1, w_se = t.wx1 * t.wy1;
        int in_nw = inside(t.x0, t.y0, H, W), in_ne = inside(t.x0 + 1, t.y0, H, W);
        int in_sw = inside(t.x0, t.y0 + 1, H, W), in_se = inside(t.x0 + 1, t.y0 + 1, H, W);
        for (int c = 0; c < C; ++c) {
          const float* s = src + ((size_t)b * C + c) * plane;
          float acc = 0.0f;
          if (in_nw) acc = acc + s[(size_t)t.y0 * W + t.x0] * w_nw;
          if (in_ne) acc = acc + s[(size_t)t.y0 * W + t.x0 + 1] * w_ne;
          if (in_sw) acc = acc + s[(size_t)(t.y0 + 1) * W + t.x0] * w_sw;
          if (in_se) acc = acc + s[(size_t)(t.y0 + 1) * W + t.x0 + 1] * w_se;
          out[((size_t)b * C + c) * plane + (size_t)y * W + x] = acc;
        }
      }
    }
  }
}

/* Gradient of the warp.  grad_flow [B,2,H,W] is always produced (the model path needs only this one:
 * frame2.requires_grad is False).  grad_src may be NULL; when given it must be
 * zero-initialised by the caller and receives the scatter-add.  floor() is treated as constant.
 * Chain through the normalisation: d(ix)/d(flow_x) = (2/(W-1)) * ((W-1)/2). */
void vfi_oracle_warp_bwd(const float* grad_out, const float* src, const float* flow, float* grad_flow,
                         float* grad_src, int B, int C, int H, int W) {
  const size_t plane = (size_t)H * W;
  const float mult_x = (float)(W - 1) / 2.0f, mult_y = (float)(H - 1) / 2.0f;
  const float den_x = (float)(W - 1 > 1 ? W - 1 : 1), den_y = (float)(H - 1 > 1 ? H - 1 : 1);
  /* serial over b,y when scattering into grad_src so the sum order is deterministic */
#pragma omp parallel for collapse(2) schedule(static) if (grad_src == NULL)
  for (int b = 0; b < B; ++b) {
    for (int y = 0; y < H; ++y) {
      const float* fl = flow + (size_t)b * 2 * plane;
      for (int x = 0; x < W; ++x) {
        const size_t p = (size_t)y * W + x;
        warp_tap t = warp_locate(x, y, fl[p], fl[plane + p], H, W);
        int in_nw = inside(t.x0, t.y0, H, W), in_ne = inside(t.x0 + 1, t.y0, H, W);
        int in_sw = inside(t.x0, t.y0 + 1, H, W), in_se = inside(t.x0 + 1, t.y0 + 1, H, W);
        double gix = 0.0, giy = 0.0;
        for (int c = 0; c < C; ++c) {
          const float* s = src + ((size_t)b * C + c) * plane;
          float g = grad_out[((size_t)b * C + c) * plane + p];
          float nw = in_nw ? s[(size_t)t.y0 * W + t.x0] : 0.0f;
          float ne = in_ne ? s[(size_t)t.y0 * W + t.x0 + 1] : 0.0f;
          float sw = in_sw ? s[(size_t)(t.y0 + 1) * W + t.x0] : 0.0f;
          float se = in_se ? s[(size_t)(t.y0 + 1) * W + t.x0 + 1] : 0.0f;
          gix += (double)(((ne - nw) * t.wy0 + (se - sw) * t.wy1) * g);
          giy += (double)(((sw - nw) * t.wx0 + (se - ne) * t.wx1) * g);
          if (grad_src) {
            float* gs = grad_src + ((size_t)b * C + c) * plane;
            if (in_nw) gs[(size_t)t.y0 * W + t.x0] += g * (t.wx0 * t.wy0);
            if (in_ne) gs[(size_t)t.y0 * W + t.x0 + 1] += g * (t.wx1 * t.wy0);
            if (in_sw) gs[(size_t)(t.y0 + 1) * W + t.x0] += g * (t.wx0 * t.wy1);
            if (in_se) gs[(size_t)(t.y0 + 1) * W + t.x0 + 1] += g * (t.wx1 * t.wy1);
          }
        }
        float ggx = mult_x * (float)gix;              /* grid_sampler backward: d/d(grid) */
        float ggy = mult_y * (float)giy;
        grad_flow[(size_t)b * 2 * plane + p] = (ggx / den_x) * 2.0f;           /* ema_vfi.py:165 */
        grad_flow[(size_t)b * 2 * plane + plane + p] = (ggy / den_y) * 2.0f;   /* ema_vfi.py:166 */
      }
    }
  }
}

/* North-star extension W3 (no reference counterpart): out = m*warp(a,fa) + (1-m)*warp(b,fb).
 * Oracle = composition of two reference warps and a lerp written as m*wa + (1-m)*wb. m is [B,1,H,W]. */
void vfi_oracle_warp_blend_fwd(const float* src_a, const float* flow_a, const float* src_b, const float* flow_b,
                               const float* m, float* out, int B, int C, int H, int W) {
  const size_t n = (size_t)B * C * H * W, plane = (size_t)H * W;
  float* wa = (float*)malloc(n * sizeof(float));
  float* wb = (float*)malloc(n * sizeof(float));
  vfi_oracle_warp_fwd(src_a, flow_a, wa, B, C, H, W);
  vfi_oracle_warp_fwd(src_b, flow_b, wb, B, C, H, W);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i) {
    size_t b = (size_t)i / ((size_t)C * plane), p = (size_t)i % plane;
    float mm = m[b * plane + p];
    out[i] = mm * wa[i] + (1.0f - mm) * wb[i];
  }
  free(wa);
  free(wb);
}

/* ------------------------------------------------------------------------------------------- */
/* Modulated deformable convolution, 3x3 / stride 1 / pad 1 / dilation 1 / one group           */
/* ------------------------------------------------------------------------------------------- */

typedef struct {
  int live;            /* 0: the whole sample is outside (-1, H) x (-1, W) and contributes nothing */
  int y0, x0;          /* floor of the sampling position */
  float lh, lw;        /* fractional parts */
  int ok00, ok01, ok10, ok11; /* corner (row y0|y0+1, col x0|x0+1) lies inside the image */
} dcn_tap;

/* Sampling position of tap (i, j) for output pixel (y, x): Appendix B.
 * offset channel 2k is the row displacement, 2k+1 the column displacement, k = 3*i + j. */
static inline dcn_tap dcn_locate(int y, int x, int i, int j, float dy, float dx, int H, int W) {
  dcn_tap t;
  float py = (float)(y - 1 + i) + dy;
  float px = (float)(x - 1 + j) + dx;
  t.live = (py > -1.0f) && (py < (float)H) && (px > -1.0f) && (px < (float)W);
  /* Far-away (or NaN) positions have no valid corner; pin them so the int cast is defined.  Note the corner
   * flags are kept even when the sample is not live: torchvision's offset gradient is NOT gated on liveness
   * (it differs from the gated form exactly when py == -1 or px == -1), see vfi_oracle_dcn_bwd. */
  if (!(py > -2.0f && py < (float)H + 1.0f)) py = -2.0f;
  if (!(px > -2.0f && px < (float)W + 1.0f)) px = -2.0f;
  float fy = floorf(py), fx = floorf(px);
  t.y0 = (int)fy; t.x0 = (int)fx;
  t.lh = py - fy; t.lw = px - fx;
  int r0 = t.y0 >= 0 && t.y0 <= H - 1, r1 = t.y0 + 1 >= 0 && t.y0 + 1 <= H - 1;
  int c0 = t.x0 >= 0 && t.x0 <= W - 1, c1 = t.x0 + 1 >= 0 && t.x0 + 1 <= W - 1;
  t.ok00 = r0 && c0; t.ok01 = r0 && c1; t.ok10 = r1 && c0; t.ok11 = r1 && c1;
  return t;
}

static inline float dcn_sample(const float* plane, const dcn_tap* t, int W) {
  if (!t->live) return 0.0f;
  float hh = 1.0f - t->lh, hw = 1.0f - t->lw;
  float v00 = t->ok00 ? plane[(size_t)t->y0 * W + t->x0] : 0.0f;
  float v01 = t->ok01 ? plane[(size_t)t->y0 * W + t->x0 + 1] : 0.0f;
  float v10 = t->ok10 ? plane[(size_t)(t->y0 + 1) * W + t->x0] : 0.0f;
  float v11 = t->ok11 ? plane[(size_t)(t->y0 + 1) * W + t->x0 + 1] : 0.0f;
  float w00 = hh * hw, w01 = hh * t->lw, w10 = t->lh * hw, w11 = t->lh * t->lw;
  float acc = w00 * v00;
  acc = acc + w01 * v01;
  acc = acc + w10 * v10;
  acc = acc + w11 * v11;
  return acc;
}

/* x [B,C,H,W], offset [B,18,H,W], mask [B,9,H,W], weight [O,C,3,3], bias [O] or NULL, out [B,O,H,W]. */
void vfi_oracle_dcn_fwd(const float* x, const float* offset, const float* mask, const float* weight,
                        const float* bias, float* out, int B, int C, int O, int H, int W) {
  const size_t plane = (size_t)H * W;
  const int K = C * 9;
#pragma omp parallel
  {
    float* col = (float*)malloc((size_t)K * sizeof(float));
#pragma omp for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
      for (int y = 0; y < H; ++y) {
        for (int xx = 0; xx < W; ++xx) {
          const size_t p = (size_t)y * W + xx;
          for (int k = 0; k < 9; ++k) {
            float dy = offset[((size_t)b * 18 + 2 * k) * plane + p];
            float dx = offset[((size_t)b * 18 + 2 * k + 1) * plane + p];
            float mk = mask[((size_t)b * 9 + k) * plane + p];
            dcn_tap t = dcn_locate(y, xx, k / 3, k % 3, dy, dx, H, W);
            for (int c = 0; c < C; ++c)
              col[c * 9 + k] = mk * dcn_sample(x + ((size_t)b * C + c) * plane, &t, W);
          }
          for (int o = 0; o < O; ++o) {
            const float* wrow = weight + (size_t)o * K;
            double acc = 0.0;
            for (int q = 0; q < K; ++q) acc += (double)wrow[q] * (double)col[q];
            if (bias) acc += (double)bias[o];
            out[((size_t)b * O + o) * plane + p] = (float)acc;
          }
        }
      }
    }
    free(col);
  }
}

/* All five gradients of the layer above.  Any output pointer may be NULL.
 * grad_x must be zero-initialised by the caller when given.  Internally every reduction is done in
 * double and rounded once at the end, serially in a fixed order (small test sizes only). */
void vfi_oracle_dcn_bwd(const float* grad_out, const float* x, const float* offset, const float* mask,
                        const float* weight, float* grad_x, float* grad_offset, float* grad_mask,
                        float* grad_weight, float* grad_bias, int B, int C, int O, int H, int W) {
  const size_t plane = (size_t)H * W;
  const int K = C * 9;
  double* gw = grad_weight ? (double*)calloc((size_t)O * K, sizeof(double)) : NULL;
  double* gb = grad_bias ? (double*)calloc((size_t)O, sizeof(double)) : NULL;
  double* gx = grad_x ? (double*)calloc((size_t)B * C * plane, sizeof(double)) : NULL;
  float* col = (float*)malloc((size_t)K * sizeof(float));
  double* gcol = (double*)malloc((size_t)K * sizeof(double));

  for (int b = 0; b < B; ++b) {
    for (int y = 0; y < H; ++y) {
      for (int xx = 0; xx < W; ++xx) {
        const size_t p = (size_t)y * W + xx;
        /* gcol = W^T * grad_out at this pixel */
        for (int q = 0; q < K; ++q) gcol[q] = 0.0;
        for (int o = 0; o < O; ++o) {
          double g = (double)grad_out[((size_t)b * O + o) * plane + p];
          if (gb) gb[o] += g;
          const float* wrow = weight + (size_t)o * K;
          for (int q = 0; q < K; ++q) gcol[q] += (double)wrow[q] * g;
        }
        for (int k = 0; k < 9; ++k) {
          float dy = offset[((size_t)b * 18 + 2 * k) * plane + p];
          float dx = offset[((size_t)b * 18 + 2 * k + 1) * plane + p];
          float mk = mask[((size_t)b * 9 + k) * plane + p];
          dcn_tap t = dcn_locate(y, xx, k / 3, k % 3, dy, dx, H, W);
          double g_m = 0.0, g_dy = 0.0, g_dx = 0.0;
          float hh = 1.0f - t.lh, hw = 1.0f - t.lw;
          for (int c = 0; c < C; ++c) {
            const float* pl = x + ((size_t)b * C + c) * plane;
            float val = dcn_sample(pl, &t, W);
            col[c * 9 + k] = mk * val;
            double gc = gcol[c * 9 + k];
            g_m += gc * (double)val;
            {
              float v00 = t.ok00 ? pl[(size_t)t.y0 * W + t.x0] : 0.0f;
              float v01 = t.ok01 ? pl[(size_t)t.y0 * W + t.x0 + 1] : 0.0f;
              float v10 = t.ok10 ? pl[(size_t)(t.y0 + 1) * W + t.x0] : 0.0f;
              float v11 = t.ok11 ? pl[(size_t)(t.y0 + 1) * W + t.x0 + 1] : 0.0f;
              /* d(val)/d(py) and d(val)/d(px); a corner outside the image counts as value 0.  Not gated on
               * t.live, as in the reference (only matters when py or px is exactly -1). */
              float d_py = t.lw * (v11 - v01) + hw * (v10 - v00);
              float d_px = t.lh * (v11 - v10) + hh * (v01 - v00);
              g_dy += gc * (double)mk * (double)d_py;
              g_dx += gc * (double)mk * (double)d_px;
              if (gx && t.live) {
                double* gp = gx + ((size_t)b * C + c) * plane;
                double gm = gc * (double)mk;
                if (t.ok00) gp[(size_t)t.y0 * W + t.x0] += gm * (double)(hh * hw);
                if (t.ok01) gp[(size_t)t.y0 * W + t.x0 + 1] += gm * (double)(hh * t.lw);
                if (t.ok10) gp[(size_t)(t.y0 + 1) * W + t.x0] += gm * (double)(t.lh * hw);
                if (t.ok11) gp[(size_t)(t.y0 + 1) * W + t.x0 + 1] += gm * (double)(t.lh * t.lw);
              }
            }
          }
          if (grad_mask) grad_mask[((size_t)b * 9 + k) * plane + p] = (float)g_m;
          if (grad_offset) {
            grad_offset[((size_t)b * 18 + 2 * k) * plane + p] = (float)g_dy;
            grad_offset[((size_t)b * 18 + 2 * k + 1) * plane + p] = (float)g_dx;
          }
        }
        if (gw) {
          for (int o = 0; o < O; ++o) {
            double g = (double)grad_out[((size_t)b * O + o) * plane + p];
            double* row = gw + (size_t)o * K;
            for (int q = 0; q < K; ++q) row[q] += g * (double)col[q];
          }
        }
      }
    }
  }
  if (gw) { for (size_t i = 0; i < (size_t)O * K; ++i) grad_weight[i] = (float)gw[i]; free(gw); }
  if (gb) { for (int o = 0; o < O; ++o) grad_bias[o] = (float)gb[o]; free(gb); }
  if (gx) { for (size_t i = 0; i < (size_t)B * C * plane; ++i) grad_x[i] = (float)gx[i]; free(gx); }
  free(col);
  free(gcol);
}

/* ModulatedDeformConvPack glue, src/models/ema_vfi.py:57-59: the 27-channel offset_conv output is split
 * into three 9-channel thirds; thirds 0 and 2 are concatenated into the 18-channel offset, the middle
 * third goes through a sigmoid and becomes the modulation mask. */
void vfi_oracle_pack_split(const float* conv27, float* offset18, float* mask9, int B, int H, int W) {
  const size_t plane = (size_t)H * W;
  for (int b = 0; b < B; ++b) {
    const float* in = conv27 + (size_t)b * 27 * plane;
    memcpy(offset18 + (size_t)b * 18 * plane, in, 9 * plane * sizeof(float));
    memcpy(offset18 + (size_t)b * 18 * plane + 9 * plane, in + 18 * plane, 9 * plane * sizeof(float));
    for (size_t i = 0; i < 9 * plane; ++i)
      mask9[(size_t)b * 9 * plane + i] = 1.0f / (1.0f + expf(-in[9 * plane + i]));
  }
}
