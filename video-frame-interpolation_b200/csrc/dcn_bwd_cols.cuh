// dcn_bwd_cols.cuh -- included by dcn_tc.cu inside namespace vfi::{anonymous}.
//
// Backward of the modulated deformable convolution with respect to the input, the offsets and the mask
// (torchvision::_deform_conv2d_backward's grad_input / grad_offset / grad_mask, reached from
// /root/reference/src/models/ema_vfi.py:60 through autograd; arithmetic: SURVEY.md Appendix B) for the bf16 training
// path, split the way the arithmetic splits:
//
//   1. gcol[p, (k, c)] = sum_o gout[p, o] * W[o, c, k]        a plain dense GEMM [P, 72] x [72, 648] with no gather in
//      it: the caller makes it with its BLAS (bf16 operands, fp32 accumulation, bf16 result, 1.3 KB per pixel);
//   2. this kernel: everything that depends on the sampling positions.  Per (pixel p, tap k), with v_j the four corner
//      rows of x (64 + 3 channels, channels-last planes) and D_j = <gcol[p, k, :], v_j>:
//        grad_mask[p, k]     = live ? sum_j w_j D_j : 0
//        grad_offset[p, 2k]  = m * (lw (D11 - D01) + hw (D10 - D00))         (y),   2k + 1 likewise in x
//        grad_x[corner_j, :] += gcol[p, k, :] * m * w_j                      (live samples, corners inside the image)
//
// Mapping: a CTA owns 32 consecutive pixels of one image (288 pixel-taps).  Phase A, one thread per pixel-tap (lanes =
// consecutive pixels, so the NCHW offset / mask reads and the grad_offset / grad_mask writes coalesce): sampling
// geometry once per pixel-tap, plus the three tail channels.  Phase B, one half-warp per (pixel, kernel row) with lane = four
// consecutive channels: 8- or 16-byte loads that make full rows (gcol and the four corners), a 16-lane shuffle
// reduction for the three dot products, and grad_x as four `red.global.add.v4.f32` per lane into a channels-last fp32
// accumulator -- at most 36 x 17 vector reductions per pixel (24 x 17 when neighbouring taps share corners, see phase B)
// where the NCHW kernel (dcn_simt.cu) issues 36 x 67 scalar ones.
// Phase C writes grad_offset / grad_mask.  HBM/L2-bound by the reductions; no tensor-core work in here.

#ifndef BC_MIN_BLOCKS
#define BC_MIN_BLOCKS 4      // 56 registers, four CTAs per SM: measured 2-7 % faster than 70 registers / three CTAs
#endif
#ifndef BC_MERGE
#define BC_MERGE 1
#endif
constexpr int BC_PIX = 32;                    // pixels per CTA
constexpr int BC_PT = BC_PIX * 9;             // pixel-taps per CTA
constexpr int BC_THREADS = BC_PT;             // 9 warps
constexpr int BC_TAP_LD = 72;                 // gcol columns per tap: 64 main, 3 tail, 5 zero

struct BcParams {
  const uint8_t* x_main; const uint8_t* x_tail;            // planes of TP (bf16 or f32): 64 main channels / the tail record (channels 64.. first)
  long long main_px, tail_px;                              // bytes from one pixel to the next (dense planes: 64 / 8 elements; one
                                                           // [B,H,W,72] record buffer: 144 bytes for both)
  const void* offset; const void* mask;                    // FUSED27: both unused fields point at the 27-channel offset_conv output
  long long f_sn, f_sc, f_sh, f_sw, m_sn, m_sc, m_sh, m_sw;
  const void* gcol; long long gcol_ld;                     // [P][gcol_ld] of TP, column k * 72 + c
  float* gx; long long gx_ld;                              // [P][gx_ld] fp32 accumulator, channel c at column c (may be null)
  float* goff; long long gf_sn, gf_sc, gf_sh, gf_sw;       // [B,18,H,W] f32 (may be null)
  float* gmask; long long gm_sn, gm_sc, gm_sh, gm_sw;      // [B,9,H,W] f32 (may be null)
  // FUSED27: goff is the gradient of the raw 27-channel tensor (channel map below), gmask unused
  int B, H, W;
};

// FUSED27 (the ModulatedDeformConvPack glue of ema_vfi.py:57-59 folded in, as in the forward kernels): channel of the raw
// offset_conv output that holds offset channel j (j = 2k: dy, 2k + 1: dx of tap k) / the mask logit of tap k.
__device__ __forceinline__ int bc_off_chan(int j) { return j < 9 ? j : j + 9; }
__device__ __forceinline__ int bc_mask_chan(int k) { return 9 + k; }

struct BcGeo {
  int pix00;                    // linear pixel index (b, y0, x0) of the upper-left corner; only dereferenced when its flag is set
  int flags;                    // bit 0..3: corner 00, 01, 10, 11 inside the image; bit 4: sample live; bit 5: pixel exists
  float lh, lw, mk;
};

struct BcSmem {
  BcGeo geo[BC_PT];             // [k * 32 + pixel]
  float part[3][BC_PT];         // (mask, dy, dx) partial sums: tail channels from phase A, + main channels from phase B
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
// Four consecutive channels of a plane / gcol row as fp32 (zeros when the corner is outside the image).
template <typename TP> __device__ __forceinline__ void load4(const uint8_t* p, bool ok, float (&v)[4]);
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const uint8_t* p, bool ok, float (&v)[4]) {
  const uint2 r = ok ? __ldg(reinterpret_cast<const uint2*>(p)) : make_uint2(0u, 0u);
  v[0] = bf_lo(r.x); v[1] = bf_hi(r.x); v[2] = bf_lo(r.y); v[3] = bf_hi(r.y);
}
template <> __device__ __forceinline__ void load4<float>(const uint8_t* p, bool ok, float (&v)[4]) {
  const float4 r = ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
__device__ __forceinline__ float dot4(const float (&g)[4], const float (&v)[4]) {
  float d = g[0] * v[0];
  d = fmaf(g[1], v[1], d);
  d = fmaf(g[2], v[2], d);
  return fmaf(g[3], v[3], d);
}

// TO: dtype of offset / mask; TP: dtype of gcol and of the x planes (bf16: the tensor-core training path; f32: the fp32 path).
// FUSED27: offsets and mask logits come from the raw 27-channel offset_conv output (sigmoid folded in, rounded to TO as the
// forward kernels do), and the gradient goes back to that tensor (d sigmoid folded in).
template <typename TO, typename TP, bool FUSED27>
__global__ void __launch_bounds__(BC_THREADS, BC_MIN_BLOCKS) dcn_bwd_cols_kernel(const BcParams q) {
  constexpr int ES = (int)sizeof(TP);
  const long long MAIN_PX = q.main_px, TAIL_PX = q.tail_px;
  const uint8_t* gcol = reinterpret_cast<const uint8_t*>(q.gcol);
  __shared__ BcSmem s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const long long HW = (long long)q.H * q.W;
  const long long p0 = (long long)blockIdx.x * BC_PIX;
  const long long img0 = (long long)b * HW;                    // linear index of this image's first pixel

  // ------------------------------------------------------------------ phase A: geometry + tail channels, thread = (tap, pixel)
  {
    const int k = warp, pxl = lane;
    const long long pp = p0 + pxl;
    BcGeo g; g.pix00 = 0; g.flags = 0; g.lh = g.lw = g.mk = 0.0f;
    float t_m = 0.0f, t_dy = 0.0f, t_dx = 0.0f;
    if (pp < HW) {
      const int y = (int)(pp / q.W), x = (int)(pp % q.W);
      const TO* off = reinterpret_cast<const TO*>(q.offset) + b * q.f_sn + y * q.f_sh + x * q.f_sw;
      float dy, dx;
      if (FUSED27) {
        dy = to_f32<TO>(__ldg(off + bc_off_chan(2 * k) * q.f_sc));
        dx = to_f32<TO>(__ldg(off + bc_off_chan(2 * k + 1) * q.f_sc));
        const float logit = to_f32<TO>(__ldg(off + bc_mask_chan(k) * q.f_sc));
        g.mk = to_f32<TO>(from_f32<TO>(__fdividef(1.0f, 1.0f + __expf(-logit))));   // the forward kernels' sigmoid, rounded to TO
      } else {
        const TO* msk = reinterpret_cast<const TO*>(q.mask) + b * q.m_sn + y * q.m_sh + x * q.m_sw;
        dy = to_f32<TO>(__ldg(off + (2 * k) * q.f_sc));
        dx = to_f32<TO>(__ldg(off + (2 * k + 1) * q.f_sc));
        g.mk = to_f32<TO>(__ldg(msk + k * q.m_sc));
      }
      float py = __fadd_rn((float)(y - 1 + k / 3), dy);
      float px = __fadd_rn((float)(x - 1 + k % 3), dx);
      const bool live = (py > -1.0f) && (py < (float)q.H) && (px > -1.0f) && (px < (float)q.W);
      if (!(py > -2.0f && py < (float)q.H + 1.0f)) py = -2.0f;  // no valid corner out there; keeps the cast defined
      if (!(px > -2.0f && px < (float)q.W + 1.0f)) px = -2.0f;
      const float fy = floorf(py), fx = floorf(px);
      const int y0 = (int)fy, x0 = (int)fx;
      g.lh = py - fy; g.lw = px - fx;
      const bool r0 = (unsigned)y0 < (unsigned)q.H, r1 = (unsigned)(y0 + 1) < (unsigned)q.H;
      const bool c0 = (unsigned)x0 < (unsigned)q.W, c1 = (unsigned)(x0 + 1) < (unsigned)q.W;
      g.flags = (r0 && c0 ? 1 : 0) | (r0 && c1 ? 2 : 0) | (r1 && c0 ? 4 : 0) | (r1 && c1 ? 8 : 0) | (live ? 16 : 0) | 32;
      g.pix00 = (int)(img0 + (long long)y0 * q.W + x0);
      // tail channels (64..66): four elements of gcol, four per corner
      float gc[4], v00[4], v01[4], v10[4], v11[4];
      load4<TP>(gcol + ((size_t)(img0 + pp) * q.gcol_ld + k * BC_TAP_LD + 64) * ES, true, gc);
      const uint8_t* xt = q.x_tail + (long long)g.pix00 * TAIL_PX;
      load4<TP>(xt, g.flags & 1, v00); load4<TP>(xt + TAIL_PX, g.flags & 2, v01);
      load4<TP>(xt + (long long)q.W * TAIL_PX, g.flags & 4, v10); load4<TP>(xt + ((long long)q.W + 1) * TAIL_PX, g.flags & 8, v11);
      const float d00 = dot4(gc, v00), d01 = dot4(gc, v01), d10 = dot4(gc, v10), d11 = dot4(gc, v11);
      const float hh = 1.0f - g.lh, hw = 1.0f - g.lw;
      const float w00 = hh * hw, w01 = hh * g.lw, w10 = g.lh * hw, w11 = g.lh * g.lw;
      t_m = live ? (w00 * d00 + w01 * d01 + w10 * d10 + w11 * d11) : 0.0f;
      t_dy = g.mk * (g.lw * (d11 - d01) + hw * (d10 - d00));
      t_dx = g.mk * (g.lh * (d11 - d10) + hh * (d01 - d00));
      if (q.gx && live) {
        float* gp = q.gx + (long long)g.pix00 * q.gx_ld + 64;
        const float m0 = gc[0] * g.mk, m1 = gc[1] * g.mk, m2 = gc[2] * g.mk, m3 = gc[3] * g.mk;   // gcol column 67 is zero when C < 68
        if (g.flags & 1) red_add_v4(gp, m0 * w00, m1 * w00, m2 * w00, m3 * w00);
        if (g.flags & 2) red_add_v4(gp + q.gx_ld, m0 * w01, m1 * w01, m2 * w01, m3 * w01);
        if (g.flags & 4) red_add_v4(gp + (long long)q.W * q.gx_ld, m0 * w10, m1 * w10, m2 * w10, m3 * w10);
        if (g.flags & 8) red_add_v4(gp + ((long long)q.W + 1) * q.gx_ld, m0 * w11, m1 * w11, m2 * w11, m3 * w11);
      }
    }
    s.geo[tid] = g;
    s.part[0][tid] = t_m; s.part[1][tid] = t_dy; s.part[2][tid] = t_dx;
  }
  __syncthreads();

  // ------------------------------------------------------------------ phase B: main channels, half-warp = (pixel, kernel row), lane = 4 channels
  // A half-warp walks the three taps of one kernel row.  With offsets that vary slowly from tap to tap (what a trained
  // offset_conv produces) the right-hand corners of tap j are the left-hand corners of tap j + 1: their grad_x contributions
  // are added in registers and leave as one reduction -- 8 instead of 12 per kernel row.  The test is on the pixel index,
  // so arbitrary offsets are handled (nothing merges, nothing is lost).
  if (warp < 8) {
    const int half = lane >> 4, l16 = lane & 15;
    struct Pending { float v[4]; int pix; bool on; };
    const auto flush = [&](Pending& pd) {
      if (pd.on) red_add_v4(q.gx + (long long)pd.pix * q.gx_ld + 4 * l16, pd.v[0], pd.v[1], pd.v[2], pd.v[3]);
      pd.on = false;
    };
    // left-hand corner `pix` with weight w: absorbs the pending right-hand corner of the previous tap when it is the same pixel
    const auto left = [&](Pending& pd, bool ok, int pix, float w, const float (&m)[4]) {
      float a[4] = {m[0] * w, m[1] * w, m[2] * w, m[3] * w};
      if (BC_MERGE && pd.on && ok && pd.pix == pix) {
#pragma unroll
        for (int c = 0; c < 4; ++c) a[c] += pd.v[c];
        pd.on = false;
      }
      flush(pd);
      if (ok) red_add_v4(q.gx + (long long)pix * q.gx_ld + 4 * l16, a[0], a[1], a[2], a[3]);
    };
    const auto right = [&](Pending& pd, bool ok, int pix, float w, const float (&m)[4]) {
      pd.on = ok; pd.pix = pix;
#pragma unroll
      for (int c = 0; c < 4; ++c) pd.v[c] = m[c] * w;
    };
#pragma unroll 1
    for (int it = 0; it < (BC_PIX * 3) / 16; ++it) {
      const int unit = it * 16 + warp * 2 + half;               // consecutive half-warps: the three kernel rows of one pixel
      const int pxl = unit / 3, krow = unit - pxl * 3;
      Pending top, bot;
      top.on = bot.on = false; top.pix = bot.pix = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) top.v[c] = bot.v[c] = 0.0f;
#pragma unroll
      for (int kj = 0; kj < 3; ++kj) {
        const int k = krow * 3 + kj;
        const int idx = k * 32 + pxl;
        const BcGeo g = s.geo[idx];
        const bool exists = (g.flags & 32) != 0, live = (g.flags & 16) != 0;
        float gc[4], v00[4], v01[4], v10[4], v11[4];
        load4<TP>(gcol + ((size_t)(img0 + p0 + pxl) * q.gcol_ld + k * BC_TAP_LD + 4 * l16) * ES, exists, gc);
        const uint8_t* xm = q.x_main + (long long)g.pix00 * MAIN_PX + 4 * ES * l16;
        load4<TP>(xm, g.flags & 1, v00); load4<TP>(xm + MAIN_PX, g.flags & 2, v01);
        load4<TP>(xm + (long long)q.W * MAIN_PX, g.flags & 4, v10); load4<TP>(xm + ((long long)q.W + 1) * MAIN_PX, g.flags & 8, v11);
        const float d00 = dot4(gc, v00), d01 = dot4(gc, v01), d10 = dot4(gc, v10), d11 = dot4(gc, v11);
        const float hh = 1.0f - g.lh, hw = 1.0f - g.lw;
        const float w00 = hh * hw, w01 = hh * g.lw, w10 = g.lh * hw, w11 = g.lh * g.lw;
        float r_m = live ? (w00 * d00 + w01 * d01 + w10 * d10 + w11 * d11) : 0.0f;
        float r_dy = g.mk * (g.lw * (d11 - d01) + hw * (d10 - d00));
        float r_dx = g.mk * (g.lh * (d11 - d10) + hh * (d01 - d00));
        if (q.gx) {
          const float m[4] = {gc[0] * g.mk, gc[1] * g.mk, gc[2] * g.mk, gc[3] * g.mk};
          left(top, live && (g.flags & 1), g.pix00, w00, m);
          left(bot, live && (g.flags & 4), g.pix00 + q.W, w10, m);
          right(top, live && (g.flags & 2), g.pix00 + 1, w01, m);
          right(bot, live && (g.flags & 8), g.pix00 + q.W + 1, w11, m);
        }
#pragma unroll
        for (int sh = 8; sh >= 1; sh >>= 1) {                   // stays inside the half-warp
          r_m += __shfl_xor_sync(0xffffffffu, r_m, sh);
          r_dy += __shfl_xor_sync(0xffffffffu, r_dy, sh);
          r_dx += __shfl_xor_sync(0xffffffffu, r_dx, sh);
        }
        if (l16 == 0) { s.part[0][idx] += r_m; s.part[1][idx] += r_dy; s.part[2][idx] += r_dx; }
      }
      if (q.gx) { flush(top); flush(bot); }
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ phase C: grad_mask / grad_offset, thread = (tap, pixel)
  {
    const int k = warp;
    const long long pp = p0 + lane;
    if (pp < HW) {
      const int y = (int)(pp / q.W), x = (int)(pp % q.W);
      if (FUSED27) {
        if (q.goff) {
          float* go = q.goff + b * q.gf_sn + y * q.gf_sh + x * q.gf_sw;
          const float mk = s.geo[tid].mk;
          go[bc_off_chan(2 * k) * q.gf_sc] = s.part[1][tid];
          go[bc_off_chan(2 * k + 1) * q.gf_sc] = s.part[2][tid];
          go[bc_mask_chan(k) * q.gf_sc] = s.part[0][tid] * mk * (1.0f - mk);        // d sigmoid
        }
      } else {
        if (q.gmask) q.gmask[b * q.gm_sn + k * q.gm_sc + y * q.gm_sh + x * q.gm_sw] = s.part[0][tid];
        if (q.goff) {
          float* go = q.goff + b * q.gf_sn + y * q.gf_sh + x * q.gf_sw;
          go[(2 * k) * q.gf_sc] = s.part[1][tid];
          go[(2 * k + 1) * q.gf_sc] = s.part[2][tid];
        }
      }
    }
  }
}
