// warp_math.h -- per-pixel coordinate arithmetic of the warp, shared by the CUDA kernels and by the host-side unit
// test (tests/host/warp_math_check.cpp compiles this header with g++ and compares it with true IEEE division).
//
// Reference sequence (all fp32, one rounding per operation, no contraction):
//   v = float(pix) + flow                       /root/reference/src/models/ema_vfi.py:162
//   g = 2.0 * v / max(size-1, 1) - 1.0          ema_vfi.py:165-166
//   i = ((g + 1) / 2) * (size - 1)              aten::grid_sampler_2d un-normalise, align_corners=True
// A 1-ulp error in i is 1.2e-4 px at 1080p and already breaks the 1e-5 parity bar on noise frames,
// so the division must be correctly rounded.  Instead of the ~15-instruction IEEE division subroutine we use the
// Markstein sequence q = a*y; r = fma(-q, d, a); q' = fma(r, y, q) with y = RN(1/d) computed once on the host;
// q' == RN(a/d) whenever no intermediate under/overflows (Markstein 1990, Thm 4; d is an integer < 2^24 here, so
// its significand is never all-ones).  Underflow only happens for |a| < 2^-100, where the following "- 1.0"
// absorbs the difference; overflow/NaN positions are pinned out of bounds by the caller either way.
#pragma once

#if defined(__CUDA_ARCH__)
#define VFI_HD __host__ __device__ __forceinline__
#define VFI_MUL(a, b) __fmul_rn((a), (b))
#define VFI_ADD(a, b) __fadd_rn((a), (b))
#define VFI_SUB(a, b) __fsub_rn((a), (b))
#define VFI_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#include <math.h>
#if defined(__CUDACC__)
#define VFI_HD __host__ __device__ inline
#else
#define VFI_HD static inline
#endif
// host build: compile with -ffp-contract=off so these stay separate roundings
#define VFI_MUL(a, b) ((a) * (b))
#define VFI_ADD(a, b) ((a) + (b))
#define VFI_SUB(a, b) ((a) - (b))
#define VFI_FMA(a, b, c) fmaf((a), (b), (c))
#endif

struct WarpAxis {
  float denom;      // float(max(size-1, 1))
  float inv_denom;  // RN(1 / denom), host-computed
  float size_m1;    // float(size-1)
  float hi;         // float(size) + 4: positions outside [-4, hi] touch no valid corner
  int recip;        // 1: a / d is computed as a * RN(1/d), which is what aten's CUDA `tensor / python_scalar` does
};

static inline WarpAxis make_warp_axis(long long size, int recip = 0) {
  WarpAxis a;
  a.recip = recip;
  a.denom = (float)(size - 1 > 1 ? size - 1 : 1);
  a.inv_denom = 1.0f / a.denom;
  a.size_m1 = (float)(size - 1);
  a.hi = (float)size + 4.0f;
  return a;
}

// Correctly rounded a / ax.denom (IEEE mode), or a * RN(1/denom) (reciprocal mode).
//
// Two references exist for this one line of ema_vfi.py: on CPU `tensor / python_scalar` is a true IEEE division, on
// CUDA aten multiplies by the fp32 reciprocal of the scalar (div_true_kernel_cuda).  The two differ by 1 ulp of the
// coordinate often enough to move noise-frame outputs by ~1e-3 at 1080p/4K, so both are offered (vfi_b200.h flags).
VFI_HD float vfi_div_exact(float a, const WarpAxis& ax) {
  float q = VFI_MUL(a, ax.inv_denom);
  if (ax.recip) return q;
  float r = VFI_FMA(-q, ax.denom, a);
  return VFI_FMA(r, ax.inv_denom, q);
}

// Source coordinate sampled by output pixel `pix` displaced by `disp`.
VFI_HD float vfi_warp_coord(int pix, float disp, const WarpAxis& ax) {
  float v = VFI_ADD((float)pix, disp);
  float g = VFI_MUL(2.0f, v);
  g = vfi_div_exact(g, ax);
  g = VFI_SUB(g, 1.0f);
  float t = VFI_ADD(g, 1.0f);
  t = VFI_MUL(t, 0.5f);  // division by 2 is exact
  float i = VFI_MUL(t, ax.size_m1);
  // far outside / NaN: pin into [-4, size + 4], where all four corners are still out of bounds (keeps float->int
  // defined; fmaxf/fminf return the non-NaN operand, so NaN lands on -4)
  return fminf(fmaxf(i, -4.0f), ax.hi);
}
