// dcn_simt.cu -- modulated deformable convolution (DCNv2, 3x3 / s1 / p1 / d1 / one group) in fp32 on the CUDA cores.
//
// This is the PARITY path (VFI_DCN_MATH_FP32): fp32 gather in torchvision's operation order and an fp32 FFMA
// contraction, so results sit within ~1e-6 of torchvision's fp32 kernels (bar: max-abs 1e-5).  It holds the fp32 forward,
// the fp32 weight gradient and the generic (any C <= 72, any strides) data gradient; the Python host prefers the
// column-gradient form of the data gradient (dcn_bwd_cols.cuh) whenever C <= 68.  The throughput path is dcn_tc.cu
// (bf16 operands, tcgen05/TMEM).
//
// Replaces torchvision::deform_conv2d / _deform_conv2d_backward as reached from
// /root/reference/src/models/ema_vfi.py:60 (geometry :45-51).  Arithmetic: SURVEY.md Appendix B.
//
// Unlike torchvision there is no materialised `columns` buffer (40 GB fp32 at 1080p B=8): every kernel is an
// implicit GEMM whose column tile lives in shared memory for one tap at a time.
//   fwd        : out[128 px, 72 o]   += col_k[128 px, 72 c] * Wk[72 c, 72 o]          per tap k
//   bwd data   : gcol_k[128 px, 72 c] = gout[128 px, 72 o]  * Wk^T[72 o, 72 c]        then gather/scatter epilogue
//   bwd weight : gW_k[72 o, 72 c]    += gout^T[72 o, 32 px] * col_k[32 px, 72 c]      persistent over pixel tiles
// Channel counts are limited to C, O <= 72 here (the model uses 67); larger -> VFI_ERR_UNSUPPORTED.
#include "common.cuh"

namespace vfi {
namespace {

constexpr int CP = 72;          // padded channel count (rows of the weight tiles)
constexpr int TP = 128;         // pixels per CTA in fwd / bwd-data
constexpr int NT = 288;         // 9 warps: warp w owns output (or input) channels 8w..8w+7
constexpr int TPW = 32;         // pixels per step in bwd-weight
constexpr int NTW = 352;        // 11 warps, 324 active lanes = 18 x 18 register tiles of 4 x 4

// Workspace layout (floats): [9][CP c][CP o] forward pack | [9][CP o][CP c] backward pack | [CP] bias
constexpr size_t WS_FWD = 0;
constexpr size_t WS_BWD = (size_t)9 * CP * CP;
constexpr size_t WS_BIAS = (size_t)2 * 9 * CP * CP;
constexpr size_t WS_FLOATS = WS_BIAS + CP;

struct DcnParams {
  const void* x; const void* offset; const void* mask; void* out;
  long long x_sn, x_sc, x_sh, x_sw;
  long long f_sn, f_sc, f_sh, f_sw;     // offset
  long long m_sn, m_sc, m_sh, m_sw;     // mask
  long long o_sn, o_sc, o_sh, o_sw;     // out (fwd) or grad_out (bwd)
  int B, C, O, H, W;
  const float* ws;                      // packed weights (+ bias)
  // backward only
  float* gx; float* goff; float* gmask;
  long long gx_sn, gx_sc, gx_sh, gx_sw;
  long long gf_sn, gf_sc, gf_sh, gf_sw;
  long long gm_sn, gm_sc, gm_sh, gm_sw;
  float* gw; float* gb;
};

// Per-(pixel, tap) sampling geometry, Appendix B.  Corner flags are kept even when the sample is dead because
// torchvision's offset gradient is not gated on liveness (matters only when py or px is exactly -1).
struct Geo {
  int off00, off01, off10, off11;   // clamped element offsets inside one x plane
  float lh, lw, mk;
  int flags;                        // bit0..3: corner 00,01,10,11 valid; bit4: sample live
  int y0, x0;                       // unclamped integer corner (for scatter targets with their own strides)
};

template <typename TO>
__device__ __forceinline__ Geo make_geo(const DcnParams& p, int b, int y, int x, int k) {
  Geo g;
  const TO* off = reinterpret_cast<const TO*>(p.offset) + b * p.f_sn + y * p.f_sh + x * p.f_sw;
  const TO* msk = reinterpret_cast<const TO*>(p.mask) + b * p.m_sn + y * p.m_sh + x * p.m_sw;
  float dy = to_f32<TO>(__ldg(off + (2 * k) * p.f_sc));
  float dx = to_f32<TO>(__ldg(off + (2 * k + 1) * p.f_sc));
  g.mk = to_f32<TO>(__ldg(msk + k * p.m_sc));
  float py = __fadd_rn((float)(y - 1 + k / 3), dy);
  float px = __fadd_rn((float)(x - 1 + k % 3), dx);
  bool live = (py > -1.0f) && (py < (float)p.H) && (px > -1.0f) && (px < (float)p.W);
  if (!(py > -2.0f && py < (float)p.H + 1.0f)) py = -2.0f;   // no valid corner out there; keeps the cast defined
  if (!(px > -2.0f && px < (float)p.W + 1.0f)) px = -2.0f;
  float fy = floorf(py), fx = floorf(px);
  int y0 = (int)fy, x0 = (int)fx;
  g.y0 = y0; g.x0 = x0;
  g.lh = py - fy;
  g.lw = px - fx;
  bool r0 = (unsigned)y0 < (unsigned)p.H, r1 = (unsigned)(y0 + 1) < (unsigned)p.H;
  bool c0 = (unsigned)x0 < (unsigned)p.W, c1 = (unsigned)(x0 + 1) < (unsigned)p.W;
  int cy0 = min(max(y0, 0), p.H - 1), cy1 = min(max(y0 + 1, 0), p.H - 1);
  int cx0 = min(max(x0, 0), p.W - 1), cx1 = min(max(x0 + 1, 0), p.W - 1);
  g.off00 = (int)(cy0 * p.x_sh + cx0 * p.x_sw);
  g.off01 = (int)(cy0 * p.x_sh + cx1 * p.x_sw);
  g.off10 = (int)(cy1 * p.x_sh + cx0 * p.x_sw);
  g.off11 = (int)(cy1 * p.x_sh + cx1 * p.x_sw);
  g.flags = (r0 && c0 ? 1 : 0) | (r0 && c1 ? 2 : 0) | (r1 && c0 ? 4 : 0) | (r1 && c1 ? 8 : 0) | (live ? 16 : 0);
  return g;
}

template <typename TX>
__device__ __forceinline__ void load_corners(const TX* plane, const Geo& g, float& v00, float& v01, float& v10, float& v11) {
  v00 = (g.flags & 1) ? ldg_f32(plane + g.off00) : 0.0f;
  v01 = (g.flags & 2) ? ldg_f32(plane + g.off01) : 0.0f;
  v10 = (g.flags & 4) ? ldg_f32(plane + g.off10) : 0.0f;
  v11 = (g.flags & 8) ? ldg_f32(plane + g.off11) : 0.0f;
}

// torchvision's bilinear_interpolate: (w1*v1 + w2*v2 + w3*v3 + w4*v4), weights from (1-lh),(1-lw),lh,lw.
__device__ __forceinline__ float bilerp(const Geo& g, float v00, float v01, float v10, float v11) {
  float hh = 1.0f - g.lh, hw = 1.0f - g.lw;
  float w00 = hh * hw, w01 = hh * g.lw, w10 = g.lh * hw, w11 = g.lh * g.lw;
  float acc = w00 * v00;
  acc = fmaf(w01, v01, acc);
  acc = fmaf(w10, v10, acc);
  acc = fmaf(w11, v11, acc);
  return (g.flags & 16) ? acc : 0.0f;
}

// ---------------------------------------------------------------------------------------------- weight pack
template <typename TW>
__global__ void dcn_pack_simt_kernel(const TW* __restrict__ w, const void* bias, int bias_dtype, int O, int C, float* ws) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = 9 * CP * CP;
  if (idx < per) {
    int k = idx / (CP * CP), r = (idx / CP) % CP, q = idx % CP;
    // forward: [k][c=r][o=q]; backward: [k][o=r][c=q]
    ws[WS_FWD + idx] = (r < C && q < O) ? to_f32<TW>(w[((size_t)q * C + r) * 9 + k]) : 0.0f;
    ws[WS_BWD + idx] = (r < O && q < C) ? to_f32<TW>(w[((size_t)r * C + q) * 9 + k]) : 0.0f;
  }
  if (idx < CP) {
    float bv = 0.0f;
    if (bias && idx < O) {
      if (bias_dtype == VFI_F32) bv = reinterpret_cast<const float*>(bias)[idx];
      else if (bias_dtype == VFI_BF16) bv = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(bias)[idx]);
      else bv = __half2float(reinterpret_cast<const __half*>(bias)[idx]);
    }
    ws[WS_BIAS + idx] = bv;
  }
}

// ---------------------------------------------------------------------------------------------- forward
struct FwdSmem {
  float col[CP][TP];     // modulated samples of the current tap, [c][pixel]
  float wts[CP][CP];     // Wk[c][o]
  Geo geo[TP];
};

template <typename TX, typename TO>
__global__ void __launch_bounds__(NT) dcn_fwd_simt_kernel(const DcnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& s = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const long long HW = (long long)p.H * p.W;
  const long long p0 = (long long)blockIdx.x * TP;
  const TX* xb = reinterpret_cast<const TX*>(p.x) + b * p.x_sn;

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  for (int k = 0; k < 9; ++k) {
    __syncthreads();   // previous tap's compute is done with col / wts / geo
    if (tid < TP) {
      long long pp = p0 + tid;
      if (pp < HW) s.geo[tid] = make_geo<TO>(p, b, (int)(pp / p.W), (int)(pp % p.W), k);
      else { Geo g; g.off00 = g.off01 = g.off10 = g.off11 = 0; g.lh = g.lw = g.mk = 0.0f; g.flags = 0; g.y0 = g.x0 = 0; s.geo[tid] = g; }
    }
    {
      const float4* src = reinterpret_cast<const float4*>(p.ws + WS_FWD + (size_t)k * CP * CP);
      float4* dst = reinterpret_cast<float4*>(&s.wts[0][0]);
      for (int i = tid; i < CP * CP / 4; i += NT) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if (tid < 2 * TP) {
      const int px = tid & (TP - 1), crow = tid >> 7;
      const Geo g = s.geo[px];
      for (int c = crow; c < CP; c += 2) {
        float val = 0.0f;
        if (c < p.C) {
          float v00, v01, v10, v11;
          load_corners<TX>(xb + c * p.x_sc, g, v00, v01, v10, v11);
          val = g.mk * bilerp(g, v00, v01, v10, v11);
        }
        s.col[c][px] = val;
      }
    }
    __syncthreads();
    const int cmax = p.C;
#pragma unroll 4
    for (int c = 0; c < cmax; ++c) {
      const float4 a = *reinterpret_cast<const float4*>(&s.col[c][4 * lane]);
      const float4 w0 = *reinterpret_cast<const float4*>(&s.wts[c][8 * warp]);
      const float4 w1 = *reinterpret_cast<const float4*>(&s.wts[c][8 * warp + 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
  }

  // epilogue: + bias, store.  Lane owns pixels p0 + 4*lane .. +3, warp owns channels 8*warp .. +7.
  TX* ob = reinterpret_cast<TX*>(p.out) + b * p.o_sn;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int o = 8 * warp + j;
    if (o >= p.O) break;
    const float bv = p.ws[WS_BIAS + o];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      long long pp = p0 + 4 * lane + i;
      if (pp < HW) {
        int y = (int)(pp / p.W), x = (int)(pp % p.W);
        ob[o * p.o_sc + y * p.o_sh + x * p.o_sw] = from_f32<TX>(acc[i][j] + bv);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward: data
struct BwdSmem {
  float gout[CP][TP];    // grad_out tile [o][pixel] (rows >= O are zero)
  float gcol[CP][TP];    // Wk^T * gout for the current tap, [c][pixel]
  float wts[CP][CP];     // Wk[o][c]
  Geo geo[TP];
  float part[2][3][TP];  // partial (mask, dy, dx) sums of the two channel-interleaved halves
};

template <typename TX, typename TO, typename TG>
__global__ void __launch_bounds__(NT) dcn_bwd_data_simt_kernel(const DcnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& s = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const long long HW = (long long)p.H * p.W;
  const long long p0 = (long long)blockIdx.x * TP;
  const TX* xb = reinterpret_cast<const TX*>(p.x) + b * p.x_sn;
  const TG* gob = reinterpret_cast<const TG*>(p.out) + b * p.o_sn;

  for (int i = tid; i < CP * TP; i += NT) {
    int o = i / TP, px = i % TP;
    long long pp = p0 + px;
    float v = 0.0f;
    if (o < p.O && pp < HW) v = to_f32<TG>(__ldg(gob + o * p.o_sc + (pp / p.W) * p.o_sh + (pp % p.W) * p.o_sw));
    s.gout[o][px] = v;
  }

  for (int k = 0; k < 9; ++k) {
    __syncthreads();
    if (tid < TP) {
      long long pp = p0 + tid;
      if (pp < HW) s.geo[tid] = make_geo<TO>(p, b, (int)(pp / p.W), (int)(pp % p.W), k);
      else { Geo g; g.off00 = g.off01 = g.off10 = g.off11 = 0; g.lh = g.lw = g.mk = 0.0f; g.flags = 0; g.y0 = g.x0 = 0; s.geo[tid] = g; }
    }
    {
      const float4* src = reinterpret_cast<const float4*>(p.ws + WS_BWD + (size_t)k * CP * CP);
      float4* dst = reinterpret_cast<float4*>(&s.wts[0][0]);
      for (int i = tid; i < CP * CP / 4; i += NT) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    {
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
      const int omax = p.O;
#pragma unroll 4
      for (int o = 0; o < omax; ++o) {
        const float4 a = *reinterpret_cast<const float4*>(&s.gout[o][4 * lane]);
        const float4 w0 = *reinterpret_cast<const float4*>(&s.wts[o][8 * warp]);
        const float4 w1 = *reinterpret_cast<const float4*>(&s.wts[o][8 * warp + 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(&s.gcol[8 * warp + j][4 * lane]) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
    }
    __syncthreads();
    if (tid < 2 * TP) {
      const int px = tid & (TP - 1), crow = tid >> 7;
      const Geo g = s.geo[px];
      const float hh = 1.0f - g.lh, hw = 1.0f - g.lw;
      const bool live = (g.flags & 16) != 0;
      float g_m = 0.0f, g_dy = 0.0f, g_dx = 0.0f;
      const int y0 = g.y0, x0 = g.x0;
      for (int c = crow; c < p.C; c += 2) {
        const float gc = s.gcol[c][px];
        float v00, v01, v10, v11;
        load_corners<TX>(xb + c * p.x_sc, g, v00, v01, v10, v11);
        const float val = bilerp(g, v00, v01, v10, v11);
        g_m = fmaf(gc, val, g_m);
        const float d_py = g.lw * (v11 - v01) + hw * (v10 - v00);
        const float d_px = g.lh * (v11 - v10) + hh * (v01 - v00);
        const float gm = gc * g.mk;
        g_dy = fmaf(gm, d_py, g_dy);
        g_dx = fmaf(gm, d_px, g_dx);
        if (p.gx && live) {
          float* gp = p.gx + b * p.gx_sn + c * p.gx_sc;
          if (g.flags & 1) atomicAdd(gp + y0 * p.gx_sh + x0 * p.gx_sw, gm * (hh * hw));
          if (g.flags & 2) atomicAdd(gp + y0 * p.gx_sh + (x0 + 1) * p.gx_sw, gm * (hh * g.lw));
          if (g.flags & 4) atomicAdd(gp + (y0 + 1) * p.gx_sh + x0 * p.gx_sw, gm * (g.lh * hw));
          if (g.flags & 8) atomicAdd(gp + (y0 + 1) * p.gx_sh + (x0 + 1) * p.gx_sw, gm * (g.lh * g.lw));
        }
      }
      s.part[crow][0][px] = g_m; s.part[crow][1][px] = g_dy; s.part[crow][2][px] = g_dx;
    }
    __syncthreads();
    if (tid < TP) {
      long long pp = p0 + tid;
      if (pp < HW) {
        int y = (int)(pp / p.W), x = (int)(pp % p.W);
        if (p.gmask) p.gmask[b * p.gm_sn + k * p.gm_sc + y * p.gm_sh + x * p.gm_sw] = s.part[0][0][tid] + s.part[1][0][tid];
        if (p.goff) {
          float* go = p.goff + b * p.gf_sn + y * p.gf_sh + x * p.gf_sw;
          go[(2 * k) * p.gf_sc] = s.part[0][1][tid] + s.part[1][1][tid];
          go[(2 * k + 1) * p.gf_sc] = s.part[0][2][tid] + s.part[1][2][tid];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward: weight
struct WgtSmem {
  float gout[CP][TPW + 1];
  float col[CP][TPW + 1];
  Geo geo[TPW];
};

template <typename TX, typename TO, typename TG>
__global__ void __launch_bounds__(NTW) dcn_bwd_weight_simt_kernel(const DcnParams p, int tiles_per_image) {
  __shared__ WgtSmem s;
  const int tid = threadIdx.x;
  const int k = blockIdx.y;
  const long long HW = (long long)p.H * p.W;
  const bool active = tid < 18 * 18;
  const int to = tid / 18, tc = tid % 18;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  float bsum = 0.0f;

  const long long total = (long long)p.B * tiles_per_image;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int b = (int)(t / tiles_per_image);
    const long long p0 = (t % tiles_per_image) * TPW;
    const TX* xb = reinterpret_cast<const TX*>(p.x) + b * p.x_sn;
    const TG* gob = reinterpret_cast<const TG*>(p.out) + b * p.o_sn;
    __syncthreads();
    if (tid < TPW) {
      long long pp = p0 + tid;
      if (pp < HW) s.geo[tid] = make_geo<TO>(p, b, (int)(pp / p.W), (int)(pp % p.W), k);
      else { Geo g; g.off00 = g.off01 = g.off10 = g.off11 = 0; g.lh = g.lw = g.mk = 0.0f; g.flags = 0; g.y0 = g.x0 = 0; s.geo[tid] = g; }
    }
    for (int i = tid; i < CP * TPW; i += NTW) {
      int o = i / TPW, px = i % TPW;
      long long pp = p0 + px;
      float v = 0.0f;
      if (o < p.O && pp < HW) v = to_f32<TG>(__ldg(gob + o * p.o_sc + (pp / p.W) * p.o_sh + (pp % p.W) * p.o_sw));
      s.gout[o][px] = v;
    }
    __syncthreads();
    for (int i = tid; i < CP * TPW; i += NTW) {
      int c = i / TPW, px = i % TPW;
      float val = 0.0f;
      if (c < p.C) {
        const Geo g = s.geo[px];
        float v00, v01, v10, v11;
        load_corners<TX>(xb + c * p.x_sc, g, v00, v01, v10, v11);
        val = g.mk * bilerp(g, v00, v01, v10, v11);
      }
      s.col[c][px] = val;
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int px = 0; px < TPW; ++px) {
        float a[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = s.gout[to + 18 * i][px]; bv[i] = s.col[tc + 18 * i][px]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
      }
    }
    if (k == 0 && p.gb && tid < p.O) {
      float sacc = 0.0f;
      for (int px = 0; px < TPW; ++px) sacc += s.gout[tid][px];
      bsum += sacc;
    }
  }
  if (active && p.gw) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int o = to + 18 * i;
      if (o >= p.O) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = tc + 18 * j;
        if (c < p.C) atomicAdd(p.gw + ((size_t)o * p.C + c) * 9 + k, acc[i][j]);
      }
    }
  }
  if (k == 0 && p.gb && tid < p.O) atomicAdd(p.gb + tid, bsum);
}

// ---------------------------------------------------------------------------------------------- host side
int check_dcn(const vfi_tensor* x, const vfi_tensor* offset, const vfi_tensor* mask, const vfi_tensor* out, long long O,
              const char* who) {
  VFI_REQUIRE(x && offset && mask && out, VFI_ERR_INVALID, "%s: null tensor descriptor", who);
  VFI_REQUIRE(x->data && offset->data && mask->data && out->data, VFI_ERR_INVALID, "%s: null data pointer", who);
  VFI_REQUIRE(x->n >= 0 && x->c > 0 && x->h >= 0 && x->w >= 0 && O > 0, VFI_ERR_INVALID, "%s: bad extent", who);
  VFI_REQUIRE(offset->n == x->n && offset->c == 18 && offset->h == x->h && offset->w == x->w, VFI_ERR_INVALID,
              "%s: offset must be [B,18,H,W] (3x3 kernel, one offset group); got [%lld,%lld,%lld,%lld]", who,
              (long long)offset->n, (long long)offset->c, (long long)offset->h, (long long)offset->w);
  VFI_REQUIRE(mask->n == x->n && mask->c == 9 && mask->h == x->h && mask->w == x->w, VFI_ERR_INVALID,
              "%s: mask must be [B,9,H,W]", who);
  VFI_REQUIRE(out->n == x->n && out->c == O && out->h == x->h && out->w == x->w, VFI_ERR_INVALID,
              "%s: out/grad_out must be [B,O,H,W]", who);
  VFI_REQUIRE(offset->dtype == mask->dtype, VFI_ERR_UNSUPPORTED, "%s: offset and mask must share a dtype", who);
  VFI_REQUIRE(x->c <= CP && O <= CP, VFI_ERR_UNSUPPORTED,
              "%s: fp32 path supports at most %d input/output channels (got C=%lld, O=%lld)", who, CP,
              (long long)x->c, O);
  VFI_REQUIRE((x->h - 1) * llabs(x->sh) + (x->w - 1) * llabs(x->sw) < 2147483647LL, VFI_ERR_UNSUPPORTED,
              "%s: one input plane must span < 2^31 elements", who);
  VFI_REQUIRE(x->n <= 65535, VFI_ERR_UNSUPPORTED, "%s: batch > 65535", who);
  return VFI_OK;
}

void fill_common(DcnParams& p, const vfi_tensor* x, const vfi_tensor* offset, const vfi_tensor* mask,
                 const vfi_tensor* out, long long O, const float* ws) {
  p = DcnParams{};
  p.x = x->data; p.offset = offset->data; p.mask = mask->data; p.out = out->data;
  p.x_sn = x->sn; p.x_sc = x->sc; p.x_sh = x->sh; p.x_sw = x->sw;
  p.f_sn = offset->sn; p.f_sc = offset->sc; p.f_sh = offset->sh; p.f_sw = offset->sw;
  p.m_sn = mask->sn; p.m_sc = mask->sc; p.m_sh = mask->sh; p.m_sw = mask->sw;
  p.o_sn = out->sn; p.o_sc = out->sc; p.o_sh = out->sh; p.o_sw = out->sw;
  p.B = (int)x->n; p.C = (int)x->c; p.O = (int)O; p.H = (int)x->h; p.W = (int)x->w;
  p.ws = ws;
}

}  // namespace

size_t dcn_simt_workspace_bytes() { return ((WS_FLOATS * sizeof(float) + 255) / 256) * 256; }

int dcn_simt_pack(const void* weight, int weight_dtype, const void* bias, int bias_dtype, long long O, long long C,
                  float* ws, cudaStream_t st) {
  const int n = 9 * CP * CP;
  VFI_DISPATCH(weight_dtype, TW, {
    dcn_pack_simt_kernel<TW><<<ceil_div(n, 256), 256, 0, st>>>(reinterpret_cast<const TW*>(weight), bias, bias_dtype,
                                                             (int)O, (int)C, ws);
  });
  VFI_LAUNCH_CHECK("dcn_pack_simt_kernel");
  return VFI_OK;
}

int dcn_simt_fwd(const vfi_tensor* x, const vfi_tensor* offset, const vfi_tensor* mask, const void* weight,
                 int weight_dtype, const void* bias, int bias_dtype, const vfi_tensor* out, long long O, void* workspace,
                 size_t workspace_bytes, cudaStream_t st) {
  int rc = check_dcn(x, offset, mask, out, O, "vfi_dcn_fwd");
  if (rc) return rc;
  VFI_REQUIRE(weight, VFI_ERR_INVALID, "vfi_dcn_fwd: null weight");
  VFI_REQUIRE(x->dtype == out->dtype, VFI_ERR_UNSUPPORTED, "vfi_dcn_fwd: x/out dtype mismatch");
  VFI_REQUIRE(workspace && workspace_bytes >= dcn_simt_workspace_bytes() && aligned(workspace, 16), VFI_ERR_WORKSPACE,
              "vfi_dcn_fwd: workspace of %zu bytes (16-byte aligned) required, got %zu", dcn_simt_workspace_bytes(),
              workspace_bytes);
  if (x->n == 0 || x->h == 0 || x->w == 0) return VFI_OK;
  float* ws = reinterpret_cast<float*>(workspace);
  rc = dcn_simt_pack(weight, weight_dtype, bias, bias_dtype, O, x->c, ws, st);
  if (rc) return rc;
  DcnParams p;
  fill_common(p, x, offset, mask, out, O, ws);
  dim3 grid(ceil_div((long long)x->h * x->w, TP), (unsigned)x->n);
  const size_t smem = sizeof(FwdSmem);
  VFI_DISPATCH(x->dtype, TX, {
    VFI_DISPATCH(offset->dtype, TO, {
      auto kern = dcn_fwd_simt_kernel<TX, TO>;
      VFI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, NT, smem, st>>>(p);
    });
  });
  VFI_LAUNCH_CHECK("dcn_fwd_simt_kernel");
  return VFI_OK;
}

}  // namespace vfi

using namespace vfi;

extern "C" int vfi_dcn_bwd_data(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* offset,
                                const vfi_tensor* mask, const void* weight, int32_t weight_dtype, int64_t O,
                                const vfi_tensor* grad_x, const vfi_tensor* grad_offset, const vfi_tensor* grad_mask,
                                void* workspace, size_t workspace_bytes, vfi_stream_t stream) {
  int rc = check_dcn(x, offset, mask, grad_out, O, "vfi_dcn_bwd_data");
  if (rc) return rc;
  VFI_REQUIRE(grad_out->dtype == x->dtype, VFI_ERR_UNSUPPORTED, "vfi_dcn_bwd_data: grad_out must have x's dtype");
  VFI_REQUIRE(weight, VFI_ERR_INVALID, "vfi_dcn_bwd_data: null weight");
  VFI_REQUIRE(workspace && workspace_bytes >= dcn_simt_workspace_bytes() && aligned(workspace, 16), VFI_ERR_WORKSPACE,
              "vfi_dcn_bwd_data: workspace of %zu bytes (16-byte aligned) required, got %zu",
              dcn_simt_workspace_bytes(), workspace_bytes);
  if (grad_x) VFI_REQUIRE(grad_x->data && grad_x->dtype == VFI_F32 && same_shape(grad_x, x), VFI_ERR_INVALID,
                          "vfi_dcn_bwd_data: grad_x must be f32 with x's shape");
  if (grad_offset) VFI_REQUIRE(grad_offset->data && grad_offset->dtype == VFI_F32 && same_shape(grad_offset, offset),
                               VFI_ERR_INVALID, "vfi_dcn_bwd_data: grad_offset must be f32 [B,18,H,W]");
  if (grad_mask) VFI_REQUIRE(grad_mask->data && grad_mask->dtype == VFI_F32 && same_shape(grad_mask, mask),
                             VFI_ERR_INVALID, "vfi_dcn_bwd_data: grad_mask must be f32 [B,9,H,W]");
  if (x->n == 0 || x->h == 0 || x->w == 0) return VFI_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = reinterpret_cast<float*>(workspace);
  rc = dcn_simt_pack(weight, weight_dtype, nullptr, VFI_F32, O, x->c, ws, st);
  if (rc) return rc;
  DcnParams p;
  fill_common(p, x, offset, mask, grad_out, O, ws);
  if (grad_x) { p.gx = (float*)grad_x->data; p.gx_sn = grad_x->sn; p.gx_sc = grad_x->sc; p.gx_sh = grad_x->sh; p.gx_sw = grad_x->sw; }
  if (grad_offset) { p.goff = (float*)grad_offset->data; p.gf_sn = grad_offset->sn; p.gf_sc = grad_offset->sc; p.gf_sh = grad_offset->sh; p.gf_sw = grad_offset->sw; }
  if (grad_mask) { p.gmask = (float*)grad_mask->data; p.gm_sn = grad_mask->sn; p.gm_sc = grad_mask->sc; p.gm_sh = grad_mask->sh; p.gm_sw = grad_mask->sw; }
  dim3 grid(ceil_div((long long)x->h * x->w, TP), (unsigned)x->n);
  const size_t smem = sizeof(BwdSmem);
  VFI_DISPATCH(x->dtype, TX, {
    VFI_DISPATCH(offset->dtype, TO, {
      auto kern = dcn_bwd_data_simt_kernel<TX, TO, TX>;
      VFI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, NT, smem, st>>>(p);
    });
  });
  VFI_LAUNCH_CHECK("dcn_bwd_data_simt_kernel");
  return VFI_OK;
}

extern "C" int vfi_dcn_bwd_weight(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* offset,
                                  const vfi_tensor* mask, int64_t O, float* grad_weight, float* grad_bias,
                                  vfi_stream_t stream) {
  int rc = check_dcn(x, offset, mask, grad_out, O, "vfi_dcn_bwd_weight");
  if (rc) return rc;
  VFI_REQUIRE(grad_out->dtype == x->dtype, VFI_ERR_UNSUPPORTED, "vfi_dcn_bwd_weight: grad_out must have x's dtype");
  if (x->n == 0 || x->h == 0 || x->w == 0 || (!grad_weight && !grad_bias)) return VFI_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DcnParams p;
  fill_common(p, x, offset, mask, grad_out, O, nullptr);
  p.gw = grad_weight; p.gb = grad_bias;
  const int tiles_per_image = ceil_div((long long)x->h * x->w, TPW);
  const long long total = (long long)x->n * tiles_per_image;
  dim3 grid((unsigned)(total < 296 ? total : 296), 9);
  VFI_DISPATCH(x->dtype, TX, {
    VFI_DISPATCH(offset->dtype, TO, {
      dcn_bwd_weight_simt_kernel<TX, TO, TX><<<grid, NTW, 0, st>>>(p, tiles_per_image);
    });
  });
  VFI_LAUNCH_CHECK("dcn_bwd_weight_simt_kernel");
  return VFI_OK;
}
