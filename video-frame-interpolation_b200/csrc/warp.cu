// warp.cu -- flow-guided backward warp (K1 forward, K2 backward, W3 blend extension) for sm_100a.
//
// Replaces /root/reference/src/models/ema_vfi.py:149-171: the reference builds a pixel grid on the CPU, copies it
// to the device, adds the flow, normalises, permutes and calls F.grid_sample.  Here the grid never exists: every
// thread derives its coordinates from its index, replays the normalise/un-normalise round trip bit-exactly
// (warp_math.h) and gathers the four corners straight from the source planes.
//
// HBM-bound: algorithmic traffic is (C_in + 2 + C_out) elements per pixel (32 B/px fp32, 16 B/px bf16, C = 3).
// The vectorised kernel moves flow and output as 128-bit (fp32) / 64-bit (bf16) accesses, four pixels per thread;
// the corner gathers go through the read-only path and hit L1/L2 for everything but the compulsory first touch.
#include "common.cuh"
#include "warp_math.h"

namespace vfi {
namespace {

struct WarpParams {
  const void* src;
  const void* flow;
  void* out;
  long long s_sn, s_sc, s_sh, s_sw;   // src strides
  long long f_sn, f_sc, f_sh, f_sw;   // flow strides
  long long o_sn, o_sc, o_sh, o_sw;   // out strides
  int B, C, H, W;
  WarpAxis ax, ay;
};

// Everything a pixel needs to know about where it samples.
struct Corners {
  int off00, off01, off10, off11;   // element offsets inside one source plane (clamped, always in range)
  float w00, w01, w10, w11;         // nw, ne, sw, se weights
  bool v00, v01, v10, v11;          // corner lies inside the frame
  float wx0, wx1, wy0, wy1;
  int x0, y0;                       // unclamped north-west corner
};

__device__ __forceinline__ Corners locate(int x, int y, float fx, float fy, const WarpAxis& ax, const WarpAxis& ay,
                                          int H, int W, long long sh, long long sw) {
  Corners c;
  float ix = vfi_warp_coord(x, fx, ax);
  float iy = vfi_warp_coord(y, fy, ay);
  float x0f = floorf(ix), y0f = floorf(iy);
  int x0 = (int)x0f, y0 = (int)y0f;
  c.x0 = x0; c.y0 = y0;
  c.wx1 = ix - x0f;
  c.wx0 = (x0f + 1.0f) - ix;
  c.wy1 = iy - y0f;
  c.wy0 = (y0f + 1.0f) - iy;
  bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
  bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
  int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
  int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
  c.off00 = (int)(cy0 * sh + cx0 * sw);
  c.off01 = (int)(cy0 * sh + cx1 * sw);
  c.off10 = (int)(cy1 * sh + cx0 * sw);
  c.off11 = (int)(cy1 * sh + cx1 * sw);
  c.v00 = vx0 && vy0; c.v01 = vx1 && vy0; c.v10 = vx0 && vy1; c.v11 = vx1 && vy1;
  c.w00 = c.wx0 * c.wy0; c.w01 = c.wx1 * c.wy0; c.w10 = c.wx0 * c.wy1; c.w11 = c.wx1 * c.wy1;
  return c;
}

template <typename TS>
__device__ __forceinline__ float sample(const TS* plane, const Corners& c) {
  // predicated loads: an out-of-frame corner contributes nothing (zeros padding), as in aten::grid_sampler_2d
  float a = c.v00 ? ldg_f32(plane + c.off00) : 0.0f;
  float b = c.v01 ? ldg_f32(plane + c.off01) : 0.0f;
  float d = c.v10 ? ldg_f32(plane + c.off10) : 0.0f;
  float e = c.v11 ? ldg_f32(plane + c.off11) : 0.0f;
  float acc = a * c.w00;
  acc = fmaf(b, c.w01, acc);
  acc = fmaf(d, c.w10, acc);
  acc = fmaf(e, c.w11, acc);
  return acc;
}

// ------------------------------------------------------------------------------------------------ forward
// One thread = VEC consecutive pixels of one row.  VEC = 4 requires unit W-stride and 4-element alignment of flow
// and out rows (checked on the host); VEC = 1 is the fully strided path.  CT = compile-time channel count (0 = any).
template <typename TS, typename TF, int VEC, int CT>
__global__ void __launch_bounds__(256) warp_fwd_kernel(const WarpParams p) {
  const int Wv = (p.W + VEC - 1) / VEC;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)p.B * p.H * Wv;
  if (idx >= total) return;
  int xv = (int)(idx % Wv);
  long long t = idx / Wv;
  int y = (int)(t % p.H);
  int b = (int)(t / p.H);
  int x = xv * VEC;
  const int C = CT ? CT : p.C;

  const TF* fl = reinterpret_cast<const TF*>(p.flow) + b * p.f_sn + y * p.f_sh + x * p.f_sw;
  const TS* src = reinterpret_cast<const TS*>(p.src) + b * p.s_sn;
  TS* out = reinterpret_cast<TS*>(p.out) + b * p.o_sn + y * p.o_sh + x * p.o_sw;

  float fx[VEC], fy[VEC];
  if constexpr (VEC == 4) {
    if constexpr (sizeof(TF) == 4) {
      float4 a = __ldcs(reinterpret_cast<const float4*>(fl));
      float4 c = __ldcs(reinterpret_cast<const float4*>(fl + p.f_sc));
      fx[0] = a.x; fx[1] = a.y; fx[2] = a.z; fx[3] = a.w;
      fy[0] = c.x; fy[1] = c.y; fy[2] = c.z; fy[3] = c.w;
    } else {
      uint2 a = __ldcs(reinterpret_cast<const uint2*>(fl));
      uint2 c = __ldcs(reinterpret_cast<const uint2*>(fl + p.f_sc));
      const TF* ap = reinterpret_cast<const TF*>(&a);
      const TF* cp = reinterpret_cast<const TF*>(&c);
#pragma unroll
      for (int i = 0; i < 4; ++i) { fx[i] = to_f32<TF>(ap[i]); fy[i] = to_f32<TF>(cp[i]); }
    }
  } else {
    fx[0] = to_f32<TF>(fl[0]);
    fy[0] = to_f32<TF>(fl[p.f_sc]);
  }

  Corners cr[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) cr[i] = locate(x + i, y, fx[i], fy[i], p.ax, p.ay, p.H, p.W, p.s_sh, p.s_sw);

  auto do_channel = [&](int cc) {
    const TS* plane = src + cc * p.s_sc;
    float r[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) r[i] = sample<TS>(plane, cr[i]);
    TS* o = out + cc * p.o_sc;
    if constexpr (VEC == 4) {
      if constexpr (sizeof(TS) == 4) {
        __stcs(reinterpret_cast<float4*>(o), make_float4(r[0], r[1], r[2], r[3]));
      } else {
        TS v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = from_f32<TS>(r[i]);
        __stcs(reinterpret_cast<uint2*>(o), *reinterpret_cast<uint2*>(v));
      }
    } else {
      o[0] = from_f32<TS>(r[0]);
    }
  };
  if constexpr (CT > 0) {
#pragma unroll
    for (int c = 0; c < CT; ++c) do_channel(c);
  } else {
    for (int c = 0; c < C; ++c) do_channel(c);
  }
}

// Record output: out is a channels-last bf16 "tail plane" (unit channel stride, 8 elements = 16 B per pixel, dense
// rows), i.e. the tail input of the tensor-core DCN kernel.  One thread = 4 consecutive pixels: two 8-byte flow loads,
// 4 x C x 4 independent gathers in flight, four 16-byte record stores (channels >= C written as zeros, so the buffer needs
// no pre-clearing).  Removes the torch.cat of ema_vfi.py:134 without paying for partial-sector writes.
template <typename TF, int CT>
__global__ void __launch_bounds__(256) warp_fwd_rec_kernel(const WarpParams p) {
  const int Wv = p.W / 4;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)p.B * p.H * Wv;
  if (idx >= total) return;
  int xv = (int)(idx % Wv);
  long long t = idx / Wv;
  int y = (int)(t % p.H);
  int b = (int)(t / p.H);
  int x = xv * 4;
  const TF* fl = reinterpret_cast<const TF*>(p.flow) + b * p.f_sn + y * p.f_sh + x;
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(p.src) + b * p.s_sn;
  float fx[4], fy[4];
  if constexpr (sizeof(TF) == 4) {
    float4 a = __ldcs(reinterpret_cast<const float4*>(fl));
    float4 c = __ldcs(reinterpret_cast<const float4*>(fl + p.f_sc));
    fx[0] = a.x; fx[1] = a.y; fx[2] = a.z; fx[3] = a.w;
    fy[0] = c.x; fy[1] = c.y; fy[2] = c.z; fy[3] = c.w;
  } else {
    uint2 a = __ldcs(reinterpret_cast<const uint2*>(fl));
    uint2 c = __ldcs(reinterpret_cast<const uint2*>(fl + p.f_sc));
    const TF* ap = reinterpret_cast<const TF*>(&a);
    const TF* cp = reinterpret_cast<const TF*>(&c);
#pragma unroll
    for (int i = 0; i < 4; ++i) { fx[i] = to_f32<TF>(ap[i]); fy[i] = to_f32<TF>(cp[i]); }
  }
  Corners cr[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) cr[i] = locate(x + i, y, fx[i], fy[i], p.ax, p.ay, p.H, p.W, p.s_sh, p.s_sw);
  float r[4][CT];
#pragma unroll
  for (int c = 0; c < CT; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i][c] = sample<__nv_bfloat16>(src + c * p.s_sc, cr[i]);
  uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + b * p.o_sn + y * p.o_sh + x * p.o_sw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __align__(16) __nv_bfloat16 rec[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) rec[c] = __float2bfloat16_rn(c < CT ? r[i][c < CT ? c : 0] : 0.0f);
    __stcs(o + i, *reinterpret_cast<uint4*>(rec));
  }
}

// ------------------------------------------------------------------------------------------------ blend (W3)
struct BlendParams {
  WarpParams a;          // src_a / flow_a / out
  const void* src_b;
  const void* flow_b;
  const void* m;
  long long b_sn, b_sc, b_sh, b_sw;      // src_b strides
  long long g_sn, g_sc, g_sh, g_sw;      // flow_b strides
  long long m_sn, m_sh, m_sw;            // mask strides
};

template <typename TS, typename TF>
__global__ void __launch_bounds__(256) warp_blend_kernel(const BlendParams q) {
  const WarpParams& p = q.a;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)p.B * p.H * p.W;
  if (idx >= total) return;
  int x = (int)(idx % p.W);
  long long t = idx / p.W;
  int y = (int)(t % p.H);
  int b = (int)(t / p.H);
  const TF* fa = reinterpret_cast<const TF*>(p.flow) + b * p.f_sn + y * p.f_sh + x * p.f_sw;
  const TF* fb = reinterpret_cast<const TF*>(q.flow_b) + b * q.g_sn + y * q.g_sh + x * q.g_sw;
  Corners ca = locate(x, y, to_f32<TF>(fa[0]), to_f32<TF>(fa[p.f_sc]), p.ax, p.ay, p.H, p.W, p.s_sh, p.s_sw);
  Corners cb = locate(x, y, to_f32<TF>(fb[0]), to_f32<TF>(fb[q.g_sc]), p.ax, p.ay, p.H, p.W, q.b_sh, q.b_sw);
  float m = to_f32<TS>(reinterpret_cast<const TS*>(q.m)[b * q.m_sn + y * q.m_sh + x * q.m_sw]);
  float m1 = 1.0f - m;
  const TS* sa = reinterpret_cast<const TS*>(p.src) + b * p.s_sn;
  const TS* sb = reinterpret_cast<const TS*>(q.src_b) + b * q.b_sn;
  TS* out = reinterpret_cast<TS*>(p.out) + b * p.o_sn + y * p.o_sh + x * p.o_sw;
  for (int c = 0; c < p.C; ++c) {
    float wa = sample<TS>(sa + c * p.s_sc, ca);
    float wb = sample<TS>(sb + c * q.b_sc, cb);
    out[c * p.o_sc] = from_f32<TS>(__fadd_rn(__fmul_rn(m, wa), __fmul_rn(m1, wb)));
  }
}

// ------------------------------------------------------------------------------------------------ backward
struct WarpBwdParams {
  WarpParams f;               // src / flow as in forward; f.out unused
  const void* gout;
  float* gflow;
  float* gsrc;                // may be null
  long long go_sn, go_sc, go_sh, go_sw;
  long long gf_sn, gf_sc, gf_sh, gf_sw;
  long long gs_sn, gs_sc, gs_sh, gs_sw;
  float mult_x, mult_y;       // (W-1)/2, (H-1)/2: aten grid_sampler backward
};

template <typename TS, typename TF, typename TG>
__global__ void __launch_bounds__(256) warp_bwd_kernel(const WarpBwdParams q) {
  const WarpParams& p = q.f;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)p.B * p.H * p.W;
  if (idx >= total) return;
  int x = (int)(idx % p.W);
  long long t = idx / p.W;
  int y = (int)(t % p.H);
  int b = (int)(t / p.H);
  const TF* fl = reinterpret_cast<const TF*>(p.flow) + b * p.f_sn + y * p.f_sh + x * p.f_sw;
  Corners c = locate(x, y, to_f32<TF>(fl[0]), to_f32<TF>(fl[p.f_sc]), p.ax, p.ay, p.H, p.W, p.s_sh, p.s_sw);
  const TS* src = reinterpret_cast<const TS*>(p.src) + b * p.s_sn;
  const TG* go = reinterpret_cast<const TG*>(q.gout) + b * q.go_sn + y * q.go_sh + x * q.go_sw;
  float gix = 0.0f, giy = 0.0f;
  for (int ch = 0; ch < p.C; ++ch) {
    const TS* plane = src + ch * p.s_sc;
    float g = to_f32<TG>(__ldg(go + ch * q.go_sc));
    float nw = c.v00 ? ldg_f32(plane + c.off00) : 0.0f;
    float ne = c.v01 ? ldg_f32(plane + c.off01) : 0.0f;
    float sw = c.v10 ? ldg_f32(plane + c.off10) : 0.0f;
    float se = c.v11 ? ldg_f32(plane + c.off11) : 0.0f;
    gix += ((ne - nw) * c.wy0 + (se - sw) * c.wy1) * g;
    giy += ((sw - nw) * c.wx0 + (se - ne) * c.wx1) * g;
    if (q.gsrc) {
      // same pixel geometry as src but addressed with grad_src's own strides
      float* gs = q.gsrc + b * q.gs_sn + ch * q.gs_sc;
      const int x0 = c.x0, y0 = c.y0;
      if (c.v00) atomicAdd(gs + y0 * q.gs_sh + x0 * q.gs_sw, g * c.w00);
      if (c.v01) atomicAdd(gs + y0 * q.gs_sh + (x0 + 1) * q.gs_sw, g * c.w01);
      if (c.v10) atomicAdd(gs + (y0 + 1) * q.gs_sh + x0 * q.gs_sw, g * c.w10);
      if (c.v11) atomicAdd(gs + (y0 + 1) * q.gs_sh + (x0 + 1) * q.gs_sw, g * c.w11);
    }
  }
  // d(grid)/d(flow): aten scales by (size-1)/2, autograd of "2.0 * v / denom" divides by denom and doubles.
  float* gf = q.gflow + b * q.gf_sn + y * q.gf_sh + x * q.gf_sw;
  const float ggx = __fmul_rn(q.mult_x, gix), ggy = __fmul_rn(q.mult_y, giy);
  gf[0] = __fmul_rn(p.ax.recip ? __fmul_rn(ggx, p.ax.inv_denom) : __fdiv_rn(ggx, p.ax.denom), 2.0f);
  gf[q.gf_sc] = __fmul_rn(p.ay.recip ? __fmul_rn(ggy, p.ay.inv_denom) : __fdiv_rn(ggy, p.ay.denom), 2.0f);
}

// ------------------------------------------------------------------------------------------------ host side
int check_common(const vfi_tensor* src, const vfi_tensor* flow, const vfi_tensor* out, const char* who) {
  VFI_REQUIRE(src && flow && out, VFI_ERR_INVALID, "%s: null tensor descriptor", who);
  VFI_REQUIRE(src->n >= 0 && src->c >= 0 && src->h >= 0 && src->w >= 0, VFI_ERR_INVALID, "%s: negative extent", who);
  const bool empty = src->n == 0 || src->c == 0 || src->h == 0 || src->w == 0;
  VFI_REQUIRE(empty || (src->data && flow->data && out->data), VFI_ERR_INVALID, "%s: null data pointer", who);
  VFI_REQUIRE(same_shape(src, out), VFI_ERR_INVALID, "%s: out shape must equal src shape", who);
  VFI_REQUIRE(flow->n == src->n && flow->c == 2 && flow->h == src->h && flow->w == src->w, VFI_ERR_INVALID,
              "%s: flow must be [B,2,H,W] matching src (got [%lld,%lld,%lld,%lld])", who, (long long)flow->n,
              (long long)flow->c, (long long)flow->h, (long long)flow->w);
  VFI_REQUIRE(src->dtype == out->dtype, VFI_ERR_INVALID, "%s: src/out dtype mismatch", who);
  VFI_REQUIRE(flow->dtype == VFI_F32 || flow->dtype == src->dtype, VFI_ERR_UNSUPPORTED,
              "%s: flow dtype must be f32 or the dtype of src", who);
  VFI_REQUIRE(src->h < (1 << 23) && src->w < (1 << 23), VFI_ERR_UNSUPPORTED, "%s: frame too large", who);
  // plane offsets are held in 32-bit ints
  VFI_REQUIRE((src->h - 1) * llabs(src->sh) + (src->w - 1) * llabs(src->sw) < 2147483647LL, VFI_ERR_UNSUPPORTED,
              "%s: one source plane must span < 2^31 elements", who);
  return VFI_OK;
}

WarpParams make_params(const vfi_tensor* src, const vfi_tensor* flow, const vfi_tensor* out, int flags) {
  WarpParams p;
  p.src = src->data; p.flow = flow->data; p.out = out ? out->data : nullptr;
  p.s_sn = src->sn; p.s_sc = src->sc; p.s_sh = src->sh; p.s_sw = src->sw;
  p.f_sn = flow->sn; p.f_sc = flow->sc; p.f_sh = flow->sh; p.f_sw = flow->sw;
  if (out) { p.o_sn = out->sn; p.o_sc = out->sc; p.o_sh = out->sh; p.o_sw = out->sw; }
  else { p.o_sn = p.o_sc = p.o_sh = p.o_sw = 0; }
  p.B = (int)src->n; p.C = (int)src->c; p.H = (int)src->h; p.W = (int)src->w;
  p.ax = make_warp_axis(src->w, flags & VFI_WARP_DIV_RECIPROCAL);
  p.ay = make_warp_axis(src->h, flags & VFI_WARP_DIV_RECIPROCAL);
  return p;
}

bool rows_vec4(const vfi_tensor* t) {
  size_t es = dtype_size(t->dtype);
  return t->sw == 1 && t->w % 4 == 0 && t->sh % 4 == 0 && t->sc % 4 == 0 && t->sn % 4 == 0 && aligned(t->data, 4 * es);
}

template <typename TS, typename TF>
int launch_fwd(const WarpParams& p, bool vec4, cudaStream_t st) {
  if (vec4) {
    long long total = (long long)p.B * p.H * (p.W / 4);
    int blocks = ceil_div(total, 256);
    if (p.C == 3) warp_fwd_kernel<TS, TF, 4, 3><<<blocks, 256, 0, st>>>(p);
    else warp_fwd_kernel<TS, TF, 4, 0><<<blocks, 256, 0, st>>>(p);
  } else {
    long long total = (long long)p.B * p.H * p.W;
    int blocks = ceil_div(total, 256);
    if (p.C == 3) warp_fwd_kernel<TS, TF, 1, 3><<<blocks, 256, 0, st>>>(p);
    else warp_fwd_kernel<TS, TF, 1, 0><<<blocks, 256, 0, st>>>(p);
  }
  VFI_LAUNCH_CHECK("warp_fwd_kernel");
  return VFI_OK;
}

}  // namespace
}  // namespace vfi

using namespace vfi;

extern "C" int vfi_warp_fwd(const vfi_tensor* src, const vfi_tensor* flow, const vfi_tensor* out, int32_t flags,
                            vfi_stream_t stream) {
  int rc = check_common(src, flow, out, "vfi_warp_fwd");
  if (rc) return rc;
  if (src->n == 0 || src->c == 0 || src->h == 0 || src->w == 0) return VFI_OK;
  VFI_REQUIRE((long long)src->n * src->h * src->w < (1LL << 40), VFI_ERR_UNSUPPORTED, "vfi_warp_fwd: too many pixels");
  WarpParams p = make_params(src, flow, out, flags);
  bool vec4 = rows_vec4(flow) && rows_vec4(out);
  cudaStream_t st = (cudaStream_t)stream;
  // tail-plane record output (see warp_fwd_rec_kernel): [B,H,W,8] bf16 records, C = 3
  if (src->dtype == VFI_BF16 && src->c == 3 && out->sc == 1 && out->sw == 8 && out->sh == out->w * 8 && out->sn % 8 == 0 &&
      aligned(out->data, 16) && rows_vec4(flow)) {
    long long total = (long long)p.B * p.H * (p.W / 4);
    int blocks = ceil_div(total, 256);
    if (flow->dtype == VFI_F32) warp_fwd_rec_kernel<float, 3><<<blocks, 256, 0, st>>>(p);
    else warp_fwd_rec_kernel<__nv_bfloat16, 3><<<blocks, 256, 0, st>>>(p);
    VFI_LAUNCH_CHECK("warp_fwd_rec_kernel");
    return VFI_OK;
  }
  VFI_DISPATCH(src->dtype, TS, {
    if (flow->dtype == VFI_F32) { rc = launch_fwd<TS, float>(p, vec4, st); }
    else { rc = launch_fwd<TS, TS>(p, vec4, st); }
  });
  return rc;
}

extern "C" int vfi_warp_blend_fwd(const vfi_tensor* src_a, const vfi_tensor* flow_a, const vfi_tensor* src_b,
                                  const vfi_tensor* flow_b, const vfi_tensor* m, const vfi_tensor* out,
                                  int32_t flags, vfi_stream_t stream) {
  int rc = check_common(src_a, flow_a, out, "vfi_warp_blend_fwd");
  if (rc) return rc;
  rc = check_common(src_b, flow_b, out, "vfi_warp_blend_fwd");
  if (rc) return rc;
  VFI_REQUIRE(m && m->n == out->n && m->c == 1 && m->h == out->h && m->w == out->w, VFI_ERR_INVALID,
              "vfi_warp_blend_fwd: m must be [B,1,H,W]");
  VFI_REQUIRE(m->dtype == out->dtype && flow_a->dtype == flow_b->dtype, VFI_ERR_INVALID,
              "vfi_warp_blend_fwd: dtype mismatch");
  if (out->n == 0 || out->c == 0 || out->h == 0 || out->w == 0) return VFI_OK;
  BlendParams q;
  q.a = make_params(src_a, flow_a, out, flags);
  q.src_b = src_b->data; q.flow_b = flow_b->data; q.m = m->data;
  q.b_sn = src_b->sn; q.b_sc = src_b->sc; q.b_sh = src_b->sh; q.b_sw = src_b->sw;
  q.g_sn = flow_b->sn; q.g_sc = flow_b->sc; q.g_sh = flow_b->sh; q.g_sw = flow_b->sw;
  q.m_sn = m->sn; q.m_sh = m->sh; q.m_sw = m->sw;
  long long total = (long long)out->n * out->h * out->w;
  int blocks = ceil_div(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  VFI_DISPATCH(out->dtype, TS, {
    if (flow_a->dtype == VFI_F32) warp_blend_kernel<TS, float><<<blocks, 256, 0, st>>>(q);
    else warp_blend_kernel<TS, TS><<<blocks, 256, 0, st>>>(q);
  });
  VFI_LAUNCH_CHECK("warp_blend_kernel");
  return VFI_OK;
}

extern "C" int vfi_warp_bwd(const vfi_tensor* grad_out, const vfi_tensor* src, const vfi_tensor* flow,
                            const vfi_tensor* grad_flow, const vfi_tensor* grad_src, int32_t flags,
                            vfi_stream_t stream) {
  VFI_REQUIRE(grad_out && grad_flow && src && flow, VFI_ERR_INVALID, "vfi_warp_bwd: null tensor descriptor");
  const bool empty_b = src->n == 0 || src->c == 0 || src->h == 0 || src->w == 0;
  VFI_REQUIRE(empty_b || (grad_out->data && src->data && flow->data), VFI_ERR_INVALID, "vfi_warp_bwd: null data pointer");
  VFI_REQUIRE(same_shape(src, grad_out), VFI_ERR_INVALID, "vfi_warp_bwd: grad_out shape must equal src shape");
  VFI_REQUIRE(flow->n == src->n && flow->c == 2 && flow->h == src->h && flow->w == src->w, VFI_ERR_INVALID,
              "vfi_warp_bwd: flow must be [B,2,H,W] matching src");
  VFI_REQUIRE(flow->dtype == VFI_F32 || flow->dtype == src->dtype, VFI_ERR_UNSUPPORTED,
              "vfi_warp_bwd: flow dtype must be f32 or the dtype of src");
  VFI_REQUIRE((src->h - 1) * llabs(src->sh) + (src->w - 1) * llabs(src->sw) < 2147483647LL, VFI_ERR_UNSUPPORTED,
              "vfi_warp_bwd: one source plane must span < 2^31 elements");
  int rc = VFI_OK;
  VFI_REQUIRE(same_shape(grad_flow, flow) && grad_flow->dtype == VFI_F32 && (empty_b || grad_flow->data), VFI_ERR_INVALID,
              "vfi_warp_bwd: grad_flow must be f32 [B,2,H,W]");
  if (grad_src) {
    VFI_REQUIRE(same_shape(grad_src, src) && grad_src->dtype == VFI_F32 && grad_src->data, VFI_ERR_INVALID,
                "vfi_warp_bwd: grad_src must be f32 with src's shape");
  }
  if (src->n == 0 || src->h == 0 || src->w == 0) return VFI_OK;
  WarpBwdParams q;
  q.f = make_params(src, flow, nullptr, flags);
  q.gout = grad_out->data; q.gflow = (float*)grad_flow->data; q.gsrc = grad_src ? (float*)grad_src->data : nullptr;
  q.go_sn = grad_out->sn; q.go_sc = grad_out->sc; q.go_sh = grad_out->sh; q.go_sw = grad_out->sw;
  q.gf_sn = grad_flow->sn; q.gf_sc = grad_flow->sc; q.gf_sh = grad_flow->sh; q.gf_sw = grad_flow->sw;
  if (grad_src) { q.gs_sn = grad_src->sn; q.gs_sc = grad_src->sc; q.gs_sh = grad_src->sh; q.gs_sw = grad_src->sw; }
  else { q.gs_sn = q.gs_sc = q.gs_sh = q.gs_sw = 0; }
  q.mult_x = (float)(src->w - 1) / 2.0f;
  q.mult_y = (float)(src->h - 1) / 2.0f;
  long long total = (long long)src->n * src->h * src->w;
  int blocks = ceil_div(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  VFI_DISPATCH(src->dtype, TS, {
    VFI_DISPATCH(grad_out->dtype, TG, {
      if (flow->dtype == VFI_F32) warp_bwd_kernel<TS, float, TG><<<blocks, 256, 0, st>>>(q);
      else warp_bwd_kernel<TS, TS, TG><<<blocks, 256, 0, st>>>(q);
    });
  });
  VFI_LAUNCH_CHECK("warp_bwd_kernel");
  return VFI_OK;
}
