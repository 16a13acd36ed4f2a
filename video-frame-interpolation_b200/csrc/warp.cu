// warp.cu -- flow-guided backward warp (K1 forward, K2 backward, W3 blend extension) for sm_100a.
//
// Replaces /root/reference/src/models/ema_vfi.py:149-171: the reference builds a pixel grid on the CPU, copies it
// to the device, adds the flow, normalises, permutes and calls F.grid_sample.  Here the grid never exists: every
// thread derives its coordinates from its index, replays the normalise/un-normalise round trip bit-exactly
// (warp_math.h) and gathers the four corners straight from the source planes.
//
// HBM-bound by design: algorithmic traffic is (C_in + 2 + C_out) elements per pixel (32 B/px fp32, 16 B/px bf16, C = 3);
// the corner gathers go through the read-only path and hit L1/L2 for everything but the compulsory first touch.
#include <cstdlib>
#include <map>
#include <mutex>
#include <type_traits>
#include <utility>

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

// Everything a pixel needs to know about where it samples.  Plane offsets are 32-bit (checked on the host); a corner
// outside the frame has offset -1 and is not loaded (zeros padding).
struct Corners {
  int off00, off01, off10, off11;   // element offsets inside one source plane, -1 = corner outside the frame
  float w00, w01, w10, w11;         // nw, ne, sw, se weights
  float wx0, wx1, wy0, wy1;
  int x0, y0;                       // north-west corner (may lie outside)
};

__device__ __forceinline__ Corners locate(int x, int y, float fx, float fy, const WarpAxis& ax, const WarpAxis& ay,
                                          int H, int W, int sh, int sw) {
  Corners c;
  const float ix = vfi_warp_coord(x, fx, ax);
  const float iy = vfi_warp_coord(y, fy, ay);
  const int x0 = __float2int_rd(ix), y0 = __float2int_rd(iy);      // floor; positions are pinned to [-4, size + 4]
  const float x0f = (float)x0, y0f = (float)y0;
  c.x0 = x0; c.y0 = y0;
  c.wx1 = ix - x0f;
  c.wx0 = (x0f + 1.0f) - ix;
  c.wy1 = iy - y0f;
  c.wy0 = (y0f + 1.0f) - iy;
  const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
  const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
  const int r0 = y0 * sh, r1 = r0 + sh, c0 = x0 * sw, c1 = c0 + sw;
  c.off00 = (vx0 && vy0) ? r0 + c0 : -1;
  c.off01 = (vx1 && vy0) ? r0 + c1 : -1;
  c.off10 = (vx0 && vy1) ? r1 + c0 : -1;
  c.off11 = (vx1 && vy1) ? r1 + c1 : -1;
  c.w00 = c.wx0 * c.wy0; c.w01 = c.wx1 * c.wy0; c.w10 = c.wx0 * c.wy1; c.w11 = c.wx1 * c.wy1;
  return c;
}

template <typename TS>
__device__ __forceinline__ float corner(const TS* plane, int off) {
  return off >= 0 ? ldg_f32(plane + (unsigned)off) : 0.0f;   // predicated load: outside corners contribute nothing
}

template <typename TS>
__device__ __forceinline__ float sample(const TS* plane, const Corners& c) {
  const float a = corner(plane, c.off00), b = corner(plane, c.off01);
  const float d = corner(plane, c.off10), e = corner(plane, c.off11);
  float acc = a * c.w00;
  acc = fmaf(b, c.w01, acc);
  acc = fmaf(d, c.w10, acc);
  acc = fmaf(e, c.w11, acc);
  return acc;
}

// ------------------------------------------------------------------------------------------------ forward
// Thread mapping: lanes are CONSECUTIVE pixels of one row, so every one of the 4 x C corner gathers of a warp lands in
// one or two 128-byte lines (smooth flow) instead of the 3+ lines a 4-pixels-per-thread mapping touches; the LSU replays
// a multi-line request at ~2 cycles per line, which is what bounded the first version (0.23 ms for bf16 and fp32 alike).
// Instruction-level parallelism comes from PPT pixels per thread spaced one block apart (x, x + 256): all 2 x 4 x C
// gathers of a thread are independent and in flight together.  grid = (ceil(W / 512), H, B).
// REC = true: out is a channels-last bf16 "tail plane" (unit channel stride, 8 elements = 16 B per pixel) -- the tail
// input of the tensor-core DCN kernel: [c0 c1 c2 0 | c0 c1 c2 0], every byte written (no pre-clearing, no torch.cat).
constexpr int WARP_BLOCK = 256;
constexpr int WARP_PPT = 2;

template <typename TS, typename TF, int CT, bool REC>
__global__ void __launch_bounds__(WARP_BLOCK) warp_fwd_kernel(const WarpParams p) {
  const int y = blockIdx.y, b = blockIdx.z;
  const int x0 = blockIdx.x * (WARP_BLOCK * WARP_PPT) + threadIdx.x;
  const int C = CT ? CT : p.C;
  const TF* fl = reinterpret_cast<const TF*>(p.flow) + b * p.f_sn + y * p.f_sh;
  const TF* fl_y = fl + p.f_sc;
  const TS* src = reinterpret_cast<const TS*>(p.src) + b * p.s_sn;
  TS* out = reinterpret_cast<TS*>(p.out) + b * p.o_sn + y * p.o_sh;
  const int f_sw = (int)p.f_sw, o_sw = (int)p.o_sw, s_sh = (int)p.s_sh, s_sw = (int)p.s_sw;   // 32-bit inner strides

  Corners cr[WARP_PPT];
  bool ok[WARP_PPT];
#pragma unroll
  for (int i = 0; i < WARP_PPT; ++i) {
    const int x = x0 + i * WARP_BLOCK;
    ok[i] = x < p.W;
    const int xs = ok[i] ? x : 0;
    const float fx = to_f32<TF>(__ldcs(fl + xs * f_sw)), fy = to_f32<TF>(__ldcs(fl_y + xs * f_sw));
    cr[i] = locate(xs, y, fx, fy, p.ax, p.ay, p.H, p.W, s_sh, s_sw);
  }
  if constexpr (REC) {
    float r[WARP_PPT][CT ? CT : 1];
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
      for (int i = 0; i < WARP_PPT; ++i) r[i][c] = sample<TS>(src + c * p.s_sc, cr[i]);
#pragma unroll
    for (int i = 0; i < WARP_PPT; ++i) {
      if (!ok[i]) continue;
      __align__(16) __nv_bfloat16 rec[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) rec[c] = __float2bfloat16_rn((c & 3) < CT ? r[i][(c & 3) < CT ? (c & 3) : 0] : 0.0f);   // mirrored halves
      __stcs(reinterpret_cast<uint4*>(out + (x0 + i * WARP_BLOCK) * o_sw), *reinterpret_cast<uint4*>(rec));
    }
  } else {
    auto do_channel = [&](int c) {
      float r[WARP_PPT];
#pragma unroll
      for (int i = 0; i < WARP_PPT; ++i) r[i] = sample<TS>(src + c * p.s_sc, cr[i]);
#pragma unroll
      for (int i = 0; i < WARP_PPT; ++i)
        if (ok[i]) __stcs(out + c * p.o_sc + (x0 + i * WARP_BLOCK) * o_sw, from_f32<TS>(r[i]));
    };
    if constexpr (CT > 0) {
#pragma unroll
      for (int c = 0; c < CT; ++c) do_channel(c);
    } else {
      for (int c = 0; c < C; ++c) do_channel(c);
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward, fast path
// The form the hot path calls: C = 3 planar source with unit pixel stride (NCHW frames), flow with unit pixel stride, out
// planar or tail-plane records.  The generic kernel above spends ~250 instructions per pixel on 64-bit strided
// addressing, run-time division modes and predicated corner loads; this one is specialised down to ~110:
//   * the division mode is a template parameter; all plane offsets are 32-bit and the three planes share them;
//   * no predicated loads and two addresses per plane: the 2 x 2 patch is clamped into the frame as a whole, the 1-D weights
//     are re-slotted / zeroed for patches hanging over an edge (zeros padding), so all twelve gathers of a pixel are
//     unconditional, in flight together, and ten of them use immediate offsets.
// Arithmetic per corner and the coordinate replay are those of the generic kernel (warp_math.h), bit for bit.
#ifndef VFI_WARPF_BLOCK
#define VFI_WARPF_BLOCK 128
#endif
#ifndef VFI_WARPF_PPT
#define VFI_WARPF_PPT 2
#endif
constexpr int WARPF_BLOCK = VFI_WARPF_BLOCK;

template <bool RECIP> __device__ __forceinline__ float warp_coord_t(int pix, float disp, const WarpAxis& ax) {
  float g = VFI_MUL(2.0f, VFI_ADD((float)pix, disp));
  float q = VFI_MUL(g, ax.inv_denom);
  if (!RECIP) q = VFI_FMA(VFI_FMA(-q, ax.denom, g), ax.inv_denom, q);      // Markstein: q = RN(g / denom)
  const float i = VFI_MUL(VFI_MUL(VFI_ADD(VFI_SUB(q, 1.0f), 1.0f), 0.5f), ax.size_m1);
  return fminf(fmaxf(i, -4.0f), ax.hi);
}

// Three planar channels sampled at source position (ix, iy): the sampler of the fast kernels.
template <typename TS>
__device__ __forceinline__ void sample3_at(const TS* s0, const TS* s1, const TS* s2, int pitch, int H, int W, float ix, float iy,
                                           float (&r)[3]) {
  const int x0 = __float2int_rd(ix), y0 = __float2int_rd(iy);
  const float x0f = (float)x0, y0f = (float)y0;
  // The 2 x 2 patch is addressed from its clamped north-west pixel (xc, yc) in [0, W-2] x [0, H-2]: the other three
  // corners are at compile-time byte offsets / one row pitch.  d = x0 - xc is 0 inside the frame; -1 / +1 when the
  // true patch hangs over the left / right edge by one pixel (its inner column then sits in the other slot); anything
  // else means no valid corner.  Zero weights stand for aten's skipped corners, and the order of the non-zero
  // products (nw, ne, sw, se) is unchanged.
  const int xc = min(max(x0, 0), W - 2), yc = min(max(y0, 0), H - 2);
  const int dx = x0 - xc, dy = y0 - yc;
  const float ax1 = ix - x0f, ax0 = (x0f + 1.0f) - ix, ay1 = iy - y0f, ay0 = (y0f + 1.0f) - iy;
  const float wxa = dx == 0 ? ax0 : (dx == -1 ? ax1 : 0.0f), wxb = dx == 0 ? ax1 : (dx == 1 ? ax0 : 0.0f);
  const float wya = dy == 0 ? ay0 : (dy == -1 ? ay1 : 0.0f), wyb = dy == 0 ? ay1 : (dy == 1 ? ay0 : 0.0f);
  const float w00 = wxa * wya, w01 = wxb * wya, w10 = wxa * wyb, w11 = wxb * wyb;
  const unsigned o = (unsigned)(yc * pitch + xc);
  auto lerp = [&](const TS* pl) {
    const TS* q0 = pl + o;
    const TS* q1 = q0 + pitch;
    const float a = ldg_f32(q0), bb = ldg_f32(q0 + 1), d = ldg_f32(q1), e = ldg_f32(q1 + 1);
    return fmaf(e, w11, fmaf(d, w10, fmaf(bb, w01, a * w00)));
  };
  r[0] = lerp(s0); r[1] = lerp(s1); r[2] = lerp(s2);
}
// ... of one output pixel (x, y) displaced by (fx, fy)
template <typename TS, bool RECIP>
__device__ __forceinline__ void sample3(const TS* s0, const TS* s1, const TS* s2, int pitch, int H, int W, int x, int y, float fx,
                                        float fy, const WarpAxis& ax, const WarpAxis& ay, float (&r)[3]) {
  sample3_at<TS>(s0, s1, s2, pitch, H, W, warp_coord_t<RECIP>(x, fx, ax), warp_coord_t<RECIP>(y, fy, ay), r);
}

constexpr int WARPF_PPT = VFI_WARPF_PPT;                      // pixels per thread, one block apart (x, x + 128)

// Lanes are CONSECUTIVE pixels: the 2-byte gathers of a warp then span ~64 bytes + the flow's variation, i.e. one or two
// 128-byte lines per request.  (A thread owning two ADJACENT pixels halves the flow / store requests but doubles that span;
// measured: 3.7 L1 wavefronts per request and the LSU data pipe at 73 % -- the kernel's bound -- against ~1.3 here.)
template <typename TS, typename TF, bool REC, bool RECIP>
__global__ void __launch_bounds__(WARPF_BLOCK) warp_fwd_fast_kernel(const WarpParams p) {
  const int y = blockIdx.y, b = blockIdx.z;
  const int xb = blockIdx.x * (WARPF_BLOCK * WARPF_PPT) + threadIdx.x;
  const TF* fl = reinterpret_cast<const TF*>(p.flow) + b * p.f_sn + (long long)y * p.f_sh;
  const TF* fly = fl + p.f_sc;
  const TS* s0 = reinterpret_cast<const TS*>(p.src) + b * p.s_sn;
  const TS* s1 = s0 + p.s_sc;
  const TS* s2 = s1 + p.s_sc;
  const int pitch = (int)p.s_sh, H = p.H, W = p.W;
  float fx[WARPF_PPT], fy[WARPF_PPT];
#pragma unroll
  for (int i = 0; i < WARPF_PPT; ++i) {
    const int x = min(xb + i * WARPF_BLOCK, W - 1);             // out-of-range threads recompute the last pixel (not stored)
    fx[i] = to_f32<TF>(__ldcs(fl + x));
    fy[i] = to_f32<TF>(__ldcs(fly + x));
  }
  float r[WARPF_PPT][3];
#pragma unroll
  for (int i = 0; i < WARPF_PPT; ++i)
    sample3<TS, RECIP>(s0, s1, s2, pitch, H, W, min(xb + i * WARPF_BLOCK, W - 1), y, fx[i], fy[i], p.ax, p.ay, r[i]);
#pragma unroll
  for (int i = 0; i < WARPF_PPT; ++i) {
    const int x = xb + i * WARPF_BLOCK;
    if (x >= W) continue;
    if constexpr (REC) {
      // tail-plane record: 16 bytes per pixel = [c0 c1 c2 0 | c0 c1 c2 0] (mirrored halves, see dcn_tc.cu)
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + b * p.o_sn + (long long)y * p.o_sh + (long long)x * p.o_sw;
      const __nv_bfloat162 c01 = __floats2bfloat162_rn(r[i][0], r[i][1]), c2z = __floats2bfloat162_rn(r[i][2], 0.0f);
      const uint32_t lo = *reinterpret_cast<const uint32_t*>(&c01), hi = *reinterpret_cast<const uint32_t*>(&c2z);
      __stcs(reinterpret_cast<uint4*>(out), make_uint4(lo, hi, lo, hi));
    } else {
      TS* out = reinterpret_cast<TS*>(p.out) + b * p.o_sn + (long long)y * p.o_sh + x;
#pragma unroll
      for (int c = 0; c < 3; ++c) __stcs(out + c * p.o_sc, from_f32<TS>(r[i][c]));
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward, staged path
// The form the north star names: the source window of a tile is brought into shared memory by the copy engine (TMA), and the
// twelve corner gathers of a pixel become shared-memory loads.
//
// Persistent CTAs (as many as fit the SMs) walk over 32 x 32 tiles of output pixels (8 warps; a half-warp is one tile row, a
// thread owns two pairs of adjacent pixels) in a two-stage software pipeline -- per iteration, for tiles t (being produced), t + G (next)
// and t + 2 G (G = grid size):
//   A. flow of tile t + G (loaded during the previous iteration) -> source coordinates, replaying the reference's arithmetic bit
//      for bit as the fast kernel does; a warp reduction (redux.sync) + one shared-memory exchange gives the bounding box of the
//      tile's corners.  If it fits a staging window -- 48 x 40 source pixels for near-identity flow (what the reference's model
//      produces), else 64 x WsBox::H -- thread 0 issues ONE cp.async.bulk.tensor (three planes) into the other window buffer,
//      anchored at the box's north-west pixel (x rounded down to a 16-byte boundary: the copy engine faults on a misaligned
//      innermost coordinate).  The copy engine zero-fills whatever lies outside the frame, so the zeros padding of
//      F.grid_sample needs no clamping, no re-slotting of weights and no predicated loads.
//   B. the flow loads of tile t + 2 G are issued (they have a whole iteration to arrive).
//   C. tile t: wait for its window (issued one iteration ago), gather from shared memory (2-byte LDS by 32 consecutive pixels:
//      one 64-byte wavefront per request for small flows), same products in the same order (nw, ne, sw, se) as every other warp
//      kernel here, store.
// One __syncthreads per tile.  Tiles whose corners do not fit a window (large incoherent flow) take the fast kernel's L1 path
// inside the same launch, so the result never depends on the route (up to the sign of a zero: a zero-weight corner with a
// negative value gives -0 on the L1 path, the zero-filled window always +0, which is also what aten returns).
// What this buys over the L1 kernel: ~100 instead of ~190 issue slots per pixel (that kernel's bound),
// no data-dependent L1 wavefronts, and every DRAM / L2 latency of a tile hidden behind the previous tile's arithmetic.
#ifndef VFI_WS_MIN_CTAS
#define VFI_WS_MIN_CTAS 3
#endif
constexpr int WS_TILE = 32, WS_THREADS = 256, WS_MIN_CTAS = VFI_WS_MIN_CTAS;  // resident CTAs per SM the register budget allows
constexpr int WS_BW = 64, WS_SMALL_W = 48, WS_SMALL_H = 40;                // staging windows (source pixels): big is 64 x WsBox::H
// fp32 frames get a 64 x 32 big window (the bytes of the 16-bit one): three resident CTAs then leave the L1 the fall-back tiles
// gather through as large as in the 16-bit case (with 64 x 48 it shrank to ~30 KB and incoherent flow ran 3.5x slower).
template <typename TS> struct WsBox { static constexpr int H = sizeof(TS) == 4 ? 32 : 64; };

template <typename TS>
struct __align__(128) WsSmem {
  TS box[2][3 * WsBox<TS>::H * WS_BW];                                     // TMA destinations (128-byte aligned)
  unsigned long long bar[2];
  int red[2][8][4];
  int plan[2][4];                                                          // anchor x, anchor y, window width (0 = L1 path), window height
};

struct WarpStagedArgs {
  WarpParams p;
  unsigned long long* tile_counts;                                         // optional [2]: tiles staged / tiles on the L1 path
  int tiles_x, tiles_y, num_tiles;
  int dbg;                                                                 // VFI_WARP_DEBUG (diagnostics): 1 = never stage, 2 = big window only
  alignas(128) CUtensorMap tm_big;                                         // src {W, H, 3, B}, box {64, WsBox::H, 3, 1}
  alignas(128) CUtensorMap tm_small;                                       // same tensor, box {48, 40, 3, 1}
};

__device__ __forceinline__ uint32_t ws_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_tma_load(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(0), "r"(c3), "r"(bar)
               : "memory");
}

// Two adjacent elements as one load / store (flow pairs, output pairs).
template <typename T> struct Pair;
template <> struct Pair<float> { using type = float2; };
template <> struct Pair<__nv_bfloat16> { using type = uint32_t; };
template <> struct Pair<__half> { using type = uint32_t; };
__device__ __forceinline__ void unpack_pair(float2 v, const float*, float& a, float& b) { a = v.x; b = v.y; }
__device__ __forceinline__ void unpack_pair(uint32_t v, const __nv_bfloat16*, float& a, float& b) {
  a = __uint_as_float(v << 16); b = __uint_as_float(v & 0xffff0000u);
}
__device__ __forceinline__ void unpack_pair(uint32_t v, const __half*, float& a, float& b) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
  a = f.x; b = f.y;
}
__device__ __forceinline__ float2 pack_pair(float a, float b, const float*) { return make_float2(a, b); }
__device__ __forceinline__ uint32_t pack_pair(float a, float b, const __nv_bfloat16*) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_pair(float a, float b, const __half*) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Where a thread's pixels of one tile are: batch entry, row of its first pixel pair, column of the pair (always even).
struct WsOrigin { int b, y, x, tx, ty; };

template <typename TS, typename TF, bool REC, bool RECIP>
__global__ void __launch_bounds__(WS_THREADS, WS_MIN_CTAS) warp_fwd_staged_kernel(const __grid_constant__ WarpStagedArgs a) {
  constexpr int BH = WsBox<TS>::H;
  using FP = typename Pair<TF>::type;
  using SP = typename Pair<TS>::type;
  const WarpParams& p = a.p;
  extern __shared__ uint8_t ws_raw[];
  WsSmem<TS>& s = *reinterpret_cast<WsSmem<TS>*>(ws_raw + ((128 - (ws_smem_u32(ws_raw) & 127)) & 127));
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int H = p.H, W = p.W, G = (int)gridDim.x, N = a.num_tiles;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ws_smem_u32(&s.bar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ws_smem_u32(&s.bar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Thread -> pixels of a tile: a half-warp is one row of 32 pixels (lane & 15 = pixel pair), a warp two neighbouring rows, and
  // a thread owns the pairs at rows 4 w + h and 4 w + 2 + h (w = warp, h = lane / 16): flow loads and planar stores move two
  // pixels each.  Tiles advance by G = gridDim.x with carries instead of divisions.
  const int per_img = a.tiles_x * a.tiles_y;
  const int gb = G / per_img, gy = (G - gb * per_img) / a.tiles_x, gx = G - gb * per_img - gy * a.tiles_x;
  const int yin = 4 * wrp + (lane >> 4), xin = 2 * (lane & 15);
  auto first = [&](int t) {
    WsOrigin o;
    o.b = t / per_img;
    const int r = t - o.b * per_img;
    o.ty = r / a.tiles_x; o.tx = r - o.ty * a.tiles_x;
    o.y = o.ty * WS_TILE + yin; o.x = o.tx * WS_TILE + xin;
    return o;
  };
  auto advance = [&](WsOrigin o) {
    o.tx += gx;
    if (o.tx >= a.tiles_x) { o.tx -= a.tiles_x; ++o.ty; }
    o.ty += gy;
    if (o.ty >= a.tiles_y) { o.ty -= a.tiles_y; ++o.b; }
    o.b += gb;
    o.y = o.ty * WS_TILE + yin; o.x = o.tx * WS_TILE + xin;
    return o;
  };
  // (B) flow of a tile: the loads only -- the values stay raw so that nothing waits for them before the next iteration
  auto load_flow = [&](const WsOrigin& o, FP (&fx)[2], FP (&fy)[2]) {
    const TF* fl = reinterpret_cast<const TF*>(p.flow) + o.b * p.f_sn + min(o.x, W - 2);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const TF* q = fl + (long long)min(o.y + 2 * j, H - 1) * p.f_sh;
      fx[j] = __ldcs(reinterpret_cast<const FP*>(q));
      fy[j] = __ldcs(reinterpret_cast<const FP*>(q + p.f_sc));
    }
  };
  // (A) coordinates of a tile (pixel 2 j + e: row pair j, column x + e), its bounding box, the plan and the copy into window
  // `slot`.  Contains the tile's __syncthreads.
  auto plan_tile = [&](const WsOrigin& o, int slot, const FP (&fx)[2], const FP (&fy)[2], float (&ix)[4], float (&iy)[4]) {
    int mnx = 0x7fffffff, mxx = -0x7fffffff, mny = 0x7fffffff, mxy = -0x7fffffff;
    const int xc = min(o.x, W - 2);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int y = o.y + 2 * j;
      float f0, f1, g0, g1;
      unpack_pair(fx[j], static_cast<const TF*>(nullptr), f0, f1);
      unpack_pair(fy[j], static_cast<const TF*>(nullptr), g0, g1);
      ix[2 * j] = warp_coord_t<RECIP>(xc, f0, p.ax);
      ix[2 * j + 1] = warp_coord_t<RECIP>(xc + 1, f1, p.ax);
      iy[2 * j] = warp_coord_t<RECIP>(min(y, H - 1), g0, p.ay);
      iy[2 * j + 1] = warp_coord_t<RECIP>(min(y, H - 1), g1, p.ay);
      if (o.x < W && y < H) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int xi = __float2int_rd(ix[2 * j + e]), yi = __float2int_rd(iy[2 * j + e]);
          mnx = min(mnx, xi); mxx = max(mxx, xi); mny = min(mny, yi); mxy = max(mxy, yi);
        }
      }
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
    mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
    if (lane == 0) *reinterpret_cast<int4*>(&s.red[slot][wrp][0]) = make_int4(mnx, mxx, mny, mxy);
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const int4 v = *reinterpret_cast<const int4*>(&s.red[slot][w][0]);
        mnx = min(mnx, v.x); mxx = max(mxx, v.y); mny = min(mny, v.z); mxy = max(mxy, v.w);
      }
      mnx = (mnx >> 3) << 3;                                               // 16-byte aligned first byte of every box row
      const int ex = mxx - mnx + 2, ey = mxy - mny + 2;                    // source pixels from the anchor to the last corner
      int bw = 0, bh = 0;
      if (ex <= WS_SMALL_W && ey <= WS_SMALL_H && !(a.dbg & 2)) { bw = WS_SMALL_W; bh = WS_SMALL_H; }
      else if (ex <= WS_BW && ey <= BH) { bw = WS_BW; bh = BH; }
      if (a.dbg & 1) bw = bh = 0;
      *reinterpret_cast<int4*>(&s.plan[slot][0]) = make_int4(mnx, mny, bw, bh);
      if (bw) {
        const uint32_t bar_a = ws_smem_u32(&s.bar[slot]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)(3 * bw * bh * sizeof(TS))) : "memory");
        if (bw == WS_SMALL_W) ws_tma_load(ws_smem_u32(&s.box[slot][0]), &a.tm_small, mnx, mny, o.b, bar_a);
        else ws_tma_load(ws_smem_u32(&s.box[slot][0]), &a.tm_big, mnx, mny, o.b, bar_a);
      }
      if (a.tile_counts) atomicAdd(a.tile_counts + (bw ? 0 : 1), 1ull);
    }
  };

  int t = (int)blockIdx.x;
  if (t >= N) return;
  WsOrigin oc, on = first(t), of;                                          // tile being produced / planned / whose flow is in flight
  FP fx[2], fy[2];
  float nix[4], niy[4];
  uint32_t phase = 0;                                                      // bit s: parity of window s's next completed copy
  load_flow(on, fx, fy);
  plan_tile(on, 0, fx, fy, nix, niy);
  of = advance(on);
  if (t + G < N) load_flow(of, fx, fy);
  for (int it = 0; t < N; ++it, t += G) {
    const int slot = it & 1;
    float ix[4], iy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ix[i] = nix[i]; iy[i] = niy[i]; }
    oc = on; on = of;
    if (t + G < N) plan_tile(on, slot ^ 1, fx, fy, nix, niy);              // (A); its barrier also publishes plan[slot]
    else __syncthreads();
    of = advance(on);
    if (t + 2 * G < N) load_flow(of, fx, fy);                              // (B)
    // ---- (C) tile t
    const int4 pl = *reinterpret_cast<const int4*>(&s.plan[slot][0]);
    const int ax0 = pl.x, ay0 = pl.y, bw = pl.z, bh = pl.w;
    const bool ok = oc.x < W;                                              // W is even: both pixels of a pair or neither
    float r[4][3];
    if (bw) {
      const uint32_t bar_a = ws_smem_u32(&s.bar[slot]), par = (phase >> slot) & 1u;
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a), "r"(par) : "memory");
      phase ^= 1u << slot;
      const TS* box = &s.box[slot][0];
      const int plane = bw * bh;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int xi = __float2int_rd(ix[i]), yi = __float2int_rd(iy[i]);
        const float x0f = (float)xi, y0f = (float)yi;
        const float wx1 = ix[i] - x0f, wx0 = (x0f + 1.0f) - ix[i], wy1 = iy[i] - y0f, wy0 = (y0f + 1.0f) - iy[i];
        const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;
        const int o = (ok && oc.y + 2 * (i >> 1) < H) ? (yi - ay0) * bw + (xi - ax0) : 0;   // pixels past the frame edge are not stored
        const TS* q = box + o;
#pragma unroll
        for (int c = 0; c < 3; ++c, q += plane) {
          const float v00 = to_f32<TS>(q[0]), v01 = to_f32<TS>(q[1]), v10 = to_f32<TS>(q[bw]), v11 = to_f32<TS>(q[bw + 1]);
          r[i][c] = fmaf(v11, w11, fmaf(v10, w10, fmaf(v01, w01, v00 * w00)));
        }
      }
    } else {
      // the fast kernel's L1 path (the tile's corners do not fit a window)
      const TS* s0 = reinterpret_cast<const TS*>(p.src) + oc.b * p.s_sn;
#pragma unroll
      for (int i = 0; i < 4; ++i) sample3_at<TS>(s0, s0 + p.s_sc, s0 + 2 * p.s_sc, (int)p.s_sh, H, W, ix[i], iy[i], r[i]);
    }
    if (ok) {
      if constexpr (REC) {
        uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + oc.b * p.o_sn + (long long)oc.y * p.o_sh +
                                              (long long)oc.x * p.o_sw);
        const long long row2 = 2 * p.o_sh / 8, px = p.o_sw / 8;            // in 16-byte records
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (oc.y + 2 * (i >> 1) >= H) continue;
          const uint32_t lo = pack_pair(r[i][0], r[i][1], static_cast<const __nv_bfloat16*>(nullptr));
          const uint32_t hi = pack_pair(r[i][2], 0.0f, static_cast<const __nv_bfloat16*>(nullptr));
          __stcs(out + (i >> 1) * row2 + (i & 1) * px, make_uint4(lo, hi, lo, hi));
        }
      } else {
        TS* out = reinterpret_cast<TS*>(p.out) + oc.b * p.o_sn + (long long)oc.y * p.o_sh + oc.x;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (oc.y + 2 * j >= H) continue;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            __stcs(reinterpret_cast<SP*>(out + c * p.o_sc + 2 * j * p.o_sh), pack_pair(r[2 * j][c], r[2 * j + 1][c], static_cast<const TS*>(nullptr)));
        }
      }
    }
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
  Corners ca = locate(x, y, to_f32<TF>(fa[0]), to_f32<TF>(fa[p.f_sc]), p.ax, p.ay, p.H, p.W, (int)p.s_sh, (int)p.s_sw);
  Corners cb = locate(x, y, to_f32<TF>(fb[0]), to_f32<TF>(fb[q.g_sc]), p.ax, p.ay, p.H, p.W, (int)q.b_sh, (int)q.b_sw);
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

// Fast path of the blend (three planar channels, unit pixel strides): the two warps share the fast kernel's sampler, one
// thread per pixel, lanes consecutive.  out = m * warp(a, fa) + (1 - m) * warp(b, fb), rounded as the composition is.
template <typename TS, typename TF, bool RECIP>
__global__ void __launch_bounds__(WARPF_BLOCK) warp_blend_fast_kernel(const BlendParams q) {
  const WarpParams& p = q.a;
  const int y = blockIdx.y, b = blockIdx.z;
  const int x = blockIdx.x * WARPF_BLOCK + threadIdx.x;
  if (x >= p.W) return;
  const TF* fa = reinterpret_cast<const TF*>(p.flow) + b * p.f_sn + (long long)y * p.f_sh + x;
  const TF* fb = reinterpret_cast<const TF*>(q.flow_b) + b * q.g_sn + (long long)y * q.g_sh + x;
  const float fax = to_f32<TF>(__ldcs(fa)), fay = to_f32<TF>(__ldcs(fa + p.f_sc));
  const float fbx = to_f32<TF>(__ldcs(fb)), fby = to_f32<TF>(__ldcs(fb + q.g_sc));
  const float m = to_f32<TS>(__ldcs(reinterpret_cast<const TS*>(q.m) + b * q.m_sn + (long long)y * q.m_sh + x));
  const TS* a0 = reinterpret_cast<const TS*>(p.src) + b * p.s_sn;
  const TS* b0 = reinterpret_cast<const TS*>(q.src_b) + b * q.b_sn;
  float ra[3], rb[3];
  sample3<TS, RECIP>(a0, a0 + p.s_sc, a0 + 2 * p.s_sc, (int)p.s_sh, p.H, p.W, x, y, fax, fay, p.ax, p.ay, ra);
  sample3<TS, RECIP>(b0, b0 + q.b_sc, b0 + 2 * q.b_sc, (int)q.b_sh, p.H, p.W, x, y, fbx, fby, p.ax, p.ay, rb);
  const float m1 = 1.0f - m;
  TS* out = reinterpret_cast<TS*>(p.out) + b * p.o_sn + (long long)y * p.o_sh + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) __stcs(out + c * p.o_sc, from_f32<TS>(__fadd_rn(__fmul_rn(m, ra[c]), __fmul_rn(m1, rb[c]))));
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
  Corners c = locate(x, y, to_f32<TF>(fl[0]), to_f32<TF>(fl[p.f_sc]), p.ax, p.ay, p.H, p.W, (int)p.s_sh, (int)p.s_sw);
  const TS* src = reinterpret_cast<const TS*>(p.src) + b * p.s_sn;
  const TG* go = reinterpret_cast<const TG*>(q.gout) + b * q.go_sn + y * q.go_sh + x * q.go_sw;
  float gix = 0.0f, giy = 0.0f;
  for (int ch = 0; ch < p.C; ++ch) {
    const TS* plane = src + ch * p.s_sc;
    float g = to_f32<TG>(__ldg(go + ch * q.go_sc));
    float nw = corner(plane, c.off00), ne = corner(plane, c.off01);
    float sw = corner(plane, c.off10), se = corner(plane, c.off11);
    gix += ((ne - nw) * c.wy0 + (se - sw) * c.wy1) * g;
    giy += ((sw - nw) * c.wx0 + (se - ne) * c.wx1) * g;
    if (q.gsrc) {
      // same pixel geometry as src but addressed with grad_src's own strides
      float* gs = q.gsrc + b * q.gs_sn + ch * q.gs_sc;
      const int x0 = c.x0, y0 = c.y0;
      if (c.off00 >= 0) atomicAdd(gs + y0 * q.gs_sh + x0 * q.gs_sw, g * c.w00);
      if (c.off01 >= 0) atomicAdd(gs + y0 * q.gs_sh + (x0 + 1) * q.gs_sw, g * c.w01);
      if (c.off10 >= 0) atomicAdd(gs + (y0 + 1) * q.gs_sh + x0 * q.gs_sw, g * c.w10);
      if (c.off11 >= 0) atomicAdd(gs + (y0 + 1) * q.gs_sh + (x0 + 1) * q.gs_sw, g * c.w11);
    }
  }
  // d(grid)/d(flow): aten scales by (size-1)/2, autograd of "2.0 * v / denom" divides by denom and doubles.
  float* gf = q.gflow + b * q.gf_sn + y * q.gf_sh + x * q.gf_sw;
  const float ggx = __fmul_rn(q.mult_x, gix), ggy = __fmul_rn(q.mult_y, giy);
  gf[0] = __fmul_rn(p.ax.recip ? __fmul_rn(ggx, p.ax.inv_denom) : __fdiv_rn(ggx, p.ax.denom), 2.0f);
  gf[q.gf_sc] = __fmul_rn(p.ay.recip ? __fmul_rn(ggy, p.ay.inv_denom) : __fdiv_rn(ggy, p.ay.denom), 2.0f);
}

// ------------------------------------------------------------------------------------------------ backward, fast path
// grad_flow of the model path (EMA_VFI.warp is differentiated with respect to the flow only: frame2 is an input image):
// C = 3 planar frames, unit pixel strides, no grad_src.  Same treatment as the forward fast path -- lanes are consecutive
// pixels, the 2 x 2 patch is clamped into the frame as a whole so all twelve gathers are unconditional and in flight
// together; corners outside the frame are then ZEROED by re-slotting (the generic kernel skips their loads), so the
// expression below is the generic kernel's, operand for operand.
template <typename TS, typename TF, typename TG, bool RECIP>
__global__ void __launch_bounds__(WARPF_BLOCK) warp_bwd_fast_kernel(const WarpBwdParams q) {
  const WarpParams& p = q.f;
  const int y = blockIdx.y, b = blockIdx.z;
  const int x = blockIdx.x * WARPF_BLOCK + threadIdx.x;
  if (x >= p.W) return;
  const int H = p.H, W = p.W, pitch = (int)p.s_sh;
  const TF* fl = reinterpret_cast<const TF*>(p.flow) + b * p.f_sn + (long long)y * p.f_sh + x;
  const float ix = warp_coord_t<RECIP>(x, to_f32<TF>(__ldcs(fl)), p.ax);
  const float iy = warp_coord_t<RECIP>(y, to_f32<TF>(__ldcs(fl + p.f_sc)), p.ay);
  const int x0 = __float2int_rd(ix), y0 = __float2int_rd(iy);
  const float x0f = (float)x0, y0f = (float)y0;
  const float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix, wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
  const int xc = min(max(x0, 0), W - 2), yc = min(max(y0, 0), H - 2);
  const int dx = x0 - xc, dy = y0 - yc;       // 0 inside; -1 / +1: the patch hangs over the low / high edge by one pixel
  const unsigned o = (unsigned)(yc * pitch + xc);
  const TS* s0 = reinterpret_cast<const TS*>(p.src) + b * p.s_sn;
  const TG* go = reinterpret_cast<const TG*>(q.gout) + b * q.go_sn + (long long)y * q.go_sh + x;
  float v[3][4], g[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const TS* q0 = s0 + c * p.s_sc + o;
    v[c][0] = ldg_f32(q0); v[c][1] = ldg_f32(q0 + 1); v[c][2] = ldg_f32(q0 + pitch); v[c][3] = ldg_f32(q0 + pitch + 1);
    g[c] = to_f32<TG>(__ldcs(go + c * q.go_sc));
  }
  float gix = 0.0f, giy = 0.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // rows first (top / bottom of the true patch), then columns
    const float tl = dy == 0 ? v[c][0] : (dy == 1 ? v[c][2] : 0.0f), tr = dy == 0 ? v[c][1] : (dy == 1 ? v[c][3] : 0.0f);
    const float bl = dy == 0 ? v[c][2] : (dy == -1 ? v[c][0] : 0.0f), br = dy == 0 ? v[c][3] : (dy == -1 ? v[c][1] : 0.0f);
    const float nw = dx == 0 ? tl : (dx == 1 ? tr : 0.0f), ne = dx == 0 ? tr : (dx == -1 ? tl : 0.0f);
    const float sw = dx == 0 ? bl : (dx == 1 ? br : 0.0f), se = dx == 0 ? br : (dx == -1 ? bl : 0.0f);
    gix += ((ne - nw) * wy0 + (se - sw) * wy1) * g[c];
    giy += ((sw - nw) * wx0 + (se - ne) * wx1) * g[c];
  }
  float* gf = q.gflow + b * q.gf_sn + (long long)y * q.gf_sh + x;
  const float ggx = __fmul_rn(q.mult_x, gix), ggy = __fmul_rn(q.mult_y, giy);
  __stcs(gf, __fmul_rn(RECIP ? __fmul_rn(ggx, p.ax.inv_denom) : __fdiv_rn(ggx, p.ax.denom), 2.0f));
  __stcs(gf + q.gf_sc, __fmul_rn(RECIP ? __fmul_rn(ggy, p.ay.inv_denom) : __fdiv_rn(ggy, p.ay.denom), 2.0f));
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

// Tensor map over the three source planes of every frame: dims {W, H, 3, B} (innermost first), box {bw, bh, 3, 1}, zero fill
// outside the tensor.  False when the copy engine cannot describe the tensor (alignment / stride rules of cuTensorMapEncodeTiled).
template <typename TS>
bool warp_src_map(CUtensorMap* m, const WarpParams& p, int bw, int bh) {
  EncodeTiledFn enc = tensor_map_encoder();
  const long long es = (long long)sizeof(TS);
  if (!enc || !aligned(p.src, 16) || p.s_sw != 1) return false;
  const long long sb[3] = {p.s_sh * es, p.s_sc * es, (p.B > 1 ? p.s_sn : p.s_sc * 3) * es};
  for (int i = 0; i < 3; ++i)
    if (sb[i] <= 0 || sb[i] % 16 != 0 || sb[i] >= (1LL << 40)) return false;
  const cuuint64_t gd[4] = {(cuuint64_t)p.W, (cuuint64_t)p.H, 3, (cuuint64_t)p.B};
  const cuuint64_t gs[3] = {(cuuint64_t)sb[0], (cuuint64_t)sb[1], (cuuint64_t)sb[2]};
  const cuuint32_t bx[4] = {(cuuint32_t)bw, (cuuint32_t)bh, 3, 1}, el[4] = {1, 1, 1, 1};
  return enc(m, sizeof(TS) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(p.src), gd, gs, bx,
             el, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ unsigned long long g_warp_tile_counts[2];                       // VFI_WARP_COUNT_TILES: tiles staged / on the L1 path

template <typename TS, typename TF>
int launch_fwd(const WarpParams& p, bool rec, bool fast, int flags, cudaStream_t st) {
  // the staged kernel moves pixel PAIRS: even width, flow rows / planes and (planar) output rows / planes aligned to a pair
  const auto pair_ok = [](const void* base, size_t es, long long a, long long b, long long c) {
    return aligned(base, 2 * es) && a % 2 == 0 && b % 2 == 0 && c % 2 == 0;
  };
  const bool pairs = p.W % 2 == 0 && pair_ok(p.flow, sizeof(TF), p.f_sh, p.f_sc, p.f_sn) &&
                     (rec || pair_ok(p.out, sizeof(TS), p.o_sh, p.o_sc, p.o_sn));
  if (fast && pairs && !(flags & VFI_WARP_NO_STAGING) && p.W >= WS_TILE && p.H >= 8) {
    // staged path (TMA window + shared-memory gathers); falls through to the L1 kernels when the source cannot be mapped
    WarpStagedArgs a;
    a.p = p;
    a.tile_counts = nullptr;
    const char* dbg_env = getenv("VFI_WARP_DEBUG");
    a.dbg = dbg_env ? atoi(dbg_env) : 0;
    if (warp_src_map<TS>(&a.tm_big, p, WS_BW, WsBox<TS>::H) && warp_src_map<TS>(&a.tm_small, p, WS_SMALL_W, WS_SMALL_H)) {
      if (flags & VFI_WARP_COUNT_TILES) VFI_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&a.tile_counts), g_warp_tile_counts));
      a.tiles_x = ceil_div(p.W, WS_TILE); a.tiles_y = ceil_div(p.H, WS_TILE);
      const long long tiles = (long long)a.tiles_x * a.tiles_y * p.B;
      VFI_REQUIRE(tiles < (1LL << 30), VFI_ERR_UNSUPPORTED, "vfi_warp_fwd: too many tiles");
      a.num_tiles = (int)tiles;
      const bool recip = p.ax.recip != 0;
      const size_t smem = sizeof(WsSmem<TS>) + 128;
      // persistent grid: every CTA the SMs can hold at once
      auto launch = [&](void (*kern)(const WarpStagedArgs)) -> int {
        // resident CTAs per SM of this instantiation on this device: asked once (the answer also records that the kernel's
        // dynamic shared-memory limit has been raised)
        static std::mutex mu;
        static std::map<std::pair<const void*, int>, int> resident;
        static int sms_of[64] = {0};
        int dev = 0;
        VFI_CUDA(cudaGetDevice(&dev));
        int ctas = 0, sms = 0;
        {
          std::lock_guard<std::mutex> lock(mu);
          auto it = resident.find({reinterpret_cast<const void*>(kern), dev});
          if (it != resident.end()) { ctas = it->second; sms = sms_of[dev & 63]; }
        }
        if (!ctas) {
          VFI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
          VFI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          VFI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, WS_THREADS, smem));
          if (ctas < 1) ctas = 1;
          std::lock_guard<std::mutex> lock(mu);
          resident[{reinterpret_cast<const void*>(kern), dev}] = ctas;
          sms_of[dev & 63] = sms;
        }
        const int grid = (int)(tiles < (long long)sms * ctas ? tiles : (long long)sms * ctas);
        kern<<<grid, WS_THREADS, smem, st>>>(a);
        return VFI_OK;
      };
      bool launched = false;
      int lrc = VFI_OK;
      if constexpr (std::is_same<TS, __nv_bfloat16>::value) {
        if (rec) {
          lrc = recip ? launch(warp_fwd_staged_kernel<TS, TF, true, true>) : launch(warp_fwd_staged_kernel<TS, TF, true, false>);
          launched = true;
        }
      }
      if (!launched) lrc = recip ? launch(warp_fwd_staged_kernel<TS, TF, false, true>) : launch(warp_fwd_staged_kernel<TS, TF, false, false>);
      if (lrc) return lrc;
      VFI_LAUNCH_CHECK("warp_fwd_staged_kernel");
      return VFI_OK;
    }
  }
  if (fast) {
    dim3 grid(ceil_div(p.W, WARPF_BLOCK * WARPF_PPT), p.H, p.B);
    const bool recip = p.ax.recip != 0;
    if constexpr (std::is_same<TS, __nv_bfloat16>::value) {
      if (rec) {
        if (recip) warp_fwd_fast_kernel<TS, TF, true, true><<<grid, WARPF_BLOCK, 0, st>>>(p);
        else warp_fwd_fast_kernel<TS, TF, true, false><<<grid, WARPF_BLOCK, 0, st>>>(p);
        VFI_LAUNCH_CHECK("warp_fwd_fast_kernel<rec>");
        return VFI_OK;
      }
    }
    if (recip) warp_fwd_fast_kernel<TS, TF, false, true><<<grid, WARPF_BLOCK, 0, st>>>(p);
    else warp_fwd_fast_kernel<TS, TF, false, false><<<grid, WARPF_BLOCK, 0, st>>>(p);
    VFI_LAUNCH_CHECK("warp_fwd_fast_kernel");
    return VFI_OK;
  }
  dim3 grid(ceil_div(p.W, WARP_BLOCK * WARP_PPT), p.H, p.B);
  if constexpr (std::is_same<TS, __nv_bfloat16>::value) {
    if (rec) {
      warp_fwd_kernel<TS, TF, 3, true><<<grid, WARP_BLOCK, 0, st>>>(p);
      VFI_LAUNCH_CHECK("warp_fwd_kernel<rec>");
      return VFI_OK;
    }
  }
  if (p.C == 3) warp_fwd_kernel<TS, TF, 3, false><<<grid, WARP_BLOCK, 0, st>>>(p);
  else warp_fwd_kernel<TS, TF, 0, false><<<grid, WARP_BLOCK, 0, st>>>(p);
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
  VFI_REQUIRE(src->h <= 65535 && src->n <= 65535, VFI_ERR_UNSUPPORTED, "vfi_warp_fwd: H and B must be <= 65535");
  WarpParams p = make_params(src, flow, out, flags);
  cudaStream_t st = (cudaStream_t)stream;
  // tail-plane record output ([B,H,W,8] bf16 records holding C = 3 channels + zeros): only on request, because whole
  // 16-byte records are written -- a caller's [:, :3] slice of an ordinary channels-last tensor must keep channels 3..7
  const bool rec_shape = src->dtype == VFI_BF16 && src->c == 3 && out->sc == 1 && out->sw >= 8 && out->sw % 8 == 0 && out->sh % 8 == 0 &&
                         out->sn % 8 == 0 && aligned(out->data, 16);
  const bool rec = (flags & VFI_WARP_OUT_TAIL_RECORD) != 0;
  VFI_REQUIRE(!rec || rec_shape, VFI_ERR_INVALID,
              "vfi_warp_fwd: VFI_WARP_OUT_TAIL_RECORD needs a bf16 [B,3,H,W] view of 16-byte channels-last records");
  // fast path: three planar channels with unit pixel strides
  const bool fast = src->c == 3 && src->sw == 1 && flow->sw == 1 && src->w >= 2 && src->h >= 2 && src->sh >= 0 && src->sc >= 0 &&
                    src->sn >= 0 && (rec || out->sw == 1);
  VFI_DISPATCH(src->dtype, TS, {
    if (flow->dtype == VFI_F32) { rc = launch_fwd<TS, float>(p, rec, fast, flags, st); }
    else { rc = launch_fwd<TS, TS>(p, rec, fast, flags, st); }
  });
  return rc;
}

extern "C" int vfi_warp_tile_counts(uint64_t* staged, uint64_t* direct, int32_t reset) {
  unsigned long long h[2] = {0, 0};
  VFI_CUDA(cudaMemcpyFromSymbol(h, g_warp_tile_counts, sizeof(h)));
  if (staged) *staged = h[0];
  if (direct) *direct = h[1];
  if (reset) {
    const unsigned long long z[2] = {0, 0};
    VFI_CUDA(cudaMemcpyToSymbol(g_warp_tile_counts, z, sizeof(z)));
  }
  return VFI_OK;
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
  auto planar3 = [](const vfi_tensor* t) { return t->c == 3 && t->sw == 1 && t->sh >= 0 && t->sc >= 0 && t->sn >= 0; };
  const bool fast = planar3(src_a) && planar3(src_b) && planar3(out) && flow_a->sw == 1 && flow_b->sw == 1 && m->sw == 1 &&
                    out->w >= 2 && out->h >= 2 && out->h <= 65535 && out->n <= 65535;
  if (fast) {
    dim3 grid(ceil_div(out->w, WARPF_BLOCK), (unsigned)out->h, (unsigned)out->n);
    const bool recip = (flags & VFI_WARP_DIV_RECIPROCAL) != 0;
    VFI_DISPATCH(out->dtype, TS, {
      if (flow_a->dtype == VFI_F32) {
        if (recip) warp_blend_fast_kernel<TS, float, true><<<grid, WARPF_BLOCK, 0, st>>>(q);
        else warp_blend_fast_kernel<TS, float, false><<<grid, WARPF_BLOCK, 0, st>>>(q);
      } else {
        if (recip) warp_blend_fast_kernel<TS, TS, true><<<grid, WARPF_BLOCK, 0, st>>>(q);
        else warp_blend_fast_kernel<TS, TS, false><<<grid, WARPF_BLOCK, 0, st>>>(q);
      }
    });
    VFI_LAUNCH_CHECK("warp_blend_fast_kernel");
    return VFI_OK;
  }
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
  const bool fast = !grad_src && src->c == 3 && src->sw == 1 && flow->sw == 1 && grad_out->sw == 1 && grad_flow->sw == 1 &&
                    src->h >= 2 && src->w >= 2 && src->sh >= 0 && src->n <= 65535 && src->h <= 65535;
  if (fast) {
    dim3 grid(ceil_div(src->w, WARPF_BLOCK), (unsigned)src->h, (unsigned)src->n);
    const bool recip = q.f.ax.recip != 0;
    VFI_DISPATCH(src->dtype, TS, {
      VFI_DISPATCH(grad_out->dtype, TG, {
        if (flow->dtype == VFI_F32) {
          if (recip) warp_bwd_fast_kernel<TS, float, TG, true><<<grid, WARPF_BLOCK, 0, st>>>(q);
          else warp_bwd_fast_kernel<TS, float, TG, false><<<grid, WARPF_BLOCK, 0, st>>>(q);
        } else {
          if (recip) warp_bwd_fast_kernel<TS, TS, TG, true><<<grid, WARPF_BLOCK, 0, st>>>(q);
          else warp_bwd_fast_kernel<TS, TS, TG, false><<<grid, WARPF_BLOCK, 0, st>>>(q);
        }
      });
    });
    VFI_LAUNCH_CHECK("warp_bwd_fast_kernel");
    return VFI_OK;
  }
  VFI_DISPATCH(src->dtype, TS, {
    VFI_DISPATCH(grad_out->dtype, TG, {
      if (flow->dtype == VFI_F32) warp_bwd_kernel<TS, float, TG><<<blocks, 256, 0, st>>>(q);
      else warp_bwd_kernel<TS, TS, TG><<<blocks, 256, 0, st>>>(q);
    });
  });
  VFI_LAUNCH_CHECK("warp_bwd_kernel");
  return VFI_OK;
}
