// dcn_tc.cu -- DCNv2 on the tensor cores of sm_100a (bf16 operands, fp32 accumulate): the translation unit of the throughput path.
//
// What lives here:   host side of every tensor-core entry point (dispatch, argument checks, workspace layout, TMA tensor maps),
//                    the weight and input packers, the PTX wrappers the kernels share, and the tcgen05 self test;
// dcn_tc7.cuh        the forward kernel (v7): A operand written to tensor memory by the gather warps, source box / offsets /
//                    output through TMA tensor maps, tail channels gathered by the geometry warps -- its header has the design;
// dcn_tc6.cuh        v6, the round-1 kernel: kept as the bit-exactness reference of v7 (VFI_DCN_KERNEL=v6, one instantiation)
//                    and for the box / geometry roles the weight-gradient kernel shares;
// dcn_tc6_wgrad.cuh  weight / bias gradient (pixel-reduction GEMM, accumulators persistent in tensor memory);
// dcn_gcol.cuh       column gradient gcol = grad_out x W (A in tensor memory, weights resident in shared memory, TMA stores);
// dcn_bwd_cols.cuh   input / offset / mask gradients from the column gradient (no tensor-core work; shares the packers).
//
// Replaces torchvision::deform_conv2d (call site /root/reference/src/models/ema_vfi.py:60) on the throughput path.
// torchvision materialises columns[603, P] in HBM (40 GB fp32 at 1080p batch 8) and calls a BLAS GEMM; here the modulated
// bilinear im2col is the A-operand PRODUCER of the GEMM and never leaves the SM:
//
//     D[128 px, 80 o] (TMEM, fp32)  +=  A[128 px, 64 k] (TMEM, written by the gather warps)  x  B[80 o, 64 k]^T (smem, bulk copy)
//
// Activation layout ("planes"): channels-last bf16, main = 64 channels (128 B per pixel = exactly one aligned line per
// bilinear corner) and tail = a 16-byte record per pixel (channels 64.. and zeros; with at most four tail channels the upper
// 8 bytes MIRROR the lower 8, so that a gather may read either half and spread its 8-byte reads over all 32 banks).  The two
// may be separate dense buffers ([B,H,W,64] + [B,H,W,8]) or the channel ranges 0..63 / 64..71 of ONE [B,H,W,72] buffer (v7
// addresses them through tensor maps).  This is how feat (64 ch) and the warped frame (3 ch) exist before the reference's
// torch.cat (ema_vfi.py:134), and it is what the kernel writes for the next layer.  Any other input layout is converted by
// pack_input_kernel (workspace).
//
// (The round-1 v4 kernel -- L1 gathers into a shared-memory A ring, fp32 "HQ" blend, C up to 72 -- was removed in round 2: the
// reference geometry is C = 67 (ema_vfi.py:96-99) and v7 takes every 16-bit offset layout, so nothing reached it any more.)
#include <cuda.h>   // CUtensorMap (types only: the encoder is resolved at run time through cudaGetDriverEntryPoint)

#include <cstddef>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace vfi {

constexpr int TC_M = 128;
constexpr int TC_N = 80;
constexpr int TC_CMAIN = 64;                         // channels in the main plane (8 chunks of 16 B)
constexpr int TC_CTAIL = 8;                          // channels in the tail plane (1 chunk)
constexpr int TC_CMAX = TC_CMAIN + TC_CTAIL;         // 72
constexpr int TC_TH = 8, TC_TW = 16;                 // output tile (rows x cols) = 128 pixels
constexpr int TC_KBLOCKS = 11;                       // 128-byte swizzle atoms along K (9 main + 2 tail)
constexpr int TC_A_BYTES = TC_M * 128;               // 16384
constexpr int TC_B_BYTES = TC_N * 128;               // 10240
constexpr int TC_ACC_STRIDE = 128;

bool dcn_tc_available() { return true; }
size_t dcn_tc_packed_weight_bytes() { return (size_t)TC_KBLOCKS * TC_B_BYTES; }   // 112,640

// workspace: [packed weight image | bias f32[80] (512 B) | main plane P*64 bf16 | tail plane P*8 bf16]
static size_t ws_bias_off() { return dcn_tc_packed_weight_bytes(); }
static size_t ws_main_off() { return dcn_tc_packed_weight_bytes() + 512; }
static size_t ws_tail_off(long long P) { return ws_main_off() + (((size_t)P * TC_CMAIN * 2 + 255) / 256) * 256; }
size_t dcn_tc_workspace_bytes(long long B, long long H, long long W) {
  const long long P = B * H * W;
  return ws_tail_off(P) + (((size_t)P * TC_CTAIL * 2 + 255) / 256) * 256;
}

namespace {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
// Bounded wait.  try_wait suspends the thread in hardware for up to the hinted time, so a waiting role does not burn
// issue slots (the un-hinted form returns every ~150 cycles; measured: 29% of all executed instructions were spin loops).
// A protocol bug traps (launch failure) after ~2 s instead of hanging the GPU.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done)
               : "r"(bar), "r"(parity), "r"(hint_ns)
               : "memory");
  return done != 0;
}
// Slow path of a wait.  try_wait returns whenever ANY mbarrier of the CTA sees an arrival (measured: a waiting warp came
// back every ~170 cycles regardless of the hint), so a bare retry loop burns the issue slots the producers need; the
// explicit sleep bounds that cost.  NS trades issue slots against wake-up latency per wait site.  A protocol bug ends the
// kernel through the watchdog below instead of hanging the GPU.
// Diagnostics of a stuck pipeline: the first waiter that times out records who it is and raises the flag; every other
// waiter sees the flag and leaves too, so the kernel ends (with garbage results) and vfi_debug_abort_info can be read.
__device__ unsigned int g_abort_flag = 0;
__device__ unsigned long long g_abort_info[4 + 32] = {0};        // [block << 32 | warp, barrier smem address, parity, count,
                                                                 //  then per warp of that block: barrier address | parity << 32]

// Slow path of a wait.  Nothing but the sleep, the counter and the retry sit on the hot path: every waiting warp runs it.
// (A %globaltimer read at entry, or an out-of-line watchdog function -- any CALL in this body makes the compiler save
// registers around it -- cost the forward kernel 6-9 %.)  Every 1024 spins the waiter looks at the abort flag; after 2^21
// spins (0.2 .. 2.7 s: a spin is one nanosleep plus one try_wait of at most 1 us; legitimate waits are milliseconds) it
// raises the flag itself.  Either way it records what it was waiting on and leaves, so the kernel ends instead of hanging.
template <int NS>
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  do {
    __nanosleep(NS);
#ifndef VFI_WATCHDOG
    // default build: a protocol bug traps (launch failure) after 2^22 spins (<= 5.5 s) -- costs nothing measurable.  The
    // recording watchdog below is 2.6 % slower on the forward kernel; build it with scripts/build_variant.sh wd -DVFI_WATCHDOG
    // and load it through VFI_B200_LIB when a pipeline needs debugging (that is how the parity aliasing of the weight-
    // gradient kernel was found).
    if (++spins > (1u << 22)) __trap();
#else
    if ((++spins & 1023u) == 0u) {
      const bool expired = spins >= (1u << 21);
      if (expired || *(volatile unsigned int*)&g_abort_flag) {
        if (expired && atomicExch(&g_abort_flag, 1u) == 0u) {
          g_abort_info[0] = ((unsigned long long)blockIdx.x << 32) | (threadIdx.x >> 5);
          g_abort_info[1] = bar;
          g_abort_info[2] = parity;
          __threadfence();
        }
        if ((g_abort_info[0] >> 32) == blockIdx.x) g_abort_info[4 + (threadIdx.x >> 5)] = bar | ((unsigned long long)parity << 32);
        atomicAdd(&g_abort_info[3], 1ull);
        return;
      }
    }
#endif
  } while (!mbar_try_wait(bar, parity, 1000u));
}
template <int NS>
__device__ __forceinline__ void mbar_wait_ns(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity, 2000u)) return;
  mbar_wait_slow<NS>(bar, parity);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_ns<64>(bar, parity); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 1-D bulk copy global -> shared (TMA engine, no tensor map), completion counted in bytes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// Shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor bits).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);       // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                        // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                        // layout type 2 = SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, both operands K-major (cute::UMMA::InstrDescriptor bits).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) /* D = f32 */ | (1u << 7) /* A = bf16 */ | (1u << 10) /* B = bf16 */ | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ operand packers
// K index (kb, kk) -> (tap, channel) of the weight it multiplies; channel -1 = zero padding / bias slot.
__host__ __device__ inline void v6_k_to_tap_channel(int kb, int kk, int& tap, int& c);   // dcn_tc6.cuh

__device__ __forceinline__ float load_bias(const void* bias, int bias_dtype, int i) {
  if (bias_dtype == VFI_F32) return reinterpret_cast<const float*>(bias)[i];
  if (bias_dtype == VFI_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(bias)[i]);
  return __half2float(reinterpret_cast<const __half*>(bias)[i]);
}

// the K order of the v6 / v7 kernels (10 blocks; block 10 of the 11-block image is left zero)
template <typename TW>
__global__ void pack_weight_kernel(const TW* __restrict__ w, const void* bias, int bias_dtype, int O, int C,
                                   uint8_t* __restrict__ packed, float* __restrict__ bias_out, int variant) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < TC_KBLOCKS * TC_N * 64) {
    int kb = idx / (TC_N * 64), o = (idx / 64) % TC_N, kk = idx % 64;
    int tap, c;
    v6_k_to_tap_channel(kb, kk, tap, c);
    float v = 0.0f;
    if (o < O && c >= 0 && c < C) v = to_f32<TW>(w[((size_t)o * C + c) * 9 + tap]);
    if (kb == 9 && (kk == 36 || kk == 37) && o < O && bias) {
      // v6 adds the bias on the tensor core: A holds 1.0 at K elements 36 and 37 of the tail block, the weight image the
      // bias split into two bf16 terms (hi + lo carries 16 significant bits)
      const float bv = load_bias(bias, bias_dtype, o);
      const float hi = __bfloat162float(__float2bfloat16_rn(bv));
      v = kk == 36 ? hi : bv - hi;
    }
    size_t off = (size_t)kb * TC_B_BYTES + (size_t)o * 128 + ((((kk >> 3) ^ (o & 7))) << 4) + (kk & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(packed + off) = __float2bfloat16_rn(v);
  }
  if (bias_out && idx < TC_N) {
    const float bv = (bias && idx < O) ? load_bias(bias, bias_dtype, idx) : 0.0f;
    bias_out[idx] = bv;
  }
}

// Any strided [B,C,H,W] tensor -> main plane [P][64] + tail plane [P][8] (bf16, zero padded).
template <typename TX>
__global__ void pack_input_kernel(const TX* __restrict__ x, long long sn, long long sc, long long sh, long long sw, int B,
                                  int C, int H, int W, __nv_bfloat16* __restrict__ main_plane,
                                  __nv_bfloat16* __restrict__ tail_plane) {
  // one thread per (pixel, 8-channel chunk); lanes run over pixels so NCHW reads coalesce -- over chunks when the channel
  // is the unit-stride dimension (channels_last), so a warp reads a contiguous run of ~3.5 pixel rows
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long npix = (long long)B * H * W;
  if (idx >= npix * (TC_CMAX / 8)) return;
  int chunk = sc == 1 ? (int)(idx % (TC_CMAX / 8)) : (int)(idx / npix);
  long long pix = sc == 1 ? idx / (TC_CMAX / 8) : idx % npix;
  int xx = (int)(pix % W);
  long long t = pix / W;
  int y = (int)(t % H);
  int b = (int)(t / H);
  const TX* src = x + b * sn + y * sh + xx * sw;
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int c = chunk * 8 + i;
    if (chunk == 8 && i >= 4 && C <= TC_CMAIN + 4) c -= 4;      // tail of <= 4 channels: upper half mirrors the lower
    v[i] = __float2bfloat16_rn(c < C ? to_f32<TX>(__ldg(src + c * sc)) : 0.0f);
  }
  __nv_bfloat16* dst = chunk < 8 ? main_plane + pix * TC_CMAIN + chunk * 8 : tail_plane + pix * TC_CTAIL;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(v);
}

// Same planes in fp32 (main [P][64], tail [P][8] = channels 64..71), for the fp32 column-gradient backward.
template <typename TX>
__global__ void pack_input_f32_kernel(const TX* __restrict__ x, long long sn, long long sc, long long sh, long long sw, int B,
                                      int C, int H, int W, float* __restrict__ main_plane, float* __restrict__ tail_plane) {
  constexpr int CH = TC_CMAX / 4;                           // 18 chunks of four channels
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long npix = (long long)B * H * W;
  if (idx >= npix * CH) return;
  int chunk = sc == 1 ? (int)(idx % CH) : (int)(idx / npix);
  long long pix = sc == 1 ? idx / CH : idx % npix;
  int xx = (int)(pix % W);
  long long t = pix / W;
  int y = (int)(t % H);
  int b = (int)(t / H);
  const TX* src = x + b * sn + y * sh + xx * sw;
  float4 v;
  float* f = reinterpret_cast<float*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = chunk * 4 + i;
    f[i] = c < C ? to_f32<TX>(__ldg(src + c * sc)) : 0.0f;
  }
  float* dst = chunk < 16 ? main_plane + pix * TC_CMAIN + chunk * 4 : tail_plane + pix * TC_CTAIL + (chunk - 16) * 4;
  *reinterpret_cast<float4*>(dst) = v;
}

// ------------------------------------------------------------------------------------------------ main kernel
struct TcParams {
  const uint8_t* x_main; const uint8_t* x_tail;   // planes (see file header)
  uint32_t main_stride, tail_stride;               // bytes per pixel (128 / 16 for dense planes)
  const void* offset; const void* mask;            // fused27: both point at the 27-channel offset_conv output
  long long f_sn, f_sc, f_sh, f_sw;
  long long m_sn, m_sc, m_sh, m_sw;
  int fused27;                                     // 1: offset ch j -> conv27[j < 9 ? j : j + 9], mask = sigmoid(conv27[9 + k])
  int geo_cl;                                      // fused27 input is dense channels-last ([P][27]): v6 copies tile rows as they lie
  const uint8_t* wpacked;                          // [11][80][128 B] swizzled
  const float* bias;                               // [80]
  void* out; void* out_tail;                       // out_tail != null: planes out (main [P][64], tail [P][8] bf16)
  long long o_sn, o_sc, o_sh, o_sw;                // generic strided output otherwise
  int out_rows;                                    // generic output is a dense channels-last 16-bit tensor ([P][O] rows): staged stores
  int B, H, W, O;
  int tiles_x, tiles_y, num_tiles;
  int experiment;                                  // VFI_DCN_EXPERIMENT (diagnostics only, results are wrong when non-zero)
  unsigned long long* debug;                       // optional [grid][32 warps][8] cycle counters (VFI_DCN_DEBUG), else null
};

__device__ __forceinline__ void unpack2(uint32_t v, float& lo, float& hi) {
  lo = __uint_as_float(v << 16);
  hi = __uint_as_float(v & 0xffff0000u);
}
__device__ __forceinline__ __nv_bfloat162 as_bf162(uint32_t v) { return *reinterpret_cast<__nv_bfloat162*>(&v); }

// One 16-byte chunk (8 channels) of the modulated bilinear sample from its four corner chunks.
// Fast: packed HMUL2/HFMA2.BF16 (every step rounds to bf16; measured cost on the layer output: 2.8e-3 vs 1.4e-3
// max-rel).  HQ: fp32 arithmetic, one rounding at the end.
__device__ __forceinline__ uint4 lerp_chunk(const uint4& a, const uint4& b, const uint4& c, const uint4& d, const uint2& w) {
  const __nv_bfloat162 w0 = as_bf162(__byte_perm(w.x, 0, 0x1010)), w1 = as_bf162(__byte_perm(w.x, 0, 0x3232));
  const __nv_bfloat162 w2 = as_bf162(__byte_perm(w.y, 0, 0x1010)), w3 = as_bf162(__byte_perm(w.y, 0, 0x3232));
  auto f = [&](uint32_t va, uint32_t vb, uint32_t vc, uint32_t vd) {
    __nv_bfloat162 r = __hmul2(w0, as_bf162(va));
    r = __hfma2(w1, as_bf162(vb), r);
    r = __hfma2(w2, as_bf162(vc), r);
    r = __hfma2(w3, as_bf162(vd), r);
    return *reinterpret_cast<uint32_t*>(&r);
  };
  return make_uint4(f(a.x, b.x, c.x, d.x), f(a.y, b.y, c.y, d.y), f(a.z, b.z, c.z, d.z), f(a.w, b.w, c.w, d.w));
}

__device__ __forceinline__ void store_geo_w(uint2& dst, float w0, float w1, float w2, float w3) {
  __nv_bfloat162 a = __floats2bfloat162_rn(w0, w1), b = __floats2bfloat162_rn(w2, w3);
  dst = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

// tile index -> (batch, top row, left column)
__device__ __forceinline__ void tile_origin(const TcParams& p, int tile, int& b, int& y0, int& x0) {
  int per_img = p.tiles_x * p.tiles_y;
  b = tile / per_img;
  int t = tile % per_img;
  y0 = (t / p.tiles_x) * TC_TH;
  x0 = (t % p.tiles_x) * TC_TW;
}

#include "dcn_tc6.cuh"   // v6: TMEM-resident A operand, source box staged in shared memory
#include "dcn_tc7.cuh"   // v7 (default): v6 + TMA tensor maps, tail channels gathered by the geometry warps, four producer groups
#include "dcn_tc6_wgrad.cuh"   // weight / bias gradient on tcgen05 (pixel-reduction GEMM, accumulators persistent in TMEM)
#include "dcn_bwd_cols.cuh"    // grad_x / grad_offset / grad_mask from the column gradient (channels-last, vector reductions)
#include "dcn_gcol.cuh"        // the column gradient itself: gcol = grad_out x W on tcgen05 (A in TMEM, weights resident in smem)

// ------------------------------------------------------------------------------------------------ UMMA self test
// D[128, 80] = A[128, K] * Bm[80, K]^T with A, Bm row-major bf16 in global memory, K a multiple of 64.  Uses exactly
// the descriptor / swizzle / tcgen05 helpers of the kernel above, serialised (one stage, block barriers), so that a
// wrong answer from the DCN kernel can be attributed either to the tensor-core plumbing or to the gather.
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const __nv_bfloat16* __restrict__ A,
                                                                const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D,
                                                                int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint8_t* sa = base;                       // 16384
  uint8_t* sb = base + TC_A_BYTES;          // 10240
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(base + TC_A_BYTES + TC_B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(smem_u32(bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = umma_idesc_bf16(TC_M, TC_N);
  uint32_t parity = 0;
  for (int kb = 0; kb < K / 64; ++kb) {
    for (int i = tid; i < TC_M * 8; i += 128) {
      int r = i >> 3, jj = i & 7;
      uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * K + kb * 64 + jj * 8);
      *reinterpret_cast<uint4*>(sa + r * 128 + ((jj ^ (r & 7)) << 4)) = v;
    }
    for (int i = tid; i < TC_N * 8; i += 128) {
      int r = i >> 3, jj = i & 7;
      uint4 v = *reinterpret_cast<const uint4*>(Bm + (size_t)r * K + kb * 64 + jj * 8);
      *reinterpret_cast<uint4*>(sb + r * 128 + ((jj ^ (r & 7)) << 4)) = v;
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t adesc = umma_desc_sw128(smem_u32(sa)), bdesc = umma_desc_sw128(smem_u32(sb));
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
      umma_commit(smem_u32(bar));
    }
    mbar_wait(smem_u32(bar), parity);
    parity ^= 1;
    tc_fence_after();
    __syncthreads();
  }
  uint32_t d[TC_N];
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
  for (int c = 0; c < TC_N / 16; ++c) tmem_ld16(taddr + c * 16, d + c * 16);
  tmem_ld_wait();
  const int row = warp * 32 + lane;
#pragma unroll
  for (int c = 0; c < TC_N; ++c) D[row * TC_N + c] = __uint_as_float(d[c]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

}  // namespace

// VFI_DCN_DEBUG=1: per-warp wait-cycle counters of the last launch, [256 CTAs][32 warps][8] u64 (diagnostics only).
unsigned long long* dcn_tc_debug_buffer() {
  static unsigned long long* buf = [] {
    unsigned long long* b = nullptr;
    const char* e = getenv("VFI_DCN_DEBUG");
    if (e && e[0] == '1' && cudaMalloc(&b, 256 * 32 * 8 * sizeof(unsigned long long)) != cudaSuccess) b = nullptr;
    return b;
  }();
  return buf;
}

// dense channels-last bf16 plane [B,H,W,C] with a pixel stride of exactly `stride` elements, 16-byte aligned
static bool is_plane(const vfi_tensor* t, long long stride) {
  return t->dtype == VFI_BF16 && t->sc == 1 && t->sw == stride && t->sh == t->w * stride && t->sn == t->h * t->w * stride &&
         aligned(t->data, 16);
}
// ... or with any pixel stride >= min_stride that keeps pixels 16-byte aligned (rows and images dense): e.g. the main / tail
// channel ranges of one [B,H,W,72] activation buffer.  The v7 kernel addresses planes through TMA tensor maps and takes these.
static bool is_strided_plane(const vfi_tensor* t, long long min_stride) {
  return t->dtype == VFI_BF16 && t->sc == 1 && t->sw >= min_stride && t->sw % 8 == 0 && t->sh == t->w * t->sw &&
         t->sn == t->h * t->w * t->sw && aligned(t->data, 16);
}

int dcn_tc_pack_weight(const void* weight, int weight_dtype, const void* bias, int bias_dtype, long long O, long long C,
                       void* packed, float* bias_out, cudaStream_t st, int variant) {
  VFI_REQUIRE(weight && packed, VFI_ERR_INVALID, "vfi_dcn_pack_weight: null pointer");
  VFI_REQUIRE(O > 0 && O <= TC_N && C > 0 && C <= TC_CMAX, VFI_ERR_UNSUPPORTED,
              "vfi_dcn_pack_weight: tensor-core path supports C <= %d, O <= %d (got C=%lld, O=%lld)", TC_CMAX, TC_N, C, O);
  VFI_REQUIRE(aligned(packed, 16), VFI_ERR_INVALID, "vfi_dcn_pack_weight: destination must be 16-byte aligned");
  const int n = TC_KBLOCKS * TC_N * 64;
  VFI_DISPATCH(weight_dtype, TW, {
    pack_weight_kernel<TW><<<ceil_div(n, 256), 256, 0, st>>>(reinterpret_cast<const TW*>(weight), bias, bias_dtype, (int)O,
                                                           (int)C, reinterpret_cast<uint8_t*>(packed), bias_out, variant);
  });
  VFI_LAUNCH_CHECK("pack_weight_kernel");
  return VFI_OK;
}

int dcn_tc_pack_input(const vfi_tensor* x, void* main_plane, void* tail_plane, cudaStream_t st) {
  VFI_REQUIRE(x && main_plane && tail_plane, VFI_ERR_INVALID, "vfi_dcn_pack_input: null pointer");
  VFI_REQUIRE(x->c > 0 && x->c <= TC_CMAX, VFI_ERR_UNSUPPORTED, "vfi_dcn_pack_input: C must be <= %d", TC_CMAX);
  VFI_REQUIRE(aligned(main_plane, 16) && aligned(tail_plane, 16), VFI_ERR_INVALID,
              "vfi_dcn_pack_input: destinations must be 16-byte aligned");
  long long total = (long long)x->n * x->h * x->w * (TC_CMAX / 8);
  if (total == 0) return VFI_OK;
  VFI_REQUIRE(x->data, VFI_ERR_INVALID, "vfi_dcn_pack_input: null data pointer");
  VFI_DISPATCH(x->dtype, TX, {
    pack_input_kernel<TX><<<ceil_div(total, 256), 256, 0, st>>>(
        reinterpret_cast<const TX*>(x->data), x->sn, x->sc, x->sh, x->sw, (int)x->n, (int)x->c, (int)x->h, (int)x->w,
        reinterpret_cast<__nv_bfloat16*>(main_plane), reinterpret_cast<__nv_bfloat16*>(tail_plane));
  });
  VFI_LAUNCH_CHECK("pack_input_kernel");
  return VFI_OK;
}


// 4-D tiled map over 16-bit elements: dims / box innermost first, strides (bytes) of dims 1..3.  False when the copy engine
// cannot describe the tensor (alignment, stride range) -- the caller then uses another path or reports VFI_ERR_UNSUPPORTED.
static bool make_map4(CUtensorMap* m, const void* base, const long long dim[4], const long long stride_bytes[3], const int box[4],
                      bool swizzle128) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc || !aligned(base, 16)) return false;
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    if (dim[i] <= 0 || box[i] <= 0 || box[i] > 256) return false;
    gd[i] = (cuuint64_t)dim[i];
    bx[i] = (cuuint32_t)box[i];
  }
  for (int i = 0; i < 3; ++i) {
    if (stride_bytes[i] <= 0 || stride_bytes[i] % 16 != 0 || stride_bytes[i] >= (1LL << 40)) return false;
    gs[i] = (cuuint64_t)stride_bytes[i];
  }
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// channels-last 16-bit plane [B,H,W,c] with dense rows (sh = W * sw), any pixel stride that is a multiple of 16 bytes
static bool plane_map(CUtensorMap* m, const void* base, long long c, long long px_bytes, long long B, long long H, long long W, int box_w,
                      int box_h, bool swizzle128) {
  const long long dim[4] = {c, W, H, B}, st[3] = {px_bytes, px_bytes * W, px_bytes * W * H};
  const int box[4] = {(int)c, box_w, box_h, 1};
  return make_map4(m, base, dim, st, box, swizzle128);
}
// NCHW-like 16-bit tensor (unit pixel stride): box of 16 x 8 pixels x all channels
static bool nchw_map(CUtensorMap* m, const vfi_tensor* t) {
  if (t->sw != 1 || t->sh <= 0 || t->sc <= 0 || t->sn <= 0) return false;
  const long long dim[4] = {t->w, t->h, t->c, t->n}, st[3] = {t->sh * 2, t->sc * 2, t->sn * 2};
  const int box[4] = {TC_TW, TC_TH, (int)t->c, 1};
  return make_map4(m, t->data, dim, st, box, false);
}

// Shared implementation.  x_tail == null: x_main is any [B,C,H,W] tensor (packed into planes in the workspace).
// x_tail != null: planes in.  conv27 != null selects the fused offset/mask form.  out_tail != null: planes out.
int dcn_tc_run(const vfi_tensor* x_main, const vfi_tensor* x_tail, const vfi_tensor* offset, const vfi_tensor* mask,
               const vfi_tensor* conv27, const void* weight, int weight_dtype, const void* bias, int bias_dtype,
               const vfi_tensor* out, const vfi_tensor* out_tail, long long O, bool hq, void* workspace,
               size_t workspace_bytes, cudaStream_t st, const char* who) {
  const vfi_tensor* x = x_main;
  VFI_REQUIRE(x && out && weight && (conv27 || (offset && mask)), VFI_ERR_INVALID, "%s: null argument", who);
  const long long C = x->c + (x_tail ? x_tail->c : 0);
  VFI_REQUIRE(C <= TC_CMAX && O <= TC_N && O > 0 && x->c > 0, VFI_ERR_UNSUPPORTED,
              "%s(bf16_tc): supports C <= %d and O <= %d (got C=%lld, O=%lld)", who, TC_CMAX, TC_N, C, O);
  if (conv27) {
    VFI_REQUIRE(conv27->n == x->n && conv27->c == 27 && conv27->h == x->h && conv27->w == x->w, VFI_ERR_INVALID,
                "%s: conv27 must be [B,27,H,W]", who);
    offset = mask = conv27;
  } else {
    VFI_REQUIRE(offset->n == x->n && offset->c == 18 && offset->h == x->h && offset->w == x->w, VFI_ERR_INVALID,
                "%s: offset must be [B,18,H,W]", who);
    VFI_REQUIRE(mask->n == x->n && mask->c == 9 && mask->h == x->h && mask->w == x->w, VFI_ERR_INVALID,
                "%s: mask must be [B,9,H,W]", who);
    VFI_REQUIRE(offset->dtype == mask->dtype, VFI_ERR_UNSUPPORTED, "%s: offset and mask must share a dtype", who);
  }
  const long long P = (long long)x->n * x->h * x->w;
  if (P == 0) return VFI_OK;
  VFI_REQUIRE(x->data && offset->data && mask->data && out->data, VFI_ERR_INVALID, "%s: null data pointer", who);
  VFI_REQUIRE(P < 2147483647LL / 2, VFI_ERR_UNSUPPORTED, "%s(bf16_tc): more than 2^30 pixels per call", who);
  VFI_REQUIRE(27 * llabs(offset->sc) < 2147483647LL && 27 * llabs(mask->sc) < 2147483647LL, VFI_ERR_UNSUPPORTED,
              "%s(bf16_tc): offset/mask channel stride too large for 32-bit indexing", who);

  TcParams p;
  if (out_tail) {
    VFI_REQUIRE(O > TC_CMAIN && O <= TC_CMAX && out->c == TC_CMAIN && out_tail->c == O - TC_CMAIN && out_tail->data &&
                    out->n == x->n && out->h == x->h && out->w == x->w && out_tail->n == x->n && out_tail->h == x->h &&
                    out_tail->w == x->w && is_strided_plane(out, TC_CMAIN) && is_strided_plane(out_tail, TC_CTAIL),
                VFI_ERR_UNSUPPORTED,
                "%s: plane output needs out [B,64,H,W] and out_tail [B,O-64,H,W] as channels-last bf16 planes "
                "(pixel strides >= 64 / 8 elements and multiples of 8, dense rows)", who);
    p.out = out->data; p.out_tail = out_tail->data;
    p.o_sn = p.o_sc = p.o_sh = p.o_sw = 0; p.out_rows = 0;
  } else {
    VFI_REQUIRE(out->n == x->n && out->c == O && out->h == x->h && out->w == x->w, VFI_ERR_INVALID,
                "%s: out must be [B,O,H,W]", who);
    p.out = out->data; p.out_tail = nullptr;
    p.o_sn = out->sn; p.o_sc = out->sc; p.o_sh = out->sh; p.o_sw = out->sw;
    // channels_last result of a 16-bit forward (what ops.py allocates): every tile row is one contiguous, 16-byte aligned
    // run of 16 x O elements, written from a staged copy of the tile instead of 2-byte stores 2 O bytes apart
    p.out_rows = (dtype_size(out->dtype) == 2 && out->sc == 1 && out->sw == O && out->sh == out->w * O && out->sn % 8 == 0 &&
                  out->w % 8 == 0 && O <= TC_CMAX && aligned(out->data, 16)) ? 1 : 0;
  }
  if (x_tail) {
    VFI_REQUIRE(x_tail->data && x->c == TC_CMAIN && x_tail->c <= TC_CTAIL && x_tail->n == x->n && x_tail->h == x->h &&
                    x_tail->w == x->w && is_strided_plane(x, TC_CMAIN) && is_strided_plane(x_tail, TC_CTAIL),
                VFI_ERR_UNSUPPORTED,
                "%s: plane input needs x_main [B,64,H,W] and x_tail [B,<=8,H,W] as channels-last bf16 planes "
                "(pixel strides >= 64 / 8 elements and multiples of 8, dense rows, pad channels of the tail zero)", who);
  }
  const bool dense_planes = (!x_tail || (is_plane(x, TC_CMAIN) && is_plane(x_tail, TC_CTAIL))) &&
                            (!out_tail || (is_plane(out, TC_CMAIN) && is_plane(out_tail, TC_CTAIL)));
  const size_t need = x_tail ? ws_main_off() : dcn_tc_workspace_bytes(x->n, x->h, x->w);
  VFI_REQUIRE(workspace && workspace_bytes >= need && aligned(workspace, 256), VFI_ERR_WORKSPACE,
              "%s(bf16_tc): workspace of %zu bytes (256-byte aligned) required, got %zu", who, need, workspace_bytes);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* bias_ws = reinterpret_cast<float*>(ws + ws_bias_off());
  // Kernel: v7 (dcn_tc7.cuh) -- packed-bf16 blend, at most four tail channels (C <= 68), 16-bit offsets / masks in any layout.
  // The reference kernel v6 brings the offsets / masks in with 16-byte bulk copies: unit pixel stride, 16-byte aligned rows
  auto bulk_ok = [](const vfi_tensor* t) {
    return (t->dtype == VFI_BF16 || t->dtype == VFI_F16) && t->sw == 1 && t->w % 8 == 0 && t->sh % 8 == 0 && t->sc % 8 == 0 &&
           t->sn % 8 == 0 && aligned(t->data, 16) && t->sh >= 0 && t->sc >= 0 && t->sn >= 0;
  };
  // ... or, for the fused form, a dense channels-last offset_conv output ([P][27], what a channels_last model hands over)
  const bool geo_cl = conv27 && (conv27->dtype == VFI_BF16 || conv27->dtype == VFI_F16) && conv27->sc == 1 && conv27->sw == 27 &&
                      conv27->sh == conv27->w * 27 && conv27->sn % 8 == 0 && conv27->w % 8 == 0 && aligned(conv27->data, 16);
  const bool use_v6 = !hq && dense_planes && C <= TC_CMAIN + 4 && (geo_cl || (bulk_ok(offset) && bulk_ok(mask)));
  const bool use_v7 = !hq && C <= TC_CMAIN + 4 && (offset->dtype == VFI_BF16 || offset->dtype == VFI_F16) &&
                      mask->dtype == offset->dtype;
  int rc = dcn_tc_pack_weight(weight, weight_dtype, bias, bias_dtype, O, C, ws, bias_ws, st, 6);
  if (rc) return rc;
  if (x_tail) {
    p.x_main = reinterpret_cast<const uint8_t*>(x->data);
    p.x_tail = reinterpret_cast<const uint8_t*>(x_tail->data);
  } else {
    rc = dcn_tc_pack_input(x, ws + ws_main_off(), ws + ws_tail_off(P), st);
    if (rc) return rc;
    p.x_main = ws + ws_main_off();
    p.x_tail = ws + ws_tail_off(P);
  }
  p.main_stride = x_tail ? (uint32_t)(x->sw * 2) : TC_CMAIN * 2; p.tail_stride = x_tail ? (uint32_t)(x_tail->sw * 2) : TC_CTAIL * 2;
  p.offset = offset->data; p.mask = mask->data; p.fused27 = conv27 ? 1 : 0; p.geo_cl = (use_v6 && geo_cl) ? 1 : 0;
  p.f_sn = offset->sn; p.f_sc = offset->sc; p.f_sh = offset->sh; p.f_sw = offset->sw;
  p.m_sn = mask->sn; p.m_sc = mask->sc; p.m_sh = mask->sh; p.m_sw = mask->sw;
  p.wpacked = ws; p.bias = bias_ws;
  p.B = (int)x->n; p.H = (int)x->h; p.W = (int)x->w; p.O = (int)O;
  p.tiles_x = ceil_div(x->w, TC_TW); p.tiles_y = ceil_div(x->h, TC_TH);
  p.num_tiles = p.B * p.tiles_x * p.tiles_y;
  p.debug = dcn_tc_debug_buffer();
  static const int experiment = [] { const char* e = getenv("VFI_DCN_EXPERIMENT"); return e ? atoi(e) : 0; }();
  p.experiment = experiment;                          // only read by the debug-counter instantiation (VFI_DCN_DEBUG=1)
  int dev = 0, sms = 148;
  VFI_CUDA(cudaGetDevice(&dev));
  VFI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  const int out_dtype = out_tail ? VFI_BF16 : out->dtype;
  // v7 (default): v6's arithmetic with TMA tensor maps, the tail channels gathered by the geometry warps and four producer
  // groups (dcn_tc7.cuh).  VFI_DCN_KERNEL=v6 selects the previous kernel (A/B runs).
  static const bool force_v6 = [] { const char* e = getenv("VFI_DCN_KERNEL"); return e && e[0] == 'v' && e[1] == '6'; }();
  VFI_REQUIRE(dense_planes || (use_v7 && !force_v6), VFI_ERR_UNSUPPORTED,
              "%s: planes with a pixel stride other than 64 / 8 elements need the default (v7) tensor-core kernel", who);
  if (use_v7 && !force_v6) {
    V7Args a;
    a.p = p;
    a.o_main_px = out_tail ? (uint32_t)(out->sw * 2) : TC_CMAIN * 2; a.o_tail_px = out_tail ? (uint32_t)(out_tail->sw * 2) : TC_CTAIL * 2;
    bool ok = plane_map(&a.tm_main, p.x_main, TC_CMAIN, p.main_stride, p.B, p.H, p.W, V6_BOX_W, V6_BOX_H, false) &&
              plane_map(&a.tm_tail, p.x_tail, TC_CTAIL, p.tail_stride, p.B, p.H, p.W, V6_BOX_W, V6_BOX_H, false);
    VFI_REQUIRE(ok, VFI_ERR_UNSUPPORTED, "%s(bf16_tc): the activation planes cannot be described by a TMA tensor map", who);
    a.tm_raw0 = a.tm_main; a.tm_raw1 = a.tm_main; a.tm_out = a.tm_main;     // defined contents for the maps a mode does not use
    static const int force_raw = [] { const char* e = getenv("VFI_DCN_RAW"); return e ? atoi(e) : -1; }();
    if (geo_cl && force_raw != V7_RAW_LDG) a.raw_mode = V7_RAW_ROWS;
    else if (force_raw != V7_RAW_LDG && (conv27 ? nchw_map(&a.tm_raw0, conv27) : (nchw_map(&a.tm_raw0, offset) && nchw_map(&a.tm_raw1, mask))))
      a.raw_mode = V7_RAW_TMA;
    else a.raw_mode = V7_RAW_LDG;
    a.p.geo_cl = a.raw_mode == V7_RAW_ROWS ? 1 : 0;
    static const bool no_tma_store = [] { const char* e = getenv("VFI_DCN_NO_TMA_STORE"); return e && e[0] == '1'; }();
    a.use_tma_store = (out_tail && !no_tma_store &&
                       plane_map(&a.tm_out, p.out, TC_CMAIN, a.o_main_px, p.B, p.H, p.W, TC_TW, TC_TH, true)) ? 1 : 0;
    const size_t smem7 = sizeof(V7Smem) + 1024;
    const bool dbg = p.debug != nullptr;
#define VFI_V7_LAUNCH(FUSED, PLANES, DBG)                                                                      \
  do {                                                                                                         \
    auto kern = dcn_tc7_fwd_kernel<TO, TOUT, FUSED, PLANES, DBG>;                                              \
    VFI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem7));             \
    kern<<<grid, V7_THREADS, smem7, st>>>(a);                                                                  \
  } while (0)
    // Two forms are instantiated -- the ones the two entry points produce: the fused form (27-channel offset_conv input) with
    // planes out (vfi_dcn_fwd_fused through ops.deform_conv2d_fused / HotPath / the fused drop-in), and the torchvision form
    // (offset + mask) with a strided tensor out (vfi_dcn_fwd, the generic drop-in).
    VFI_REQUIRE((p.fused27 != 0) == (out_tail != nullptr), VFI_ERR_UNSUPPORTED,
                "%s(bf16_tc): the fused (conv27) form writes planes (out_tail), the offset + mask form a [B,O,H,W] tensor", who);
    if (out_tail) {
      using TOUT = __nv_bfloat16;
      if (offset->dtype == VFI_BF16) {
        using TO = __nv_bfloat16;
        if (dbg) VFI_V7_LAUNCH(true, true, true); else VFI_V7_LAUNCH(true, true, false);
      } else {
        using TO = __half;
        VFI_V7_LAUNCH(true, true, false);
      }
    } else {
      VFI_DISPATCH(out_dtype, TOUT, {
        if (offset->dtype == VFI_BF16) {
          using TO = __nv_bfloat16;
          VFI_V7_LAUNCH(false, false, false);
        } else {
          using TO = __half;
          VFI_V7_LAUNCH(false, false, false);
        }
      });
    }
#undef VFI_V7_LAUNCH
    VFI_LAUNCH_CHECK("dcn_tc7_fwd_kernel");
    return VFI_OK;
  }
  // v6, the round-1 kernel, is kept as the bit-exactness reference of v7: one instantiation (bf16 offsets, planes in and out,
  // 27-channel offset_conv input), reached only through VFI_DCN_KERNEL=v6.
  if (force_v6 && use_v6 && out_tail && p.fused27 && offset->dtype == VFI_BF16) {
    const size_t smem6 = sizeof(V6Smem) + 1024;
    auto kern = dcn_tc6_fwd_kernel<__nv_bfloat16, __nv_bfloat16, true, true, false>;
    VFI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6));
    kern<<<grid, V6_THREADS, smem6, st>>>(p);
    VFI_LAUNCH_CHECK("dcn_tc6_fwd_kernel");
    return VFI_OK;
  }
  VFI_REQUIRE(false, VFI_ERR_UNSUPPORTED,
              "%s(bf16_tc): needs 16-bit offset / mask tensors of one dtype and C <= %d (got C=%lld, offset dtype %d%s); the fp32-blend "
              "'bf16_tc_hq' mode and the v4 kernel were removed in round 2", who, TC_CMAIN + 4, C, (int)offset->dtype,
              force_v6 ? "; VFI_DCN_KERNEL=v6 only takes the fused bf16 planes form" : "");
}

int dcn_tc_fwd(const vfi_tensor* x, const vfi_tensor* offset, const vfi_tensor* mask, const void* weight, int weight_dtype,
               const void* bias, int bias_dtype, const vfi_tensor* out, long long O, bool hq, void* workspace,
               size_t workspace_bytes, cudaStream_t st) {
  return dcn_tc_run(x, nullptr, offset, mask, nullptr, weight, weight_dtype, bias, bias_dtype, out, nullptr, O, hq,
                    workspace, workspace_bytes, st, "vfi_dcn_fwd");
}

int dcn_tc_fwd_fused(const vfi_tensor* x_main, const vfi_tensor* x_tail, const vfi_tensor* conv27, const void* weight,
                     int weight_dtype, const void* bias, int bias_dtype, const vfi_tensor* out, const vfi_tensor* out_tail,
                     long long O, bool hq, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  VFI_REQUIRE(conv27, VFI_ERR_INVALID, "vfi_dcn_fwd_fused: null conv27");
  return dcn_tc_run(x_main, x_tail, nullptr, nullptr, conv27, weight, weight_dtype, bias, bias_dtype, out, out_tail, O, hq,
                    workspace, workspace_bytes, st, "vfi_dcn_fwd_fused");
}

int umma_selftest(const void* A, const void* Bm, float* D, int K, cudaStream_t st) {
  VFI_REQUIRE(A && Bm && D && K > 0 && K % 64 == 0, VFI_ERR_INVALID, "vfi_selftest_umma: need A, B, D and K %% 64 == 0");
  const size_t smem = TC_A_BYTES + TC_B_BYTES + 64 + 1024;
  VFI_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(A),
                                             reinterpret_cast<const __nv_bfloat16*>(Bm), D, K);
  VFI_LAUNCH_CHECK("umma_selftest_kernel");
  return VFI_OK;
}

// Host-only: which (tap, channel) the weight image multiplies at K element kk of block kb (channel -1 = zero / bias slot).
int dcn_tc_k_order(int variant, int kb, int kk, int* tap, int* channel) {
  VFI_REQUIRE(tap && channel && kb >= 0 && kb < TC_KBLOCKS && kk >= 0 && kk < 64 && variant == 6, VFI_ERR_INVALID,
              "vfi_dcn_k_order: variant 6 (the K order of the v6 / v7 kernels), kb in [0,11), kk in [0,64)");
  v6_k_to_tap_channel(kb, kk, *tap, *channel);
  return VFI_OK;
}

// grad_weight / grad_bias on the tensor cores.  x: any [B,C,H,W] tensor (packed to planes in the workspace) -- the same
// operands the forward saw; grad_out [B,O,H,W] bf16 or f32 with unit pixel stride.  Accumulates into gw / gb (fp32).
int dcn_tc_bwd_weight(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* offset, const vfi_tensor* mask,
                      long long O, float* gw, float* gb, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const char* who = "vfi_dcn_bwd_weight_tc";
  VFI_REQUIRE(grad_out && x && offset && mask, VFI_ERR_INVALID, "%s: null tensor descriptor", who);
  const long long C = x->c;
  VFI_REQUIRE(C > 0 && C <= TC_CMAIN + 4 && O > 0 && O <= TC_M, VFI_ERR_UNSUPPORTED, "%s: supports C <= %d, O <= %d", who,
              TC_CMAIN + 4, TC_M);
  VFI_REQUIRE(offset->n == x->n && offset->c == 18 && offset->h == x->h && offset->w == x->w && mask->n == x->n && mask->c == 9 &&
                  mask->h == x->h && mask->w == x->w && grad_out->n == x->n && grad_out->c == O && grad_out->h == x->h &&
                  grad_out->w == x->w, VFI_ERR_INVALID, "%s: shape mismatch", who);
  const long long P = (long long)x->n * x->h * x->w;
  if (P == 0 || (!gw && !gb)) return VFI_OK;
  auto bulk_ok = [](const vfi_tensor* t, int es) {
    const int a = 16 / es;
    return t->sw == 1 && t->w % a == 0 && t->sh % a == 0 && t->sc % a == 0 && t->sn % a == 0 && aligned(t->data, 16) && t->sh >= 0 &&
           t->sc >= 0 && t->sn >= 0;
  };
  VFI_REQUIRE((offset->dtype == VFI_BF16 || offset->dtype == VFI_F16) && offset->dtype == mask->dtype && bulk_ok(offset, 2) &&
                  bulk_ok(mask, 2), VFI_ERR_UNSUPPORTED,
              "%s: offset / mask must be 16-bit tensors with unit pixel stride and 16-byte aligned rows (W %% 8 == 0)", who);
  VFI_REQUIRE(grad_out->dtype == VFI_BF16 || grad_out->dtype == VFI_F32, VFI_ERR_UNSUPPORTED, "%s: grad_out must be bf16 or f32",
              who);
  VFI_REQUIRE(x->w % 8 == 0, VFI_ERR_UNSUPPORTED, "%s: W must be a multiple of 8", who);
  VFI_REQUIRE(P < 2147483647LL / 2, VFI_ERR_UNSUPPORTED, "%s: more than 2^30 pixels per call", who);
  const size_t need = dcn_tc_workspace_bytes(x->n, x->h, x->w);
  VFI_REQUIRE(workspace && workspace_bytes >= need && aligned(workspace, 256), VFI_ERR_WORKSPACE,
              "%s: workspace of %zu bytes (256-byte aligned) required, got %zu", who, need, workspace_bytes);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int rc = dcn_tc_pack_input(x, ws + ws_main_off(), ws + ws_tail_off(P), st);
  if (rc) return rc;
  WgParams q;
  TcParams& p = q.t;
  p.x_main = ws + ws_main_off(); p.x_tail = ws + ws_tail_off(P);
  p.main_stride = TC_CMAIN * 2; p.tail_stride = TC_CTAIL * 2;
  p.offset = offset->data; p.mask = mask->data; p.fused27 = 0; p.geo_cl = 0;
  p.f_sn = offset->sn; p.f_sc = offset->sc; p.f_sh = offset->sh; p.f_sw = offset->sw;
  p.m_sn = mask->sn; p.m_sc = mask->sc; p.m_sh = mask->sh; p.m_sw = mask->sw;
  p.wpacked = nullptr; p.bias = nullptr; p.out = nullptr; p.out_tail = nullptr;
  p.o_sn = p.o_sc = p.o_sh = p.o_sw = 0; p.out_rows = 0;
  p.B = (int)x->n; p.H = (int)x->h; p.W = (int)x->w; p.O = (int)O;
  p.tiles_x = ceil_div(x->w, TC_TW); p.tiles_y = ceil_div(x->h, TC_TH);
  p.num_tiles = p.B * p.tiles_x * p.tiles_y;
  p.experiment = 0; p.debug = nullptr;
  q.gout = grad_out->data; q.g_sn = grad_out->sn; q.g_sc = grad_out->sc; q.g_sh = grad_out->sh; q.g_sw = grad_out->sw;
  q.g_vec = (grad_out->dtype == VFI_BF16 && bulk_ok(grad_out, 2)) ? 1 : 0;
  q.gw = gw; q.gb = gb; q.C = (int)C;
  int dev = 0, sms = 148;
  VFI_CUDA(cudaGetDevice(&dev));
  VFI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  const size_t smem = sizeof(WgSmem) + 1024;
  for (int pass = 0; pass < 2; ++pass) {
    q.pass = pass;
#define VFI_WG_LAUNCH(TO, TG)                                                                         \
  do {                                                                                                \
    auto kern = dcn_tc6_wgrad_kernel<TO, TG, false>;                                                  \
    VFI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    kern<<<grid, V6_THREADS, smem, st>>>(q);                                                          \
  } while (0)
    if (offset->dtype == VFI_BF16) {
      if (grad_out->dtype == VFI_BF16) VFI_WG_LAUNCH(__nv_bfloat16, __nv_bfloat16); else VFI_WG_LAUNCH(__nv_bfloat16, float);
    } else {
      if (grad_out->dtype == VFI_BF16) VFI_WG_LAUNCH(__half, __nv_bfloat16); else VFI_WG_LAUNCH(__half, float);
    }
#undef VFI_WG_LAUNCH
    VFI_LAUNCH_CHECK("dcn_tc6_wgrad_kernel");
  }
  return VFI_OK;
}

// The same with offsets / mask taken from the raw 27-channel offset_conv output (the backward of vfi_dcn_fwd_fused).
int dcn_tc_bwd_weight_fused(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* conv27, long long O, float* gw, float* gb,
                            void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const char* who = "vfi_dcn_bwd_weight_tc_fused";
  VFI_REQUIRE(grad_out && x && conv27, VFI_ERR_INVALID, "%s: null tensor descriptor", who);
  const long long C = x->c;
  VFI_REQUIRE(C > 0 && C <= TC_CMAIN + 4 && O > 0 && O <= TC_M, VFI_ERR_UNSUPPORTED, "%s: supports C <= %d, O <= %d", who,
              TC_CMAIN + 4, TC_M);
  VFI_REQUIRE(conv27->n == x->n && conv27->c == 27 && conv27->h == x->h && conv27->w == x->w && grad_out->n == x->n && grad_out->c == O &&
                  grad_out->h == x->h && grad_out->w == x->w, VFI_ERR_INVALID, "%s: shape mismatch", who);
  const long long P = (long long)x->n * x->h * x->w;
  if (P == 0 || (!gw && !gb)) return VFI_OK;
  auto bulk_ok = [](const vfi_tensor* t, int es) {
    const int a = 16 / es;
    return t->sw == 1 && t->w % a == 0 && t->sh % a == 0 && t->sc % a == 0 && t->sn % a == 0 && aligned(t->data, 16) && t->sh >= 0 &&
           t->sc >= 0 && t->sn >= 0;
  };
  // the staged-box geometry role streams the raw values with bulk copies: NCHW rows, or dense channels-last pixels ([P][27])
  const bool geo_cl = conv27->sc == 1 && conv27->sw == 27 && conv27->sh == conv27->w * 27 && conv27->sn == conv27->h * conv27->sh &&
                      aligned(conv27->data, 16) && (conv27->w * 27) % 8 == 0;
  VFI_REQUIRE((conv27->dtype == VFI_BF16 || conv27->dtype == VFI_F16) && (geo_cl || bulk_ok(conv27, 2)), VFI_ERR_UNSUPPORTED,
              "%s: conv27 must be a 16-bit tensor, NCHW with 16-byte aligned rows (W %% 8 == 0) or dense channels-last", who);
  VFI_REQUIRE(grad_out->dtype == VFI_BF16 || grad_out->dtype == VFI_F32, VFI_ERR_UNSUPPORTED, "%s: grad_out must be bf16 or f32",
              who);
  VFI_REQUIRE(x->w % 8 == 0, VFI_ERR_UNSUPPORTED, "%s: W must be a multiple of 8", who);
  VFI_REQUIRE(P < 2147483647LL / 2, VFI_ERR_UNSUPPORTED, "%s: more than 2^30 pixels per call", who);
  const size_t need = dcn_tc_workspace_bytes(x->n, x->h, x->w);
  VFI_REQUIRE(workspace && workspace_bytes >= need && aligned(workspace, 256), VFI_ERR_WORKSPACE,
              "%s: workspace of %zu bytes (256-byte aligned) required, got %zu", who, need, workspace_bytes);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int rc = dcn_tc_pack_input(x, ws + ws_main_off(), ws + ws_tail_off(P), st);
  if (rc) return rc;
  WgParams q;
  TcParams& p = q.t;
  p.x_main = ws + ws_main_off(); p.x_tail = ws + ws_tail_off(P);
  p.main_stride = TC_CMAIN * 2; p.tail_stride = TC_CTAIL * 2;
  p.offset = conv27->data; p.mask = conv27->data; p.fused27 = 1; p.geo_cl = geo_cl ? 1 : 0;
  p.f_sn = conv27->sn; p.f_sc = conv27->sc; p.f_sh = conv27->sh; p.f_sw = conv27->sw;
  p.m_sn = conv27->sn; p.m_sc = conv27->sc; p.m_sh = conv27->sh; p.m_sw = conv27->sw;
  p.wpacked = nullptr; p.bias = nullptr; p.out = nullptr; p.out_tail = nullptr;
  p.o_sn = p.o_sc = p.o_sh = p.o_sw = 0; p.out_rows = 0;
  p.B = (int)x->n; p.H = (int)x->h; p.W = (int)x->w; p.O = (int)O;
  p.tiles_x = ceil_div(x->w, TC_TW); p.tiles_y = ceil_div(x->h, TC_TH);
  p.num_tiles = p.B * p.tiles_x * p.tiles_y;
  p.experiment = 0; p.debug = nullptr;
  q.gout = grad_out->data; q.g_sn = grad_out->sn; q.g_sc = grad_out->sc; q.g_sh = grad_out->sh; q.g_sw = grad_out->sw;
  q.g_vec = (grad_out->dtype == VFI_BF16 && bulk_ok(grad_out, 2)) ? 1 : 0;
  q.gw = gw; q.gb = gb; q.C = (int)C;
  int dev = 0, sms = 148;
  VFI_CUDA(cudaGetDevice(&dev));
  VFI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  const size_t smem = sizeof(WgSmem) + 1024;
  for (int pass = 0; pass < 2; ++pass) {
    q.pass = pass;
#define VFI_WG_LAUNCH(TO, TG)                                                                         \
  do {                                                                                                \
    auto kern = dcn_tc6_wgrad_kernel<TO, TG, true>;                                                   \
    VFI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    kern<<<grid, V6_THREADS, smem, st>>>(q);                                                          \
  } while (0)
    if (conv27->dtype == VFI_BF16) {
      if (grad_out->dtype == VFI_BF16) VFI_WG_LAUNCH(__nv_bfloat16, __nv_bfloat16); else VFI_WG_LAUNCH(__nv_bfloat16, float);
    } else {
      if (grad_out->dtype == VFI_BF16) VFI_WG_LAUNCH(__half, __nv_bfloat16); else VFI_WG_LAUNCH(__half, float);
    }
#undef VFI_WG_LAUNCH
    VFI_LAUNCH_CHECK("dcn_tc6_wgrad_kernel<fused27>");
  }
  return VFI_OK;
}

// grad_x / grad_offset / grad_mask from the column gradient gcol = grad_out x W (see dcn_bwd_cols.cuh).
static size_t cols_tail_off(long long P, size_t es) { return (((size_t)P * TC_CMAIN * es + 255) / 256) * 256; }
size_t dcn_tc_bwd_data_cols_workspace_bytes(long long B, long long H, long long W, int gcol_dtype) {
  const long long P = B * H * W;
  const size_t es = gcol_dtype == VFI_F32 ? 4 : 2;
  return cols_tail_off(P, es) + (((size_t)P * TC_CTAIL * es + 255) / 256) * 256 + 256;
}

// Column gradient gcol[P][648] (bf16, column k * 72 + c) = grad_out[P][O] x W[O][C][k] on the tensor cores (dcn_gcol.cuh).
// Workspace: the 12-atom weight image.
size_t dcn_tc_gcol_workspace_bytes() { return (size_t)GC_ATOMS * TC_B_BYTES + 256; }

int dcn_tc_gcol(const vfi_tensor* grad_out, const void* weight, int weight_dtype, long long C, void* gcol, long long gcol_ld,
                void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const char* who = "vfi_dcn_gcol";
  VFI_REQUIRE(grad_out && weight && gcol, VFI_ERR_INVALID, "%s: null pointer", who);
  const long long O = grad_out->c;
  VFI_REQUIRE(O > 0 && O <= 68 && C > 0 && C <= TC_CMAX, VFI_ERR_UNSUPPORTED, "%s: supports O <= 68, C <= %d (got O=%lld, C=%lld)", who,
              TC_CMAX, O, C);
  VFI_REQUIRE(gcol_ld == GC_TAPS * GC_LD && aligned(gcol, 16), VFI_ERR_INVALID, "%s: gcol must be a 16-byte aligned [P][%d] bf16 matrix",
              who, GC_TAPS * GC_LD);
  const long long P = (long long)grad_out->n * grad_out->h * grad_out->w;
  if (P == 0) return VFI_OK;
  VFI_REQUIRE(grad_out->data, VFI_ERR_INVALID, "%s: null data pointer", who);
  VFI_REQUIRE(P < 2147483647LL - TC_M, VFI_ERR_UNSUPPORTED, "%s: more than 2^31 pixels per call", who);
  VFI_REQUIRE(workspace && workspace_bytes >= dcn_tc_gcol_workspace_bytes() && aligned(workspace, 256), VFI_ERR_WORKSPACE,
              "%s: workspace of %zu bytes (256-byte aligned) required, got %zu", who, dcn_tc_gcol_workspace_bytes(), workspace_bytes);
  uint8_t* wimg = reinterpret_cast<uint8_t*>(workspace);
  const int n = GC_ATOMS * TC_N * 64;
  VFI_DISPATCH(weight_dtype, TW, {
    pack_gcol_weight_kernel<TW><<<ceil_div(n, 256), 256, 0, st>>>(reinterpret_cast<const TW*>(weight), (int)O, (int)C, wimg);
  });
  VFI_LAUNCH_CHECK("pack_gcol_weight_kernel");
  GcArgs q;
  q.gout = grad_out->data; q.g_sn = grad_out->sn; q.g_sc = grad_out->sc; q.g_sh = grad_out->sh; q.g_sw = grad_out->sw;
  q.wimg = wimg;
  q.B = (int)grad_out->n; q.H = (int)grad_out->h; q.W = (int)grad_out->w; q.O = (int)O;
  q.P = P;
  q.num_tiles = ceil_div(P, TC_M);
  const long long dim[4] = {gcol_ld, P, 1, 1}, stb[3] = {gcol_ld * 2, gcol_ld * 2 * P, gcol_ld * 2 * P};
  const int box[4] = {GC_LD, TC_M, 1, 1};
  VFI_REQUIRE(make_map4(&q.tm_gcol, gcol, dim, stb, box, false), VFI_ERR_UNSUPPORTED, "%s: gcol cannot be described by a TMA tensor map", who);
  int dev = 0, sms = 148;
  VFI_CUDA(cudaGetDevice(&dev));
  VFI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = q.num_tiles < sms ? q.num_tiles : sms;
  const size_t smem = sizeof(GcSmem) + 1024;
  VFI_DISPATCH(grad_out->dtype, TG, {
    auto kern = dcn_gcol_gemm_kernel<TG>;
    VFI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, GC_THREADS, smem, st>>>(q);
  });
  VFI_LAUNCH_CHECK("dcn_gcol_gemm_kernel");
  return VFI_OK;
}

int dcn_tc_bwd_data_cols(const void* gcol, int gcol_dtype, long long gcol_ld, const vfi_tensor* x, const vfi_tensor* offset,
                         const vfi_tensor* mask, float* gx_rows, long long gx_ld, const vfi_tensor* grad_offset,
                         const vfi_tensor* grad_mask, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const char* who = "vfi_dcn_bwd_data_cols";
  VFI_REQUIRE(gcol && x && offset && mask, VFI_ERR_INVALID, "%s: null pointer", who);
  VFI_REQUIRE(gcol_dtype == VFI_BF16 || gcol_dtype == VFI_F32, VFI_ERR_UNSUPPORTED, "%s: gcol must be bf16 or f32", who);
  VFI_REQUIRE(x->c > 0 && x->c <= TC_CMAIN + 4, VFI_ERR_UNSUPPORTED, "%s: supports C <= %d", who, TC_CMAIN + 4);
  VFI_REQUIRE(offset->n == x->n && offset->c == 18 && offset->h == x->h && offset->w == x->w && mask->n == x->n && mask->c == 9 &&
                  mask->h == x->h && mask->w == x->w, VFI_ERR_INVALID, "%s: shape mismatch", who);
  VFI_REQUIRE(offset->dtype == mask->dtype, VFI_ERR_UNSUPPORTED, "%s: offset and mask must share a dtype", who);
  VFI_REQUIRE(gcol_ld >= 9 * BC_TAP_LD && gcol_ld % 4 == 0 && aligned(gcol, 16), VFI_ERR_INVALID,
              "%s: gcol rows must hold 9 x %d columns, row length a multiple of 4, 16-byte aligned", who, BC_TAP_LD);
  if (gx_rows) VFI_REQUIRE(gx_ld >= TC_CMAIN + 4 && gx_ld % 4 == 0 && aligned(gx_rows, 16), VFI_ERR_INVALID,
                           "%s: grad_x rows must hold >= %d floats, 16-byte aligned", who, TC_CMAIN + 4);
  if (grad_offset) VFI_REQUIRE(grad_offset->data && grad_offset->dtype == VFI_F32 && same_shape(grad_offset, offset), VFI_ERR_INVALID,
                               "%s: grad_offset must be f32 [B,18,H,W]", who);
  if (grad_mask) VFI_REQUIRE(grad_mask->data && grad_mask->dtype == VFI_F32 && same_shape(grad_mask, mask), VFI_ERR_INVALID,
                             "%s: grad_mask must be f32 [B,9,H,W]", who);
  const long long P = (long long)x->n * x->h * x->w;
  if (P == 0 || (!gx_rows && !grad_offset && !grad_mask)) return VFI_OK;
  VFI_REQUIRE(P < 2147483647LL / 2 && x->n <= 65535, VFI_ERR_UNSUPPORTED, "%s: more than 2^30 pixels or 65535 images per call", who);
  const size_t need = dcn_tc_bwd_data_cols_workspace_bytes(x->n, x->h, x->w, gcol_dtype);
  VFI_REQUIRE(workspace && workspace_bytes >= need && aligned(workspace, 256), VFI_ERR_WORKSPACE,
              "%s: workspace of %zu bytes (256-byte aligned) required, got %zu", who, need, workspace_bytes);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const size_t es = gcol_dtype == VFI_F32 ? 4 : 2;
  uint8_t* main_plane = ws;
  uint8_t* tail_plane = ws + cols_tail_off(P, es);
  if (gcol_dtype == VFI_F32) {
    VFI_REQUIRE(x->data, VFI_ERR_INVALID, "%s: null data pointer", who);
    const long long total = P * (TC_CMAX / 4);
    VFI_DISPATCH(x->dtype, TX, {
      pack_input_f32_kernel<TX><<<ceil_div(total, 256), 256, 0, st>>>(
          reinterpret_cast<const TX*>(x->data), x->sn, x->sc, x->sh, x->sw, (int)x->n, (int)x->c, (int)x->h, (int)x->w,
          reinterpret_cast<float*>(main_plane), reinterpret_cast<float*>(tail_plane));
    });
    VFI_LAUNCH_CHECK("pack_input_f32_kernel");
  } else {
    int rc = dcn_tc_pack_input(x, main_plane, tail_plane, st);
    if (rc) return rc;
  }
  BcParams q{};
  q.x_main = main_plane; q.x_tail = tail_plane;
  q.main_px = (long long)(TC_CMAIN * es); q.tail_px = (long long)(TC_CTAIL * es);
  q.offset = offset->data; q.mask = mask->data;
  q.f_sn = offset->sn; q.f_sc = offset->sc; q.f_sh = offset->sh; q.f_sw = offset->sw;
  q.m_sn = mask->sn; q.m_sc = mask->sc; q.m_sh = mask->sh; q.m_sw = mask->sw;
  q.gcol = gcol; q.gcol_ld = gcol_ld;
  q.gx = gx_rows; q.gx_ld = gx_ld;
  if (grad_offset) { q.goff = (float*)grad_offset->data; q.gf_sn = grad_offset->sn; q.gf_sc = grad_offset->sc; q.gf_sh = grad_offset->sh; q.gf_sw = grad_offset->sw; }
  if (grad_mask) { q.gmask = (float*)grad_mask->data; q.gm_sn = grad_mask->sn; q.gm_sc = grad_mask->sc; q.gm_sh = grad_mask->sh; q.gm_sw = grad_mask->sw; }
  q.B = (int)x->n; q.H = (int)x->h; q.W = (int)x->w;
  dim3 grid(ceil_div((long long)x->h * x->w, BC_PIX), (unsigned)x->n);
  if (gcol_dtype == VFI_F32) {
    VFI_DISPATCH(offset->dtype, TO, { dcn_bwd_cols_kernel<TO, float, false><<<grid, BC_THREADS, 0, st>>>(q); });
  } else {
    VFI_DISPATCH(offset->dtype, TO, { dcn_bwd_cols_kernel<TO, __nv_bfloat16, false><<<grid, BC_THREADS, 0, st>>>(q); });
  }
  VFI_LAUNCH_CHECK("dcn_bwd_cols_kernel");
  return VFI_OK;
}

// A plane pair as the fused entry points take it: bf16 channels-last views [B,64,H,W] / [B,<=8,H,W] with dense rows and images
// (sh = W * sw, sn = H * sh) and a pixel stride that is a multiple of 16 bytes -- two dense buffers or the channel ranges
// 0..63 / 64..71 of ONE [B,H,W,72] record buffer.
static bool is_plane_view(const vfi_tensor* t, long long c_max, long long B, long long H, long long W) {
  return t && t->data && t->dtype == VFI_BF16 && t->n == B && t->h == H && t->w == W && t->c > 0 && t->c <= c_max && t->sc == 1 &&
         t->sw >= c_max && t->sw % 8 == 0 && t->sh == W * t->sw && (B == 1 || t->sn == H * t->sh) && aligned(t->data, 16);
}

// The fused training form of the data gradients (the backward of vfi_dcn_fwd_fused): x as planes, read where they lie;
// offsets / mask from the raw 27-channel offset_conv output; the gradient goes back to that tensor.
int dcn_tc_bwd_data_cols_fused(const void* gcol, long long gcol_ld, const vfi_tensor* x_main, const vfi_tensor* x_tail,
                               const vfi_tensor* conv27, float* gx_rows, long long gx_ld, const vfi_tensor* grad_conv27,
                               cudaStream_t st) {
  const char* who = "vfi_dcn_bwd_data_cols_fused";
  VFI_REQUIRE(gcol && x_main && x_tail && conv27, VFI_ERR_INVALID, "%s: null pointer", who);
  const long long B = x_main->n, H = x_main->h, W = x_main->w;
  VFI_REQUIRE(is_plane_view(x_main, TC_CMAIN, B, H, W) && x_main->c == TC_CMAIN && is_plane_view(x_tail, TC_CTAIL, B, H, W) &&
                  x_tail->c <= 4, VFI_ERR_UNSUPPORTED,
              "%s: x must be bf16 planes (main [B,64,H,W] + tail [B,<=4,H,W] channels-last views, 16-byte aligned pixels)", who);
  VFI_REQUIRE(conv27->data && conv27->n == B && conv27->c == 27 && conv27->h == H && conv27->w == W &&
                  (conv27->dtype == VFI_BF16 || conv27->dtype == VFI_F16), VFI_ERR_INVALID, "%s: conv27 must be a 16-bit [B,27,H,W] tensor", who);
  VFI_REQUIRE(gcol_ld >= 9 * BC_TAP_LD && gcol_ld % 4 == 0 && aligned(gcol, 16), VFI_ERR_INVALID,
              "%s: gcol rows must hold 9 x %d columns, row length a multiple of 4, 16-byte aligned", who, BC_TAP_LD);
  if (gx_rows) VFI_REQUIRE(gx_ld >= TC_CMAIN + 4 && gx_ld % 4 == 0 && aligned(gx_rows, 16), VFI_ERR_INVALID,
                           "%s: grad_x rows must hold >= %d floats, 16-byte aligned", who, TC_CMAIN + 4);
  if (grad_conv27) VFI_REQUIRE(grad_conv27->data && grad_conv27->dtype == VFI_F32 && same_shape(grad_conv27, conv27), VFI_ERR_INVALID,
                               "%s: grad_conv27 must be f32 [B,27,H,W]", who);
  const long long P = B * H * W;
  if (P == 0 || (!gx_rows && !grad_conv27)) return VFI_OK;
  VFI_REQUIRE(P < 2147483647LL / 2 && B <= 65535, VFI_ERR_UNSUPPORTED, "%s: more than 2^30 pixels or 65535 images per call", who);
  BcParams q{};
  q.x_main = reinterpret_cast<const uint8_t*>(x_main->data); q.x_tail = reinterpret_cast<const uint8_t*>(x_tail->data);
  q.main_px = x_main->sw * 2; q.tail_px = x_tail->sw * 2;
  q.offset = conv27->data; q.mask = conv27->data;
  q.f_sn = conv27->sn; q.f_sc = conv27->sc; q.f_sh = conv27->sh; q.f_sw = conv27->sw;
  q.gcol = gcol; q.gcol_ld = gcol_ld;
  q.gx = gx_rows; q.gx_ld = gx_ld;
  if (grad_conv27) { q.goff = (float*)grad_conv27->data; q.gf_sn = grad_conv27->sn; q.gf_sc = grad_conv27->sc; q.gf_sh = grad_conv27->sh; q.gf_sw = grad_conv27->sw; }
  q.B = (int)B; q.H = (int)H; q.W = (int)W;
  dim3 grid(ceil_div(H * W, BC_PIX), (unsigned)B);
  if (conv27->dtype == VFI_BF16) dcn_bwd_cols_kernel<__nv_bfloat16, __nv_bfloat16, true><<<grid, BC_THREADS, 0, st>>>(q);
  else dcn_bwd_cols_kernel<__half, __nv_bfloat16, true><<<grid, BC_THREADS, 0, st>>>(q);
  VFI_LAUNCH_CHECK("dcn_bwd_cols_kernel<fused27>");
  return VFI_OK;
}

// Returns 1 when a kernel of this library gave up on a pipeline wait since the last call (results of that launch are
// invalid), with info = [block << 32 | warp, barrier shared-memory address, parity, number of waiters that gave up].
int dcn_tc_abort_info(unsigned long long* info) {
  unsigned int flag = 0, zero = 0;
  unsigned long long z[36] = {0};
  VFI_CUDA(cudaDeviceSynchronize());
  VFI_CUDA(cudaMemcpyFromSymbol(&flag, g_abort_flag, sizeof(flag)));
  if (info) VFI_CUDA(cudaMemcpyFromSymbol(info, g_abort_info, sizeof(z)));
  VFI_CUDA(cudaMemcpyToSymbol(g_abort_flag, &zero, sizeof(zero)));
  VFI_CUDA(cudaMemcpyToSymbol(g_abort_info, z, sizeof(z)));
  return flag ? 1 : 0;
}

int umma_ts_selftest(const void* A, const void* Bm, float* D, uint32_t* raw, cudaStream_t st) {
  VFI_REQUIRE(A && Bm && D && raw, VFI_ERR_INVALID, "vfi_selftest_umma_ts: need A [128,64], B [80,64], D [128,80], raw [128,32]");
  const size_t smem = TC_B_BYTES + 64 + 1024;
  VFI_CUDA(cudaFuncSetAttribute(umma_ts_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_ts_selftest_kernel<<<1, 128, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(A),
                                                reinterpret_cast<const __nv_bfloat16*>(Bm), D, raw);
  VFI_LAUNCH_CHECK("umma_ts_selftest_kernel");
  return VFI_OK;
}

}  // namespace vfi
