// dcn_tc6.cuh -- v6 of the tcgen05 DCNv2 forward.  Included by dcn_tc.cu inside namespace vfi::<anonymous>.
//
// What bounds this operator on a B200 is not the tensor pipe and not HBM but the SM's 128 B/clk load/store data path:
// every output pixel needs 9 taps x 4 corners x 134 B = 4.8 KB of activation reads to build its A row (37.7 clk per pixel
// at best, against 12.8 clk of tcgen05.mma), and in v4 the same data path also carried the A-stage stores (1.3 KB / px),
// strided epilogue stores and the tensor core's own A-operand reads from shared memory.  v6 keeps only the compulsory
// gather on that path:
//
//   * the A operand lives in TENSOR MEMORY: producers write their lerp results with tcgen05.st (own write port, 256 B/clk)
//     into an 8-stage ring of 32-column K blocks, and the MMA is issued in the A-from-TMEM form
//     (tcgen05.mma [d], [a_tmem], b_desc): no A stores to shared memory, no A reads from it;
//   * the activation box of a tile (tile + 3x3 reach + 4 px halo = 18 x 26 pixels, 66 KB) is staged in shared memory
//     by bulk copies one tile ahead (double buffered), ZERO PADDED outside the image, so every corner is an LDS.128 at a
//     compile-time offset from the sample's first corner (no clamping, no per-corner validity, no L1 tag / miss traffic);
//     samples that leave the box (|offset| > 4 px at the tile edge) read global memory lane by lane;
//   * the epilogue transposes through a swizzled staging tile so that global stores are full 128 B lines.
//
// Thread <-> data mapping of a main K block (tap t, 64 channels).  tcgen05.st.16x256b gives thread (g = lane / 4,
// u = lane % 4) the 32-bit columns {8i + 2u, 8i + 2u + 1 : i = 0..3} of TMEM lanes g and g + 8, i.e. 16 channels of two
// pixels.  The K order of a block is free as long as the weight image uses the same one, so thread u takes the 16-byte
// chunks u and u + 4 of each corner pixel: the four lanes of a pixel read 64 contiguous bytes, and the two pixels that
// share a quarter-warp LDS phase read opposite halves of their 128-byte rows (odd g loads chunk u + 4 first, then u, and
// swaps the two results back with 8 SELs), which makes every LDS.128 bank-conflict free whatever the offsets are.
//
//   K element e of main block t  ->  channel ((e / 32) * 32) + 8 * ((e % 16) / 4) + 4 * ((e / 16) % 2) + e % 4   (v6_k_to_tap_channel)
//   tail block (kb = 9)          ->  e < 36: tap e / 4, channel 64 + e % 4 (lane = pixel, tcgen05.st.32x32b); rest zero; K = 48

#ifndef V6_ROTATE
#define V6_ROTATE 0          // 1: the K blocks a producer group takes rotate from tile to tile (the tail block moves around)
#endif
#ifndef V6_FIRST_TAPS
#define V6_FIRST_TAPS 4      // taps published by the first geometry barrier
#endif
#ifndef V6_NS_PROD
#define V6_NS_PROD 128       // sleep of the producers' per-tile waits (ns); A/B on one box: 32 -> 4.43, 64 -> 4.37, 128 -> 4.25-4.33, 256 -> 4.30 ms
#endif
#ifndef V6_NS_STAGE
#define V6_NS_STAGE 64       // sleep of the producers' A-stage wait (ns)
#endif
#ifndef V6_PAIRING
#define V6_PAIRING 0         // 0: group g takes K blocks (g, g + 5); 1: (g, 9 - g), i.e. the tail block goes to group 0
#endif
#ifndef V6_HALO
#define V6_HALO 4            // halo of the staged box beyond the 3x3 reach, in pixels
#endif
#ifndef V6_NS_MMA
#define V6_NS_MMA 20         // sleep of the MMA warp's operand wait (ns)
#endif
#ifndef V6_NS_HELP
#define V6_NS_HELP 256       // sleep of the box / geometry / epilogue warps' waits (ns)
#endif
#ifndef V6_NS_LOAD
#define V6_NS_LOAD 64        // sleep of the weight loader's wait (ns)
#endif
constexpr int V6_BOX_H = TC_TH + 2 + 2 * V6_HALO, V6_BOX_W = TC_TW + 2 + 2 * V6_HALO, V6_BOX_PX = V6_BOX_H * V6_BOX_W;   // 18 x 26 = 468 pixels
constexpr int V6_BOX_TOP = 1 + V6_HALO, V6_BOX_LEFT = 1 + V6_HALO;                // box origin = tile origin - (5, 5)
constexpr int V6_MAIN_PX = TC_CMAIN * 2, V6_TAIL_PX = TC_CTAIL * 2;               // bytes per pixel: 128 / 16
constexpr int V6_MAIN_ROW = V6_BOX_W * V6_MAIN_PX, V6_TAIL_ROW = V6_BOX_W * V6_TAIL_PX;   // bytes per box row: 3328 / 416
// Warp roles.  1024 threads x 64 registers is the whole register file; the fifth producer group is worth more than the
// eight registers per thread it costs (the producers are latency bound).  setmaxnreg (72 for producers / 40 for helpers)
// was tried: ptxas fails register allocation in the producer region under a setmaxnreg budget of 72 or even 80.
constexpr int V6_GROUPS = 5;                                                      // producer groups: K blocks g and g + 5 of every tile
constexpr int V6_PRODUCER_WARPS = 4 * V6_GROUPS;                                  // x 4 TMEM sub-partitions = warps 0..19
constexpr int V6_W_MMA = 20, V6_W_BOX = 21, V6_W_BLOAD = 22;                      // warp 23 idles
constexpr int V6_W_EPI = 24, V6_EPI_WARPS = 4;                                    // warps 24..27: epilogue (one per TMEM quarter)
constexpr int V6_W_GEO = 28, V6_GEO_WARPS = 4;                                    // warps 28..31: tap geometry
constexpr int V6_THREADS = 32 * 32;                                               // 1024
constexpr int V6_KBLOCKS = 10;                                                    // 9 main + 1 tail
constexpr int V6_NA = 8, V6_NB = 3;                                               // TMEM A ring / smem B ring depth
// "Block n consumed" barriers, one per A stage.  A parity wait cannot tell "completed k times" from "k - 2 times", so a
// waiter must never be two phases (16 blocks) ahead of its barrier.  It cannot be: consecutive blocks of a producer group are
// 5 apart, the group's previous wait established that block n - 13 was consumed, and the MMA warp consumes in order, so
// block n - 16 is consumed whenever block n is attempted.  (The ring may be made longer than the stage ring -- V6_NDONE_N --
// but measured slower: 8 -> 4.62, 16 -> 5.88, 24 -> 6.19 ms on one box.  The weight-gradient kernel, whose stage ring (3) is
// SHORTER than the group stride, needs the longer ring and deadlocked without it.)
#ifndef V6_NDONE_N
#define V6_NDONE_N 8
#endif
constexpr int V6_NDONE = V6_NDONE_N;
constexpr int V6_A_COL0 = 256;                                                    // TMEM columns [256, 512): A ring
constexpr int V6_TMEM_COLS = 512;
constexpr uint32_t V6_SLOW = 0x80000000u;                                         // entry.x: sample not served by the box
// Box offset of the tile's own first pixel: always inside the image and always copied.  Zero-weight entries (dead samples,
// rows of a partial tile) point here.
constexpr uint32_t V6_SAFE = (V6_BOX_TOP * V6_BOX_W + V6_BOX_LEFT) * V6_MAIN_PX;

__host__ __device__ inline void v6_k_to_tap_channel(int kb, int kk, int& tap, int& c) {
  if (kb < 9) { tap = kb; c = ((kk >> 5) << 5) + 8 * ((kk >> 2) & 3) + 4 * ((kk >> 4) & 1) + (kk & 3); }
  else if (kb == 9 && kk < 36) { tap = kk >> 2; c = TC_CMAIN + (kk & 3); }
  else { tap = 0; c = -1; }
}

struct __align__(1024) V6Smem {
  uint8_t b[V6_NB][TC_B_BYTES];                        // weight K blocks (bulk copies, SWIZZLE_128B image)
  uint8_t box_main[2][V6_BOX_PX * V6_MAIN_PX];         // 2 x 59,904 B
  uint8_t box_tail[2][V6_BOX_PX * V6_TAIL_PX];         // 2 x  7,488 B
  uint4 geo[2][9][TC_M];                               // x: box byte offset | V6_SLOW, y/z: 4 bf16 weights, w: global pixel (slow)
  uint8_t ostage[TC_M * TC_CMAX * 2];                  // epilogue staging tile: planes use 128 B rows (chunk j of row r at j ^ (r & 7)), channels-last rows O x 2 B
  uint16_t raw[27][TC_M];                              // offset / mask values of the next tile (cp.async), [channel][tile row]
  unsigned long long full[V6_NA], done[V6_NDONE];      // per K block n: operands ready (slot n % 8) / MMAs complete (slot n % 16)
  unsigned long long acc_full[2], acc_empty[2], geo_first[2], geo_full[2], geo_empty[2], box_full[2], box_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint4 lds16(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
  return r;
}
__device__ __forceinline__ uint2 lds8(uint32_t saddr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(saddr));
  return r;
}
__device__ __forceinline__ void sts_u16(uint32_t saddr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(v) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// A-from-TMEM form of the MMA: A = 128 lanes x 8 columns (16 bf16 of K per lane) at a_tmem, B from shared memory.
// Called by the whole converged warp (operands warp-uniform); one elected lane issues.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
// 16 TMEM lanes x 32 columns: thread (g, u) supplies columns 8i + 2u, 8i + 2u + 1 of lane g (r[4i], r[4i+1]) and of
// lane g + 8 (r[4i+2], r[4i+3]).
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 32 TMEM lanes x 8 columns: thread = lane, r[j] = column j.
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// a tensor element as raw bits in a 32-bit register (no conversion instruction at load time), and back
template <typename T> __device__ __forceinline__ uint32_t ldg_bits(const T* p);
template <> __device__ __forceinline__ uint32_t ldg_bits<float>(const float* p) { return __float_as_uint(__ldg(p)); }
template <> __device__ __forceinline__ uint32_t ldg_bits<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __ldg(reinterpret_cast<const unsigned short*>(p));
}
template <> __device__ __forceinline__ uint32_t ldg_bits<__half>(const __half* p) { return __ldg(reinterpret_cast<const unsigned short*>(p)); }
template <typename T> __device__ __forceinline__ float bits_to_f32(uint32_t v);
template <> __device__ __forceinline__ float bits_to_f32<float>(uint32_t v) { return __uint_as_float(v); }
template <> __device__ __forceinline__ float bits_to_f32<__nv_bfloat16>(uint32_t v) { return __uint_as_float(v << 16); }
template <> __device__ __forceinline__ float bits_to_f32<__half>(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)v)); }

template <int OFF> __device__ __forceinline__ uint4 lds16o(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr), "n"(OFF));
  return r;
}
template <int OFF> __device__ __forceinline__ uint2 lds8o(uint32_t saddr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(r.x), "=r"(r.y) : "r"(saddr), "n"(OFF));
  return r;
}
__device__ __forceinline__ void geo_bar_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }   // the 4 geometry warps
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps

// mbar_wait that adds the cycles spent waiting to `acc` when the debug buffer is enabled (VFI_DCN_DEBUG=1)
template <int NS, bool DBG>
__device__ __forceinline__ void mbar_wait_d(uint32_t bar, uint32_t parity, long long& acc) {
  if (!DBG) { mbar_wait_ns<NS>(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait_ns<NS>(bar, parity);
  acc += clock64() - t0;
}

// Rare path of the geometry: the sample is not served by the staged box.  Appendix B of SURVEY.md with corners clamped
// into the image and zero weight where torchvision skips a corner; the producers read these from global memory.
__device__ __noinline__ uint4 v6_geo_entry_slow(int H, int W, int base, int y, int x, int k, float dy, float dx, float mk) {
  float py = (float)(y - 1 + k / 3) + dy;
  float px = (float)(x - 1 + k % 3) + dx;
  const bool live = (py > -1.0f) && (py < (float)H) && (px > -1.0f) && (px < (float)W);
  if (!live) return make_uint4(V6_SAFE, 0u, 0u, 0u);
  const float fy = floorf(py), fx = floorf(px);
  const int y0 = (int)fy, x0 = (int)fx;
  const float lh = py - fy, lw = px - fx, hh = 1.0f - lh, hw = 1.0f - lw;
  const bool r0 = (unsigned)y0 < (unsigned)H, r1 = (unsigned)(y0 + 1) < (unsigned)H;
  const bool c0 = (unsigned)x0 < (unsigned)W, c1 = (unsigned)(x0 + 1) < (unsigned)W;
  const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
  const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
  uint2 wq;
  store_geo_w(wq, (r0 && c0) ? hh * hw * mk : 0.0f, (r0 && c1) ? hh * lw * mk : 0.0f, (r1 && c0) ? lh * hw * mk : 0.0f,
              (r1 && c1) ? lh * lw * mk : 0.0f);
  return make_uint4(V6_SLOW, wq.x, wq.y,
                    (uint32_t)(base + cy0 * W + cx0) | ((uint32_t)(cx1 - cx0) << 30) | ((uint32_t)(cy1 - cy0) << 31));
}

// Sampling geometry of one tap: fy0 / fx0 = (float)(y - 1 + i) / (float)(x - 1 + j).  Fast path: no clamping and no corner
// validity -- the box is zero padded outside the image, so an out-of-image corner contributes w * 0 exactly as torchvision's
// skipped corner does, and its "dead sample" rule (py <= -1, py >= H, ...) falls out of the same zeros.
// Returns false (entry invalid) when the box does not serve the sample; the caller then uses v6_geo_entry_slow.
__device__ __forceinline__ bool v6_geo_entry(int by0, int bx0, float fy0, float fx0, float dy, float dx, float mk, uint4& e) {
  const float py = fy0 + dy, px = fx0 + dx;
  const float fy = floorf(py), fx = floorf(px);
  const float lh = py - fy, lw = px - fx, hh = 1.0f - lh, hw = 1.0f - lw;
  const int ry = (int)fy - by0, rx = (int)fx - bx0;
  uint2 wq;
  store_geo_w(wq, hh * hw * mk, hh * lw * mk, lh * hw * mk, lh * lw * mk);
  e = make_uint4((uint32_t)(ry * V6_BOX_W + rx) * V6_MAIN_PX, wq.x, wq.y, 0u);
  return (unsigned)ry <= (unsigned)(V6_BOX_H - 2) && (unsigned)rx <= (unsigned)(V6_BOX_W - 2);
}

// Two 16-byte chunks of the modulated bilinear sample `e`; bF / bS = box base + the lane's first / second chunk offset.
// FAST: the caller has checked that the sample is served by the box (straight-line code, all eight loads in flight).
template <bool FAST>
__device__ __forceinline__ void v6_sample_main(const uint4& e, uint32_t bF, uint32_t bS, const uint8_t* x_main, uint32_t main_row,
                                               uint32_t c_first, uint32_t c_second, uint4& F, uint4& S) {
  uint4 f[4], g[4];
  if (FAST || (int)e.x >= 0) {
    const uint32_t aF = bF + e.x, aS = bS + e.x;
    f[0] = lds16o<0>(aF); f[1] = lds16o<V6_MAIN_PX>(aF); f[2] = lds16o<V6_MAIN_ROW>(aF); f[3] = lds16o<V6_MAIN_ROW + V6_MAIN_PX>(aF);
    g[0] = lds16o<0>(aS); g[1] = lds16o<V6_MAIN_PX>(aS); g[2] = lds16o<V6_MAIN_ROW>(aS); g[3] = lds16o<V6_MAIN_ROW + V6_MAIN_PX>(aS);
  } else {
    // rare: the sample is not served by the staged box -> global memory (clamped corners, validity in the weights)
    const uint8_t* a00 = x_main + (unsigned long long)(e.w & 0x3fffffffu) * V6_MAIN_PX;
    const uint8_t* a01 = a00 + ((e.w & 0x40000000u) ? V6_MAIN_PX : 0);
    const uint32_t dy = (e.w & 0x80000000u) ? main_row : 0u;
    f[0] = __ldg(reinterpret_cast<const uint4*>(a00 + c_first)); f[1] = __ldg(reinterpret_cast<const uint4*>(a01 + c_first));
    f[2] = __ldg(reinterpret_cast<const uint4*>(a00 + dy + c_first)); f[3] = __ldg(reinterpret_cast<const uint4*>(a01 + dy + c_first));
    g[0] = __ldg(reinterpret_cast<const uint4*>(a00 + c_second)); g[1] = __ldg(reinterpret_cast<const uint4*>(a01 + c_second));
    g[2] = __ldg(reinterpret_cast<const uint4*>(a00 + dy + c_second)); g[3] = __ldg(reinterpret_cast<const uint4*>(a01 + dy + c_second));
  }
  const uint2 w = make_uint2(e.y, e.z);
  F = lerp_chunk(f[0], f[1], f[2], f[3], w);
  S = lerp_chunk(g[0], g[1], g[2], g[3], w);
}

// The four tail channels (8 bytes) of the modulated bilinear sample `e`.  `half` = 0 / 8: which of the two mirrored
// halves of the 16-byte record this lane reads (odd lanes take the upper one: the eight-byte reads of a phase then use
// all 32 banks instead of every other pair, measured 6.4 -> ~3 excess wavefronts per pixel with i.i.d. offsets).
__device__ __forceinline__ uint2 v6_sample_tail(const uint4& e, uint32_t box_tail, const uint8_t* x_tail, uint32_t tail_row,
                                                uint32_t half) {
  uint2 v[4];
  if ((int)e.x >= 0) {
    const uint32_t a = box_tail + (e.x >> 3) + half;               // 16 B per pixel instead of 128
    v[0] = lds8o<0>(a); v[1] = lds8o<V6_TAIL_PX>(a); v[2] = lds8o<V6_TAIL_ROW>(a); v[3] = lds8o<V6_TAIL_ROW + V6_TAIL_PX>(a);
  } else {
    const uint8_t* a00 = x_tail + (unsigned long long)(e.w & 0x3fffffffu) * V6_TAIL_PX;
    const uint8_t* a01 = a00 + ((e.w & 0x40000000u) ? V6_TAIL_PX : 0);
    const uint32_t dy = (e.w & 0x80000000u) ? tail_row : 0u;
    v[0] = __ldg(reinterpret_cast<const uint2*>(a00)); v[1] = __ldg(reinterpret_cast<const uint2*>(a01));
    v[2] = __ldg(reinterpret_cast<const uint2*>(a00 + dy)); v[3] = __ldg(reinterpret_cast<const uint2*>(a01 + dy));
  }
  const uint4 r = lerp_chunk(make_uint4(v[0].x, v[0].y, 0u, 0u), make_uint4(v[1].x, v[1].y, 0u, 0u),
                             make_uint4(v[2].x, v[2].y, 0u, 0u), make_uint4(v[3].x, v[3].y, 0u, 0u), make_uint2(e.y, e.z));
  return make_uint2(r.x, r.y);
}

// Role: source-box copies (one warp, one lane per box row).  Shared by the forward and the weight-gradient kernel; `S` has
// box_main / box_tail / box_full / box_empty.
template <bool DBG, class S>
__device__ __forceinline__ void v6_role_box(S& s, const TcParams& p, int lane, int my_tiles, int tile0, int tile_step, long long& w0) {
  for (int it = 0; it < my_tiles; ++it) {
    const int sb = it & 1;
    mbar_wait_d<V6_NS_HELP, DBG>(smem_u32(&s.box_empty[sb]), ((uint32_t)(it >> 1) & 1u) ^ 1u, w0);   // producers are done with the old box
    int b, ty0, tx0;
    tile_origin(p, tile0 + it * tile_step, b, ty0, tx0);
    const int by0 = ty0 - V6_BOX_TOP, bx0 = tx0 - V6_BOX_LEFT;
    const int ya = max(by0, 0), yb = min(by0 + V6_BOX_H, p.H), xa = max(bx0, 0), xb = min(bx0 + V6_BOX_W, p.W);
    const uint32_t ncol = (uint32_t)(xb - xa), nrow = (uint32_t)(yb - ya);
    const uint32_t bar = smem_u32(&s.box_full[sb]);
    const uint32_t dst_main = smem_u32(&s.box_main[sb][0]), dst_tail = smem_u32(&s.box_tail[sb][0]);
    if (nrow * ncol != (uint32_t)V6_BOX_PX) {
      // border tile: the part of the box outside the image is zero padding (what torchvision's skipped corners amount to)
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      for (int i = lane; i < V6_BOX_PX; i += 32) {
        const int yy = by0 + i / V6_BOX_W, xx = bx0 + i % V6_BOX_W;
        if (yy < ya || yy >= yb || xx < xa || xx >= xb) {
#pragma unroll
          for (int c = 0; c < 8; ++c) sts16(dst_main + (uint32_t)i * V6_MAIN_PX + c * 16, z);
          sts16(dst_tail + (uint32_t)i * V6_TAIL_PX, z);
        }
      }
    }
    __syncwarp();                                      // the zero padding is ordered before the arrive below (release)
    if (lane == 0) mbar_arrive_expect_tx(bar, nrow * ncol * (V6_MAIN_PX + V6_TAIL_PX));
    __syncwarp();
    const int y = by0 + lane;
    if (lane < V6_BOX_H && y >= ya && y < yb) {
      const size_t gpix = (size_t)(b * p.H + y) * p.W + xa;
      const uint32_t bpix = (uint32_t)(lane * V6_BOX_W + (xa - bx0));
      bulk_g2s(dst_main + bpix * V6_MAIN_PX, p.x_main + gpix * V6_MAIN_PX, ncol * V6_MAIN_PX, bar);
      bulk_g2s(dst_tail + bpix * V6_TAIL_PX, p.x_tail + gpix * V6_TAIL_PX, ncol * V6_TAIL_PX, bar);
    }
    __syncwarp();
  }
}

// Role: tap geometry (four warps; `row` = tile row of this thread).  Shared by the forward and the weight-gradient kernel;
// `S` has raw / geo / geo_first / geo_full / geo_empty.
template <typename TO, bool FUSED27, bool DBG, class S>
__device__ __forceinline__ void v6_role_geometry(S& s, const TcParams& p, int row, int lane, int my_tiles, int tile0, int tile_step,
                                                 long long& w0, long long& w2) {
  constexpr bool dbg = DBG;
  // =========================================================================== tap geometry (4 warps)
  // Thread = tile row, nine taps.  Runs one tile ahead of the producers (double-buffered entries); the offset / mask
  // values arrive in shared memory by cp.async a further tile ahead, so no DRAM round trip sits in this warp.
  // Offsets and masks of a tile: 27 channels x 8 tile rows x 16 pixels = 432 chunks of 16 bytes of the NCHW tensor,
  // brought into raw[channel][tile row * 16 + x] by cp.async (four per thread, no registers held) one tile ahead.
  // Channels-last offset_conv output (p.geo_cl, fused form only): a tile row is one contiguous run of 16 x 27 values, copied
  // as it lies -- raw then holds [tile pixel][27 channels] and the taps below index it accordingly.  Same 432 chunks, but
  // consecutive lanes copy consecutive 16 bytes (4 wavefronts per warp instruction instead of 32).
  uint16_t* const raw_flat = &s.raw[0][0];
  auto fetch_raw = [&](int it) {
    int b, ty0, tx0;
    tile_origin(p, tile0 + it * tile_step, b, ty0, tx0);
    const int rows = min(TC_TH, p.H - ty0), cols = min(TC_TW, p.W - tx0);
    if (FUSED27 && p.geo_cl) {
      const int per_row = cols * 27 / 8;                  // 16-byte chunks of one tile row (W % 8 == 0)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = row + 128 * j, r = i / 54, jj = i - r * 54;
        if (i < 8 * 54 && r < rows && jj < per_row) {
          const TO* src = reinterpret_cast<const TO*>(p.offset) + b * p.f_sn + ((long long)(ty0 + r) * p.W + tx0) * 27 + jj * 8;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(raw_flat + r * (TC_TW * 27) + jj * 8)), "l"(src) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = row + 128 * j, c = i >> 4, r = (i >> 1) & 7, x8 = (i & 1) * 8;
      if (i < 27 * 16 && r < rows && x8 < cols) {
        const TO* src;
        if (FUSED27) src = reinterpret_cast<const TO*>(p.offset) + b * p.f_sn + (c < 18 ? (c < 9 ? c : c + 9) : c - 9) * p.f_sc;
        else if (c < 18) src = reinterpret_cast<const TO*>(p.offset) + b * p.f_sn + c * p.f_sc;
        else src = reinterpret_cast<const TO*>(p.mask) + b * p.m_sn + (c - 18) * p.m_sc;
        src += (long long)(ty0 + r) * (c < 18 || FUSED27 ? p.f_sh : p.m_sh) + tx0 + x8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&s.raw[c][r * TC_TW + x8])), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // kernel channel c (0..17: offsets dy/dx interleaved, 18..26: mask) of this thread's pixel
  const bool geo_cl = FUSED27 && p.geo_cl;
  auto rd = [&](int c) -> uint32_t {
    if (geo_cl) return raw_flat[row * 27 + (c < 18 ? (c < 9 ? c : c + 9) : c - 9)];
    return s.raw[c][row];
  };
  if (my_tiles > 0) fetch_raw(0);
  for (int it = 0; it < my_tiles; ++it) {
    const int gb = it & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    geo_bar_sync();                                      // every thread's chunks of this tile have landed
    mbar_wait_d<V6_NS_HELP, DBG>(smem_u32(&s.geo_empty[gb]), ((uint32_t)(it >> 1) & 1u) ^ 1u, w0);   // producers are done with the old contents
    const long long tg0 = dbg ? clock64() : 0;
    int b, ty0, tx0;
    tile_origin(p, tile0 + it * tile_step, b, ty0, tx0);
    const int y = ty0 + row / TC_TW, x = tx0 + row % TC_TW;
    const int by0 = ty0 - V6_BOX_TOP, bx0 = tx0 - V6_BOX_LEFT;
    if (y < p.H && x < p.W) {
      const int base = b * p.H * p.W;
      const float fy0 = (float)(y - 1), fx0 = (float)(x - 1);
      // taps 0..3 first: they are all the first K blocks of the tile need, so the producers start on them while
      // taps 4..8 are still being computed (the geometry of a tile cannot start before the previous tile but one is done)
      auto taps = [&](auto k0_tag, auto k1_tag) {
        constexpr int K0 = decltype(k0_tag)::value, K1 = decltype(k1_tag)::value;
        uint32_t slow = 0;                               // taps the box does not serve (rare): patched below
        uint32_t raw[27];                                // [dy x9 | dx x9 | mask x9] as raw bits; only [K0, K1) is live
#pragma unroll
        for (int k = K0; k < K1; ++k) {                  // tap k: channels 2k, 2k+1 of the offsets, channel k of the mask
          raw[k] = rd(2 * k);
          raw[9 + k] = rd(2 * k + 1);
          raw[18 + k] = rd(18 + k);
        }
#pragma unroll
        for (int k = K0; k < K1; ++k) {                  // straight-line code: the taps interleave
          float mk = bits_to_f32<TO>(raw[18 + k]);
          // the sigmoid result is rounded to the tensor dtype, as torch.sigmoid on that tensor would
          if (FUSED27) mk = to_f32<TO>(from_f32<TO>(__fdividef(1.0f, 1.0f + __expf(-mk))));
          uint4 e;
          if (!v6_geo_entry(by0, bx0, fy0 + (float)(k / 3), fx0 + (float)(k % 3), bits_to_f32<TO>(raw[k]), bits_to_f32<TO>(raw[9 + k]), mk, e))
            slow |= 1u << k;
          s.geo[gb][k][row] = e;
        }
        if (slow) {
#pragma unroll
          for (int k = K0; k < K1; ++k) {                // static indices keep `raw` in registers
            if (!(slow & (1u << k))) continue;
            float mk = bits_to_f32<TO>(raw[18 + k]);
            if (FUSED27) mk = to_f32<TO>(from_f32<TO>(__fdividef(1.0f, 1.0f + __expf(-mk))));
            s.geo[gb][k][row] = v6_geo_entry_slow(p.H, p.W, base, y, x, k, bits_to_f32<TO>(raw[k]), bits_to_f32<TO>(raw[9 + k]), mk);
          }
        }
      };
      taps(std::integral_constant<int, 0>{}, std::integral_constant<int, V6_FIRST_TAPS>{});
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s.geo_first[gb]));
      taps(std::integral_constant<int, V6_FIRST_TAPS>{}, std::integral_constant<int, 9>{});
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) s.geo[gb][k][row] = make_uint4(V6_SAFE, 0u, 0u, 0u);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s.geo_first[gb]));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&s.geo_full[gb]));
    geo_bar_sync();                                      // everybody has read its values: the raw buffer is free
    if (it + 1 < my_tiles) fetch_raw(it + 1);            // lands while the producers work through this tile
    if (dbg) w2 += clock64() - tg0;
  }
}

// TO: 16-bit dtype of the offset / mask tensors; TOUT: output dtype; FUSED27: offsets and mask come from the 27-channel
// offset_conv output; PLANES: output as bf16 planes (else any strided tensor); DBG: per-role cycle counters.
template <typename TO, typename TOUT, bool FUSED27, bool PLANES, bool DBG>
__global__ void __launch_bounds__(V6_THREADS, 1) dcn_tc6_fwd_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  V6Smem& s = *reinterpret_cast<V6Smem*>(smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    // One pair of barriers per K block n (slot n % 8), so that the MMA warp does ONE wait and ONE commit per block (every
    // tcgen05 / mbarrier instruction on that warp costs ~100 cycles of issue latency, and its blocks are only 160 tensor
    // cycles long): full = the A stage (4 producer warps) and the weight block (loader's expect_tx + bytes) are ready;
    // done = the block's MMAs have completed -> A stage n % 8 and B stage n % 3 are free.  The loader reads done[n - 3]
    // before block n - 3 + 8 can complete (that block needs a weight block the loader has not issued yet), so the
    // two consumers of `done` never see the barrier two phases ahead.
    for (int i = 0; i < V6_NA; ++i) mbar_init(smem_u32(&s.full[i]), 5);
    for (int i = 0; i < V6_NDONE; ++i) mbar_init(smem_u32(&s.done[i]), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s.acc_full[i]), 1);                   // one tcgen05.commit
      mbar_init(smem_u32(&s.acc_empty[i]), V6_EPI_WARPS);
      mbar_init(smem_u32(&s.geo_first[i]), V6_GEO_WARPS);       // taps 0..3 written (what the first K blocks of a tile need)
      mbar_init(smem_u32(&s.geo_full[i]), V6_GEO_WARPS);        // all nine taps written
      mbar_init(smem_u32(&s.geo_empty[i]), V6_PRODUCER_WARPS);
      mbar_init(smem_u32(&s.box_full[i]), 1);                   // the copy warp's expect_tx arrival (+ the bytes)
      mbar_init(smem_u32(&s.box_empty[i]), V6_PRODUCER_WARPS);
    }

    fence_barrier_init();
  }
  if (warp == V6_W_MMA) tmem_alloc(smem_u32(&s.tmem_base), V6_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;
  // Tiles are dealt round-robin (neighbouring tiles run at the same time on neighbouring SMs: shared halo lines are
  // fetched from DRAM once and found in L2 by the neighbour).
  const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile0 = (int)blockIdx.x, tile_step = (int)gridDim.x;
  constexpr bool dbg = DBG;
  long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0;      // cycles spent in this role's waits (debug only)
  const long long t_begin = clock64();

  if (warp < V6_PRODUCER_WARPS) {
    // =========================================================================== A-operand producers
    // Group gi (4 warps, one per TMEM sub-partition) produces K blocks gi and gi + 5 of every tile into A-ring stage n % 8
    // (n = tile_iter * 10 + kb); warp q of a group owns tile rows (= TMEM lanes) [32q, 32q + 32).
    const int group = warp >> 2, q = warp & 3;
    const int g = lane >> 2, u = lane & 3;
    const bool par = (g & 1) != 0;
    const uint32_t c_first = (uint32_t)(u + (par ? 4 : 0)) * 16, c_second = (uint32_t)(u + (par ? 0 : 4)) * 16;
    const uint32_t main_row = V6_MAIN_PX * (uint32_t)p.W, tail_row = V6_TAIL_PX * (uint32_t)p.W;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    for (int it = 0; it < my_tiles; ++it) {
      const int gb = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait_d<V6_NS_PROD, DBG>(smem_u32(&s.geo_first[gb]), tphase, w0); // this tile's geometry, first taps, has been written
      mbar_wait_d<V6_NS_PROD, DBG>(smem_u32(&s.box_full[gb]), tphase, w1);  // this tile's source box has landed in shared memory
      bool all_taps = false;
      const uint32_t box_main = smem_u32(&s.box_main[gb][0]), box_tail = smem_u32(&s.box_tail[gb][0]);
      const uint32_t bF = box_main + c_first, bS = box_main + c_second;
      const int n0 = it * V6_KBLOCKS;
      const int kb_first = V6_ROTATE ? (group + it) % V6_GROUPS : group;
      for (int kb = kb_first; kb < V6_KBLOCKS; kb = V6_PAIRING ? (kb < V6_GROUPS ? V6_KBLOCKS - 1 - kb : V6_KBLOCKS) : kb + V6_GROUPS) {   // 10 K blocks, 5 groups: two each
        const int n = n0 + kb, sa = n % V6_NA;
        // stage sa was last used by block n - 8: wait until that block has been consumed (nothing to wait for when n < 8)
        const int m8 = n - V6_NA;
        const uint32_t empty_bar = smem_u32(&s.done[(m8 + V6_NDONE) % V6_NDONE]);
        const uint32_t empty_par = m8 < 0 ? 1u : (uint32_t)(m8 / V6_NDONE) & 1u;
        const uint32_t a_taddr = tmem_base + lane_base + (uint32_t)(V6_A_COL0 + sa * 32);
        if (kb >= V6_FIRST_TAPS && !all_taps) {                  // the later taps (and the tail block, which reads all nine)
          mbar_wait_d<V6_NS_PROD, DBG>(smem_u32(&s.geo_full[gb]), tphase, w0);
          all_taps = true;
        }
        if (kb < 9) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint4 e0 = s.geo[gb][kb][q * 32 + h * 16 + g], e1 = s.geo[gb][kb][q * 32 + h * 16 + g + 8];
            uint4 F0, S0, F1, S1;
            if ((int)(e0.x | e1.x) >= 0) {                  // both samples served by the box: one basic block
              v6_sample_main<true>(e0, bF, bS, p.x_main, main_row, c_first, c_second, F0, S0);
              v6_sample_main<true>(e1, bF, bS, p.x_main, main_row, c_first, c_second, F1, S1);
            } else {
              v6_sample_main<false>(e0, bF, bS, p.x_main, main_row, c_first, c_second, F0, S0);
              v6_sample_main<false>(e1, bF, bS, p.x_main, main_row, c_first, c_second, F1, S1);
            }
            // chunk u (X) and chunk u + 4 (Y) of both pixels: odd g loaded them in the opposite order
            const uint4 X0 = par ? S0 : F0, Y0 = par ? F0 : S0, X1 = par ? S1 : F1, Y1 = par ? F1 : S1;
            const uint32_t r[16] = {X0.x, X0.y, X1.x, X1.y, X0.z, X0.w, X1.z, X1.w,
                                    Y0.x, Y0.y, Y1.x, Y1.y, Y0.z, Y0.w, Y1.z, Y1.w};
            if (h == 0) {                                        // the ring stage is needed only now, after the gathers
              mbar_wait_d<V6_NS_STAGE, DBG>(empty_bar, empty_par, w2);
              tc_fence_after();
            }
            tmem_st_16x256b_x4(a_taddr + ((uint32_t)(h * 16) << 16), r);
          }
        } else {
          // ---- tail block: lane = tile row, four channels of each of the nine taps (K = 36 of 48, then the bias slots)
          const int row = q * 32 + lane;
          uint32_t r[24];
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const uint2 v = v6_sample_tail(s.geo[gb][k][row], box_tail, p.x_tail, tail_row, (uint32_t)(lane & 1) * 8u);
            r[2 * k] = v.x; r[2 * k + 1] = v.y;
          }
          r[18] = 0x3f803f80u;                               // K elements 36, 37 = 1.0: the weight image holds bias hi / lo there
#pragma unroll
          for (int i = 19; i < 24; ++i) r[i] = 0u;
          mbar_wait_d<V6_NS_STAGE, DBG>(empty_bar, empty_par, w2);
          tc_fence_after();
          tmem_st_32x32b_x8(a_taddr, r);
          tmem_st_32x32b_x8(a_taddr + 8, r + 8);
          tmem_st_32x32b_x8(a_taddr + 16, r + 16);
        }
        tmem_st_wait();                                          // TMEM writes complete ...
        tc_fence_before();                                       // ... and ordered before the arrive the MMA lane waits on
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s.full[sa]));
      }
      __syncwarp();
      if (lane == 0) {                                           // this warp no longer reads geometry / box buffer gb
        mbar_arrive(smem_u32(&s.geo_empty[gb]));
        mbar_arrive(smem_u32(&s.box_empty[gb]));
      }
    }
  } else if (warp == V6_W_MMA) {
    // =========================================================================== MMA issuer
    // The whole warp runs the loop so that stage indices, TMEM addresses and descriptors are warp-uniform (uniform
    // datapath, no R2UR chain in front of every UTCHMMA); lane 0 issues.  Measured before: ~130 cycles of dependent
    // scalar code per tcgen05.mma made this warp the pipeline's bottleneck at ~11k cycles per tile.
    constexpr uint32_t idesc = umma_idesc_bf16(TC_M, TC_N);
    const uint32_t b_smem = smem_u32(&s.b[0][0]);
    int n = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t acc = (uint32_t)it & 1u, acc_phase = ((uint32_t)it >> 1) & 1u;
      mbar_wait_d<32, DBG>(smem_u32(&s.acc_empty[acc]), acc_phase ^ 1, w0);   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * TC_ACC_STRIDE;
#pragma unroll 1
      for (int kb = 0; kb < V6_KBLOCKS; ++kb, ++n) {
        const int sa = n % V6_NA, sb = n % V6_NB;
        mbar_wait_d<V6_NS_MMA, DBG>(smem_u32(&s.full[sa]), (uint32_t)(n / V6_NA) & 1u, w2);
        tc_fence_after();
        const uint32_t a_tmem = tmem_base + (uint32_t)(V6_A_COL0 + sa * 32);
        const uint64_t bdesc = umma_desc_sw128(b_smem + (uint32_t)sb * TC_B_BYTES);
        const long long ti0 = dbg ? clock64() : 0;
        // 16 bf16 of K = 8 TMEM columns of A = 32 B of the B swizzle atom (+2 in the encoded start address)
        if (!(DBG && (p.experiment & 2))) {                  // diagnostics: bit 1 = issue no MMA
          umma_bf16_ts(d_tmem, a_tmem, bdesc, idesc, kb != 0);
          umma_bf16_ts(d_tmem, a_tmem + 8, bdesc + 2, idesc, 1);
          umma_bf16_ts(d_tmem, a_tmem + 16, bdesc + 4, idesc, 1);
          if (kb != V6_KBLOCKS - 1) umma_bf16_ts(d_tmem, a_tmem + 24, bdesc + 6, idesc, 1);
        }
        const long long ti1 = dbg ? clock64() : 0;
        umma_commit_elect(smem_u32(&s.done[n % V6_NDONE]));
        if (kb == V6_KBLOCKS - 1) umma_commit_elect(smem_u32(&s.acc_full[acc]));
        if (dbg) { w3 += ti1 - ti0; w4 += clock64() - ti1; }
      }
    }
    __syncwarp();
  } else if (warp == V6_W_BLOAD) {
    // =========================================================================== weight-block loader (one lane)
    if (lane == 0) {
      const int total = my_tiles * V6_KBLOCKS;
      int kb = 0;
      for (int n = 0; n < total; ++n) {
        const int sb = n % V6_NB;
        if (n >= V6_NB) {                                // block n - 3 (the previous user of stage sb) has completed
          const int m = n - V6_NB;
          mbar_wait_d<V6_NS_LOAD, DBG>(smem_u32(&s.done[m % V6_NDONE]), (uint32_t)(m / V6_NDONE) & 1u, w0);
        }
        const uint32_t bar = smem_u32(&s.full[n % V6_NA]);
        if (DBG && (p.experiment & 1) && n >= V6_NB) { mbar_arrive(bar); if (++kb == V6_KBLOCKS) kb = 0; continue; }   // diagnostics: no refill
        mbar_arrive_expect_tx(bar, TC_B_BYTES);
        bulk_g2s(smem_u32(&s.b[sb][0]), p.wpacked + (size_t)kb * TC_B_BYTES, TC_B_BYTES, bar);
        if (++kb == V6_KBLOCKS) kb = 0;
      }
    }
    __syncwarp();
  } else if (warp == V6_W_BOX) {
    // =========================================================================== source-box copies (one lane per box row)
    v6_role_box<DBG>(s, p, lane, my_tiles, tile0, tile_step, w0);
  } else if (warp >= V6_W_GEO) {
    // =========================================================================== tap geometry (4 warps)
    v6_role_geometry<TO, FUSED27, DBG>(s, p, (warp - V6_W_GEO) * 32 + lane, lane, my_tiles, tile0, tile_step, w0, w2);
  } else if (warp >= V6_W_EPI) {
    // =========================================================================== epilogue (4 warps)
    const int quad = warp & 3;                           // TMEM lanes [32*quad, 32*quad + 32) belong to this warp
    const int row = quad * 32 + lane;                    // tile row = TMEM lane of this thread
    const int etid = (warp - V6_W_EPI) * 32 + lane;      // 0..127 for the cooperative store
    const uint32_t ostage = smem_u32(&s.ostage[0]);
    // cooperative store: thread = (tile column etid / 8, 16-byte chunk etid % 8); pass i = tile row i
    const int cs_x = etid >> 3, cs_c = etid & 7;
    const uint32_t cs_src = ostage + (uint32_t)cs_x * 128 + ((uint32_t)(cs_c ^ (cs_x & 7)) << 4);   // + 2048 per tile row (16 % 8 == 0)
    const size_t cs_row = (size_t)p.W * V6_MAIN_PX;
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t acc = (uint32_t)it & 1u, acc_phase = ((uint32_t)it >> 1) & 1u;
      int b, ty0, tx0;
      tile_origin(p, tile0 + it * tile_step, b, ty0, tx0);
      mbar_wait_d<V6_NS_HELP, DBG>(smem_u32(&s.acc_full[acc]), acc_phase, w1);
      const long long te0 = dbg ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TC_ACC_STRIDE;
      const int y = ty0 + row / TC_TW, x = tx0 + row % TC_TW;
      const bool inside = y < p.H && x < p.W;
      const size_t pixel = (size_t)(b * p.H + y) * p.W + x;
      __nv_bfloat16* ot = reinterpret_cast<__nv_bfloat16*>(p.out_tail) + pixel * TC_CTAIL;
      TOUT* os = reinterpret_cast<TOUT*>(p.out) + b * p.o_sn + y * p.o_sh + x * p.o_sw;
#pragma unroll
      for (int c16 = 0; c16 < TC_N / 16; ++c16) {
        uint32_t d[16];
        tmem_ld16(taddr + c16 * 16, d);
        tmem_ld_wait();
        if (c16 == TC_N / 16 - 1) {                      // last TMEM read of this accumulator: hand it back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&s.acc_empty[acc]));
        }
        // the bias is already in the accumulator (K elements 36/37 of the tail block)
        if (PLANES) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c0 = c16 * 16 + h * 8;
            if (c0 >= TC_CMAX) break;                    // columns 72..79 are padding of the UMMA N dimension
            uint4 w4;
            uint32_t* w = reinterpret_cast<uint32_t*>(&w4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat162 hv = __floats2bfloat162_rn(__uint_as_float(d[h * 8 + 2 * i]), __uint_as_float(d[h * 8 + 2 * i + 1]));
              w[i] = *reinterpret_cast<uint32_t*>(&hv);
            }
            if (c0 < TC_CMAIN) sts16(ostage + (uint32_t)row * 128 + ((uint32_t)((c0 >> 3) ^ (row & 7)) << 4), w4);
            else if (inside) {
              if (p.O <= TC_CMAIN + 4) { w4.z = w4.x; w4.w = w4.y; }   // tail of <= 4 channels: upper half mirrors the lower
              *reinterpret_cast<uint4*>(ot) = w4;                      // 16 B records of neighbouring pixels coalesce
            }
          }
        } else if (sizeof(TOUT) == 2 && p.out_rows) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {                  // packed rows of O elements: a tile row is one contiguous run
            const int c = c16 * 16 + i;
            const TOUT v = from_f32<TOUT>(__uint_as_float(d[i]));
            if (c < p.O) sts_u16(ostage + (uint32_t)(row * p.O + c) * 2, *reinterpret_cast<const uint16_t*>(&v));
          }
        } else if (inside) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c16 * 16 + i;
            if (c < p.O) os[c * p.o_sc] = from_f32<TOUT>(__uint_as_float(d[i]));
          }
        }
      }
      if (!PLANES && sizeof(TOUT) == 2 && p.out_rows) {
        epi_bar_sync();
        const int cols = min(TC_TW, p.W - tx0);
        const int seg16 = cols * p.O / 8;                  // 16-byte units per tile row (W % 8 == 0)
        const uint32_t pitch = (uint32_t)(TC_TW * p.O * 2);
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + ((size_t)b * p.o_sn + ((size_t)ty0 * p.W + tx0) * p.O) * 2;
        for (int i = 0; i < TC_TH && ty0 + i < p.H; ++i)
          for (int j = etid; j < seg16; j += 128)
            *reinterpret_cast<uint4*>(dst + (size_t)i * p.W * p.O * 2 + 16 * j) = lds16(ostage + i * pitch + 16 * j);
        epi_bar_sync();
      }
      if (PLANES) {
        // main plane: the staged tile leaves as full 128-byte lines (8 lanes per pixel, 4 pixels per warp store)
        epi_bar_sync();
        const int xx = tx0 + cs_x;
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + ((size_t)(b * p.H + ty0) * p.W + xx) * V6_MAIN_PX + cs_c * 16;
        if (xx < p.W) {
#pragma unroll
          for (int i = 0; i < TC_TH; ++i)
            if (ty0 + i < p.H) *reinterpret_cast<uint4*>(dst + i * cs_row) = lds16(cs_src + i * (TC_TW * 128));
        }
        epi_bar_sync();                                   // the staging tile may be overwritten by the next tile
      }
      if (dbg) w3 += clock64() - te0;
    }
  }

  if (dbg && lane == 0) {
    unsigned long long* d = p.debug + ((size_t)blockIdx.x * 32 + warp) * 8;
    d[0] = (unsigned long long)(clock64() - t_begin); d[1] = w0; d[2] = w1; d[3] = w2; d[4] = w3; d[5] = my_tiles; d[6] = w4;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == V6_W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, V6_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ v6 plumbing self test
// D[128, 80] = A[128, 64] * Bm[80, 64]^T through exactly the pieces v6 adds: A written to TMEM by tcgen05.st.16x256b with
// the thread <-> (lane, column) mapping the producers assume, the A-from-TMEM MMA form, and tcgen05.st.32x32b (second
// pass, accumulated: D = 2 A B^T).  `raw` receives the TMEM image of A read back with tcgen05.ld.32x32b ([128][32] u32).
__global__ void __launch_bounds__(128, 1) umma_ts_selftest_kernel(const __nv_bfloat16* __restrict__ A,
                                                                   const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D,
                                                                   uint32_t* __restrict__ raw) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint8_t* sb = base;                       // 10240
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(base + TC_B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(smem_u32(bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
  for (int i = tid; i < TC_N * 8; i += 128) {
    int r = i >> 3, jj = i & 7;
    uint4 v = *reinterpret_cast<const uint4*>(Bm + (size_t)r * 64 + jj * 8);
    *reinterpret_cast<uint4*>(sb + r * 128 + ((jj ^ (r & 7)) << 4)) = v;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_col = 128;                // A at columns [128, 160), second copy (32x32b) at [160, 192), D at [0, 80)
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t* A32 = reinterpret_cast<const uint32_t*>(A);      // A32[row * 32 + col] = K elements 2 col, 2 col + 1
  {
    const int g = lane >> 2, u = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r0 = warp * 32 + h * 16 + g, r1 = r0 + 8;
      uint32_t r[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        r[4 * i + 0] = A32[r0 * 32 + 8 * i + 2 * u]; r[4 * i + 1] = A32[r0 * 32 + 8 * i + 2 * u + 1];
        r[4 * i + 2] = A32[r1 * 32 + 8 * i + 2 * u]; r[4 * i + 3] = A32[r1 * 32 + 8 * i + 2 * u + 1];
      }
      tmem_st_16x256b_x4(tmem_base + lane_base + ((uint32_t)(h * 16) << 16) + a_col, r);
    }
    uint32_t r2[32];
    const int row = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) r2[j] = A32[row * 32 + j];
#pragma unroll
    for (int j = 0; j < 4; ++j) tmem_st_32x32b_x8(tmem_base + lane_base + a_col + 32 + 8 * j, r2 + 8 * j);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {                           // whole warp: the MMA / commit wrappers elect their own lane
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(TC_M, TC_N);
    const uint64_t bdesc = umma_desc_sw128(smem_u32(sb));
    for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_base, tmem_base + a_col + 8 * k, bdesc + 2 * k, idesc, k != 0);
    for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_base, tmem_base + a_col + 32 + 8 * k, bdesc + 2 * k, idesc, 1);
    umma_commit_elect(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  {
    uint32_t d[TC_N];
#pragma unroll
    for (int c = 0; c < TC_N / 16; ++c) tmem_ld16(tmem_base + lane_base + c * 16, d + c * 16);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < TC_N; ++c) D[row * TC_N + c] = __uint_as_float(d[c]);
  }
  {
    uint32_t a[32];
    tmem_ld32(tmem_base + lane_base + a_col, a);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c) raw[row * 32 + c] = a[c];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}
