// dcn_tc7.cuh -- v7 of the tcgen05 DCNv2 forward.  Included by dcn_tc.cu after dcn_tc6.cuh, inside namespace vfi::<anonymous>.
//
// Same arithmetic as v6 (bit-identical results: same K order, same packed-bf16 blend, same accumulation order), rebuilt
// around what round 2 measured (profiles/r02_a_*):
//
//   * scripts/microbench/gather_floor.cu: the producers' work alone -- entry reads, 36 x LDS.128 per pixel, blend,
//     tcgen05.st -- runs at 38.5 clk per pixel (2.2 ms per cfg2 layer) with 16 warps or with 28.  v6 needs 74.5 because its
//     shared-memory data pipe is 93 % busy on 64 wavefront-cycles per pixel of which 26 are not the compulsory gather, its
//     four geometry warps are busy 88 % of the time (the producers wait on them 16 %), and its MMA warp spends ~650 clk per
//     K block on issue (it would bound the kernel as soon as the producers ran at their own speed).
//
// What changed:
//   * TMA tensor maps (cp.async.bulk.tensor) bring in the source box (main + tail planes: ONE instruction each per tile,
//     out-of-image pixels zero-filled by the copy engine -- no border code, any pixel stride, e.g. a single [B,H,W,72]
//     activation buffer) and the 27 offset / mask channels of the tile (one 3-D box instead of 432 cp.async with 32
//     different lines per warp instruction); the main plane leaves through a TMA store of the swizzled staging tile.
//   * The three tail channels (K = 27 of 603) no longer occupy a producer K block: the geometry warps, which hold the
//     sampling positions in registers anyway, gather them and write their A block straight into two reserved TMEM column
//     ranges.  The producers run nine identical main blocks per tile, dealt round-robin over FOUR groups (16 warps: the
//     microbenchmark shows 16 warps saturate the data pipe), and nobody re-reads geometry entries for the tail.
//   * EIGHT geometry warps (two threads per pixel, taps 0-3 + 8 / 4-7) halve the geometry latency per tile.
//   * Warp roles are dispatched on a warp-uniform index (shuffle broadcast) so the MMA loop lives on the uniform data path,
//     with incremental ring counters instead of divisions.
//
// TMEM columns: [0,80) / [128,208) accumulators of even / odd tiles, [80,128) and [208,256) four tail A buffers of 24 columns,
// [256,512) the main A ring (8 stages x 32 columns).

#ifndef V7_PARTS
#define V7_PARTS 3           // geometry threads per pixel: 3 -> twelve geometry warps (three taps each) + three producer groups;
#endif                       //                             2 -> eight geometry warps (taps 0-3 + 8 / 4-7) + four producer groups
static_assert(V7_PARTS == 2 || V7_PARTS == 3, "V7_PARTS");
constexpr int V7_GROUPS = V7_PARTS == 3 ? 3 : 4;
constexpr int V7_PRODUCER_WARPS = 4 * V7_GROUPS;
constexpr int V7_EPI_WARPS = 4, V7_GEO_WARPS = 4 * V7_PARTS;
constexpr int V7_NT = V7_PARTS == 3 ? 3 : 4;                                      // taps a geometry thread computes in one straight-line section
constexpr int V7_FIRST_TAPS = V7_NT;                                              // ... of which part 0's are published first (geo_first)
constexpr int V7_NRV = V7_PARTS == 3 ? 9 : 15, V7_NENT = V7_PARTS == 3 ? 3 : 5;   // raw values / entries a geometry thread holds
// Every role that touches tensor memory keeps TMEM quarter = warp % 4 (all bases are multiples of four).  The helper roles
// take the low warp ids (measured 4 % faster than the other way round).
constexpr int V7_W_GEO = 0;                                                       // warps 0 .. 4 P - 1: part = warp / 4
constexpr int V7_W_EPI = V7_GEO_WARPS;                                            // four warps
constexpr int V7_W_MMA = V7_W_EPI + 4, V7_W_COPY = V7_W_MMA + 1, V7_W_BLOAD = V7_W_MMA + 2;   // the fourth warp of this quad idles
constexpr int V7_W_PROD = V7_W_MMA + 4;                                           // 4 x V7_GROUPS warps
static_assert(V7_W_PROD + V7_PRODUCER_WARPS == 32, "32 warps");
constexpr int V7_THREADS = 1024;
constexpr int V7_NA = 8, V7_NB = 3;
constexpr int V7_MAIN_BLOCKS = 9;
// Tail A operand of tile it: TMEM columns 80 / 104 / 208 / 232 (it % 4).  Four buffers make the hand-back implicit: a geometry
// warp writes tile it's buffer only after every producer warp has finished tile it - 2, and the producer of that tile's last
// block had to see block 9 (it - 2) consumed -- the MMAs are committed in order, so the tail MMAs of tile it - 3 (and it - 4,
// the previous user of the buffer) are complete.
__host__ __device__ constexpr uint32_t v7_tail_col(int it) { return (uint32_t)(((it >> 1) & 1) * TC_ACC_STRIDE + 80 + (it & 1) * 24); }
constexpr int V7_BOX_TAIL_BYTES = 7552;                                            // 468 x 16 B rounded up to a multiple of 128
constexpr uint32_t V7_BOX_TX = (uint32_t)V6_BOX_PX * (V6_MAIN_PX + V6_TAIL_PX);    // bytes one box load signals (zero fill included)
constexpr int V7_RAW_TMA = 0, V7_RAW_ROWS = 1, V7_RAW_LDG = 2;                    // how the offsets / masks of a tile arrive

struct __align__(1024) V7Smem {
  uint8_t b[V7_NB][TC_B_BYTES];                        // weight K blocks (bulk copies, SWIZZLE_128B image)            30,720
  uint8_t box_main[2][V6_BOX_PX * V6_MAIN_PX];         // TMA destination, [18][26][128 B]                           119,808
  uint8_t ostage[TC_M * TC_CMAX * 2];                  // epilogue staging tile (1024-byte aligned: TMA store, SWIZZLE_128B) 18,432
  uint4 geo[2][9][TC_M];                               // x: box byte offset | V6_SLOW, y/z: 4 bf16 weights, w: global pixel (slow)
  uint8_t box_tail[2][V7_BOX_TAIL_BYTES];              // TMA destination, [18][26][16 B]
  uint16_t raw[27][TC_M];                              // offsets / masks of the next tile
  unsigned long long full[V7_NA], done[V7_NA];         // main K block m: operands ready / MMAs complete (slot m % 8)
  unsigned long long tail_full[4], acc_full[2], acc_empty[2], geo_first[2], geo_full[2], geo_empty[2], box_full[2], box_empty[2];
  unsigned long long raw_full, raw_empty;
  uint32_t tmem_base;
};
static_assert(offsetof(V7Smem, ostage) % 1024 == 0, "TMA store with SWIZZLE_128B needs a 1024-byte aligned tile");
static_assert(offsetof(V7Smem, box_tail) % 128 == 0 && offsetof(V7Smem, raw) % 128 == 0 && offsetof(V7Smem, box_main) % 128 == 0,
              "TMA destinations are 128-byte aligned");
static_assert(sizeof(V7Smem) + 1024 <= 232448, "V7Smem exceeds the shared memory of an SM");

struct V7Args {
  TcParams p;
  int raw_mode;                                        // V7_RAW_*
  int use_tma_store;                                   // planes out through tm_out
  uint32_t o_main_px, o_tail_px;                       // bytes per pixel of the output planes
  alignas(64) CUtensorMap tm_main;                     // x main plane  {64, W, H, B} bf16, box {64, 26, 18, 1}
  alignas(64) CUtensorMap tm_tail;                     // x tail plane  { 8, W, H, B} bf16, box { 8, 26, 18, 1}
  alignas(64) CUtensorMap tm_raw0;                     // offsets (or the 27-channel conv output) {W, H, C, B}, box {16, 8, C, 1}
  alignas(64) CUtensorMap tm_raw1;                     // mask (non-fused form only)
  alignas(64) CUtensorMap tm_out;                      // out main plane {64, W, H, B}, box {64, 16, 8, 1}, SWIZZLE_128B
};

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src)
               : "memory");
}
// Non-blocking test of an mbarrier phase.  Under a saturated load/store queue every shared-memory operation -- a barrier poll
// included -- takes several hundred cycles to return, so the helper roles issue the polls (and the loads that depend on them)
// of a tile together and only look at the answers afterwards: one trip through the queue instead of one per barrier.
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok;
}
__device__ __forceinline__ void tmem_st_32x32b_x4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x2(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void epi7_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Modulated bilinear sample of the main plane from global memory (sample not served by the box); pixel stride `px`.
__device__ __forceinline__ void v7_sample_main_slow(const uint4& e, const uint8_t* x_main, uint32_t px, uint32_t row, uint32_t c_first,
                                                    uint32_t c_second, uint4& F, uint4& S) {
  const uint8_t* a00 = x_main + (unsigned long long)(e.w & 0x3fffffffu) * px;
  const uint8_t* a01 = a00 + ((e.w & 0x40000000u) ? px : 0);
  const uint32_t dy = (e.w & 0x80000000u) ? row : 0u;
  uint4 f[4], g[4];
  f[0] = __ldg(reinterpret_cast<const uint4*>(a00 + c_first)); f[1] = __ldg(reinterpret_cast<const uint4*>(a01 + c_first));
  f[2] = __ldg(reinterpret_cast<const uint4*>(a00 + dy + c_first)); f[3] = __ldg(reinterpret_cast<const uint4*>(a01 + dy + c_first));
  g[0] = __ldg(reinterpret_cast<const uint4*>(a00 + c_second)); g[1] = __ldg(reinterpret_cast<const uint4*>(a01 + c_second));
  g[2] = __ldg(reinterpret_cast<const uint4*>(a00 + dy + c_second)); g[3] = __ldg(reinterpret_cast<const uint4*>(a01 + dy + c_second));
  const uint2 w = make_uint2(e.y, e.z);
  F = lerp_chunk(f[0], f[1], f[2], f[3], w);
  S = lerp_chunk(g[0], g[1], g[2], g[3], w);
}
__device__ __forceinline__ void v7_sample_main_fast(const uint4& e, uint32_t bF, uint32_t bS, uint4& F, uint4& S) {
  const uint32_t aF = bF + e.x, aS = bS + e.x;
  uint4 f[4], g[4];
  f[0] = lds16o<0>(aF); f[1] = lds16o<V6_MAIN_PX>(aF); f[2] = lds16o<V6_MAIN_ROW>(aF); f[3] = lds16o<V6_MAIN_ROW + V6_MAIN_PX>(aF);
  g[0] = lds16o<0>(aS); g[1] = lds16o<V6_MAIN_PX>(aS); g[2] = lds16o<V6_MAIN_ROW>(aS); g[3] = lds16o<V6_MAIN_ROW + V6_MAIN_PX>(aS);
  const uint2 w = make_uint2(e.y, e.z);
  F = lerp_chunk(f[0], f[1], f[2], f[3], w);
  S = lerp_chunk(g[0], g[1], g[2], g[3], w);
}
// The four tail channels (8 bytes) of sample `e`; global pixel stride `px` for samples the box does not serve.
__device__ __forceinline__ uint2 v7_sample_tail(const uint4& e, uint32_t box_tail, const uint8_t* x_tail, uint32_t px, uint32_t row,
                                                uint32_t half) {
  uint2 v[4];
  if ((int)e.x >= 0) {
    const uint32_t a = box_tail + (e.x >> 3) + half;               // 16 B per pixel instead of 128
    v[0] = lds8o<0>(a); v[1] = lds8o<V6_TAIL_PX>(a); v[2] = lds8o<V6_TAIL_ROW>(a); v[3] = lds8o<V6_TAIL_ROW + V6_TAIL_PX>(a);
  } else {
    const uint8_t* a00 = x_tail + (unsigned long long)(e.w & 0x3fffffffu) * px;
    const uint8_t* a01 = a00 + ((e.w & 0x40000000u) ? px : 0);
    const uint32_t dy = (e.w & 0x80000000u) ? row : 0u;
    v[0] = __ldg(reinterpret_cast<const uint2*>(a00)); v[1] = __ldg(reinterpret_cast<const uint2*>(a01));
    v[2] = __ldg(reinterpret_cast<const uint2*>(a00 + dy)); v[3] = __ldg(reinterpret_cast<const uint2*>(a01 + dy));
  }
  const uint4 r = lerp_chunk(make_uint4(v[0].x, v[0].y, 0u, 0u), make_uint4(v[1].x, v[1].y, 0u, 0u),
                             make_uint4(v[2].x, v[2].y, 0u, 0u), make_uint4(v[3].x, v[3].y, 0u, 0u), make_uint2(e.y, e.z));
  return make_uint2(r.x, r.y);
}

// TO: 16-bit dtype of the offset / mask tensors; TOUT: output dtype; FUSED27: offsets and mask come from the 27-channel
// offset_conv output; PLANES: output as bf16 planes (else any strided tensor); DBG: per-role cycle counters.
template <typename TO, typename TOUT, bool FUSED27, bool PLANES, bool DBG>
__global__ void __launch_bounds__(V7_THREADS, 1) dcn_tc7_fwd_kernel(const __grid_constant__ V7Args a) {
  const TcParams& p = a.p;
  extern __shared__ uint8_t smem_raw[];
  V7Smem& s = *reinterpret_cast<V7Smem*>(smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023));
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform: role branches stay on the uniform data path

  if (tid == 0) {
    for (int i = 0; i < V7_NA; ++i) {
      mbar_init(smem_u32(&s.full[i]), 5);                         // four producer warps + the weight loader's expect_tx arrival
      mbar_init(smem_u32(&s.done[i]), 1);                         // one tcgen05.commit
    }
    // One tail barrier per tail buffer (it % 4): without a hand-back wait, a two-slot barrier could see the geometry warps of
    // tile it + 2 arrive before tile it's phase was complete (the weight loader arrives late in the tile).
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&s.tail_full[i]), V7_GEO_WARPS + 1);   // eight geometry warps + the weight loader
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s.acc_full[i]), 1);
      mbar_init(smem_u32(&s.acc_empty[i]), V7_EPI_WARPS);
      mbar_init(smem_u32(&s.geo_first[i]), 4);                    // part 0's taps written (four warps)
      mbar_init(smem_u32(&s.geo_full[i]), V7_GEO_WARPS);
      mbar_init(smem_u32(&s.geo_empty[i]), V7_PRODUCER_WARPS);
      mbar_init(smem_u32(&s.box_full[i]), 1);
      mbar_init(smem_u32(&s.box_empty[i]), V7_PRODUCER_WARPS + V7_GEO_WARPS);   // the geometry warps gather the tail channels
    }
    mbar_init(smem_u32(&s.raw_full), 1);
    mbar_init(smem_u32(&s.raw_empty), V7_GEO_WARPS);
    fence_barrier_init();
  }
  if (warp == V7_W_COPY && lane == 0) {
    tma_prefetch_desc(&a.tm_main);
    tma_prefetch_desc(&a.tm_tail);
    if (a.raw_mode == V7_RAW_TMA) { tma_prefetch_desc(&a.tm_raw0); if (!FUSED27) tma_prefetch_desc(&a.tm_raw1); }
    if (PLANES && a.use_tma_store) tma_prefetch_desc(&a.tm_out);
  }
  if (warp == V7_W_MMA) tmem_alloc(smem_u32(&s.tmem_base), V6_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;
  const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile0 = (int)blockIdx.x, tile_step = (int)gridDim.x;
  constexpr bool dbg = DBG;
  long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0;      // cycles spent in this role's waits (debug only)
  const long long t_begin = clock64();

  if (warp >= V7_W_PROD && warp < V7_W_PROD + V7_PRODUCER_WARPS) {
    // =========================================================================== A-operand producers
    // Main block m = 9 it + kb goes to group m % 4 and A-ring stage m % 8; warp q of a group owns TMEM lanes [32q, 32q + 32).
    const int group = (warp - V7_W_PROD) >> 2, q = warp & 3;
    const int g = lane >> 2, u = lane & 3;
    const bool par = (g & 1) != 0;
    const uint32_t c_first = (uint32_t)(u + (par ? 4 : 0)) * 16, c_second = (uint32_t)(u + (par ? 0 : 4)) * 16;
    const uint32_t main_row = p.main_stride * (uint32_t)p.W;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    for (int it = 0; it < my_tiles; ++it) {
      const int gb = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      const uint32_t box_main = smem_u32(&s.box_main[gb][0]);
      const uint32_t bF = box_main + c_first, bS = box_main + c_second;
      bool first = true, all_taps = false;
      // four groups: block m = 9 it + kb goes to group m % 4 (kb = (group - it) mod 4, + 4, ...); three groups: kb = group, + 3, + 6
      for (int kb = V7_GROUPS == 4 ? ((group - it) & 3) : group; kb < V7_MAIN_BLOCKS; kb += V7_GROUPS) {
        const int m = it * V7_MAIN_BLOCKS + kb, sa = m & (V7_NA - 1);
        const uint32_t a_taddr = tmem_base + lane_base + (uint32_t)(V6_A_COL0 + sa * 32);
        if (first) {
          mbar_wait_d<V6_NS_PROD, DBG>(smem_u32(&s.geo_first[gb]), tphase, w0);   // this tile's geometry, taps 0..3
          mbar_wait_d<V6_NS_PROD, DBG>(smem_u32(&s.box_full[gb]), tphase, w1);    // this tile's source box has landed
          first = false;
        }
        if (kb >= V7_FIRST_TAPS && !all_taps) {
          mbar_wait_d<V6_NS_PROD, DBG>(smem_u32(&s.geo_full[gb]), tphase, w0);
          all_taps = true;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint4 e0 = s.geo[gb][kb][q * 32 + h * 16 + g], e1 = s.geo[gb][kb][q * 32 + h * 16 + g + 8];
          uint4 F0, S0, F1, S1;
          if ((int)(e0.x | e1.x) >= 0) {                  // both samples served by the box: one basic block
            v7_sample_main_fast(e0, bF, bS, F0, S0);
            v7_sample_main_fast(e1, bF, bS, F1, S1);
          } else {
            if ((int)e0.x >= 0) v7_sample_main_fast(e0, bF, bS, F0, S0);
            else v7_sample_main_slow(e0, p.x_main, p.main_stride, main_row, c_first, c_second, F0, S0);
            if ((int)e1.x >= 0) v7_sample_main_fast(e1, bF, bS, F1, S1);
            else v7_sample_main_slow(e1, p.x_main, p.main_stride, main_row, c_first, c_second, F1, S1);
          }
          // chunk u (X) and chunk u + 4 (Y) of both pixels: odd g loaded them in the opposite order
          const uint4 X0 = par ? S0 : F0, Y0 = par ? F0 : S0, X1 = par ? S1 : F1, Y1 = par ? F1 : S1;
          const uint32_t r[16] = {X0.x, X0.y, X1.x, X1.y, X0.z, X0.w, X1.z, X1.w,
                                  Y0.x, Y0.y, Y1.x, Y1.y, Y0.z, Y0.w, Y1.z, Y1.w};
          if (h == 0 && m >= V7_NA) {                            // the ring stage is needed only now, after the gathers
            mbar_wait_d<V6_NS_STAGE, DBG>(smem_u32(&s.done[sa]), (uint32_t)((m - V7_NA) >> 3) & 1u, w2);
            tc_fence_after();
          }
          tmem_st_16x256b_x4(a_taddr + ((uint32_t)(h * 16) << 16), r);
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
  } else if (warp == V7_W_MMA) {
    // =========================================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(TC_M, TC_N);
    const uint32_t b_smem = smem_u32(&s.b[0][0]);
    uint32_t sa = 0, a_par = 0, sb = 0;
    // The poll of block n + 1 is issued before block n is waited for: in steady state the MMA warp never pays the
    // load/store-queue latency of a barrier poll between two blocks.
    uint32_t ok_next = my_tiles > 0 ? mbar_test(smem_u32(&s.full[0]), 0u) : 0u;
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t acc = (uint32_t)it & 1u, acc_phase = ((uint32_t)it >> 1) & 1u;
      mbar_wait_d<32, DBG>(smem_u32(&s.acc_empty[acc]), acc_phase ^ 1, w0);   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * TC_ACC_STRIDE;
#pragma unroll 1
      for (int kb = 0; kb < V7_MAIN_BLOCKS; ++kb) {
        const uint32_t ok = ok_next;
        {
          const uint32_t sn = sa + 1 == V7_NA ? 0u : sa + 1, pn = sa + 1 == V7_NA ? a_par ^ 1u : a_par;
          // after the last main block of the tile comes the tail block: poll its barrier instead
          ok_next = kb + 1 < V7_MAIN_BLOCKS ? mbar_test(smem_u32(&s.full[sn]), pn) : mbar_test(smem_u32(&s.tail_full[it & 3]), (uint32_t)(it >> 2) & 1u);
        }
        if (!ok) mbar_wait_d<V6_NS_MMA, DBG>(smem_u32(&s.full[sa]), a_par, w2);
        tc_fence_after();
        const uint32_t a_tmem = tmem_base + (uint32_t)V6_A_COL0 + sa * 32u;
        const uint64_t bdesc = umma_desc_sw128(b_smem + sb * (uint32_t)TC_B_BYTES);
        const long long ti0 = dbg ? clock64() : 0;
        if (!(DBG && (p.experiment & 2))) {              // diagnostics: bit 1 = issue no MMA
          umma_bf16_ts(d_tmem, a_tmem, bdesc, idesc, kb != 0);
          umma_bf16_ts(d_tmem, a_tmem + 8, bdesc + 2, idesc, 1);
          umma_bf16_ts(d_tmem, a_tmem + 16, bdesc + 4, idesc, 1);
          umma_bf16_ts(d_tmem, a_tmem + 24, bdesc + 6, idesc, 1);
        }
        const long long ti1 = dbg ? clock64() : 0;
        umma_commit_elect(smem_u32(&s.done[sa]));
        if (dbg) { w3 += ti1 - ti0; w4 += clock64() - ti1; }
        if (++sa == V7_NA) { sa = 0; a_par ^= 1u; }
        if (++sb == V7_NB) sb = 0;
      }
      // tail block: A = what the geometry warps wrote next to the accumulator (K = 36 tail samples + the bias slots, 48 in all)
      {
        const uint32_t ok = ok_next;
        ok_next = it + 1 < my_tiles ? mbar_test(smem_u32(&s.full[sa]), a_par) : 0u;     // first block of the next tile
        if (!ok) mbar_wait_d<V6_NS_MMA, DBG>(smem_u32(&s.tail_full[it & 3]), (uint32_t)(it >> 2) & 1u, w1);
      }
      tc_fence_after();
      {
        const uint32_t a_tmem = tmem_base + v7_tail_col(it);
        const uint64_t bdesc = umma_desc_sw128(b_smem + sb * (uint32_t)TC_B_BYTES);
        umma_bf16_ts(d_tmem, a_tmem, bdesc, idesc, 1);
        umma_bf16_ts(d_tmem, a_tmem + 8, bdesc + 2, idesc, 1);
        umma_bf16_ts(d_tmem, a_tmem + 16, bdesc + 4, idesc, 1);
        umma_commit_elect(smem_u32(&s.acc_full[acc]));
        if (++sb == V7_NB) sb = 0;
      }
    }
    __syncwarp();
  } else if (warp == V7_W_BLOAD) {
    // =========================================================================== weight-block loader (one lane)
    // Ten blocks per tile into a 3-stage ring: j < 9 -> main block (arrives on full[m % 8]), j = 9 -> tail block (tail_full).
    if (lane == 0) {
      int nb = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int j = 0; j < V7_MAIN_BLOCKS + 1; ++j, ++nb) {
          const int sb = nb % V7_NB;
          if (nb >= V7_NB) {                                  // the block that used this stage last (nb - 3) has been consumed
            const int pj = j - V7_NB, pit = pj < 0 ? it - 1 : it, pjj = pj < 0 ? pj + V7_MAIN_BLOCKS + 1 : pj;
            if (pjj == V7_MAIN_BLOCKS) {
              mbar_wait_d<V6_NS_LOAD, DBG>(smem_u32(&s.acc_full[pit & 1]), (uint32_t)(pit >> 1) & 1u, w0);
            } else {
              const int pm = pit * V7_MAIN_BLOCKS + pjj;
              mbar_wait_d<V6_NS_LOAD, DBG>(smem_u32(&s.done[pm & (V7_NA - 1)]), (uint32_t)(pm >> 3) & 1u, w0);
            }
          }
          const uint32_t bar = j < V7_MAIN_BLOCKS ? smem_u32(&s.full[(it * V7_MAIN_BLOCKS + j) & (V7_NA - 1)]) : smem_u32(&s.tail_full[it & 3]);
          mbar_arrive_expect_tx(bar, TC_B_BYTES);
          bulk_g2s(smem_u32(&s.b[sb][0]), p.wpacked + (size_t)j * TC_B_BYTES, TC_B_BYTES, bar);
        }
      }
    }
    __syncwarp();
  } else if (warp == V7_W_COPY) {
    // =========================================================================== source box + offsets / masks (copy engine)
    for (int it = 0; it < my_tiles; ++it) {
      const int gb = it & 1;
      int b, ty0, tx0;
      tile_origin(p, tile0 + it * tile_step, b, ty0, tx0);
      if (a.raw_mode != V7_RAW_LDG) {
        // the raw buffer is single: tile it - 1's values have been read into registers by all geometry warps
        if (it > 0) mbar_wait_d<V6_NS_HELP, DBG>(smem_u32(&s.raw_empty), (uint32_t)(it - 1) & 1u, w1);
        const uint32_t bar = smem_u32(&s.raw_full);
        if (a.raw_mode == V7_RAW_TMA) {
          if (lane == 0) {
            mbar_arrive_expect_tx(bar, 27u * TC_M * 2u);
            if (FUSED27) {
              tma_load_4d(smem_u32(&s.raw[0][0]), &a.tm_raw0, tx0, ty0, 0, b, bar);
            } else {
              tma_load_4d(smem_u32(&s.raw[0][0]), &a.tm_raw0, tx0, ty0, 0, b, bar);
              tma_load_4d(smem_u32(&s.raw[18][0]), &a.tm_raw1, tx0, ty0, 0, b, bar);
            }
          }
        } else {
          // dense channels-last offset_conv output: a tile row is one contiguous run of cols x 27 values, copied as it lies
          const int rows = min(TC_TH, p.H - ty0), cols = min(TC_TW, p.W - tx0);
          const uint32_t row_bytes = (uint32_t)cols * 54u;
          if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)rows * row_bytes);
          __syncwarp();
          if (lane < rows) {
            const TO* src = reinterpret_cast<const TO*>(p.offset) + b * p.f_sn + ((long long)(ty0 + lane) * p.W + tx0) * 27;
            bulk_g2s(smem_u32(&s.raw[0][0]) + (uint32_t)lane * (TC_TW * 54u), src, row_bytes, bar);
          }
        }
      }
      mbar_wait_d<V6_NS_HELP, DBG>(smem_u32(&s.box_empty[gb]), ((uint32_t)(it >> 1) & 1u) ^ 1u, w0);   // readers are done with the old box
      if (lane == 0) {
        const uint32_t bar = smem_u32(&s.box_full[gb]);
        mbar_arrive_expect_tx(bar, V7_BOX_TX);
        tma_load_4d(smem_u32(&s.box_main[gb][0]), &a.tm_main, 0, tx0 - V6_BOX_LEFT, ty0 - V6_BOX_TOP, b, bar);
        tma_load_4d(smem_u32(&s.box_tail[gb][0]), &a.tm_tail, 0, tx0 - V6_BOX_LEFT, ty0 - V6_BOX_TOP, b, bar);
      }
      __syncwarp();
    }
  } else if (warp >= V7_W_GEO && warp < V7_W_GEO + V7_GEO_WARPS) {
    // =========================================================================== tap geometry + tail channels
    // Thread = (tile row, part).  Three parts: taps 3 part .. 3 part + 2.  Two parts: part 0 = taps 0..3 and then tap 8,
    // part 1 = taps 4..7.  Part 0's first section is published first (geo_first: what the first K blocks of a tile need).
    // The entries stay in registers for the tail gather, whose results go straight to this tile's tail A columns in TMEM.
    const int part = (warp - V7_W_GEO) >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t tail_row = p.tail_stride * (uint32_t)p.W;
    const uint16_t* const raw_flat = &s.raw[0][0];
    const int raw_mode = a.raw_mode;
    const int kbase = part * V7_NT;                      // taps of the straight-line section: kbase .. kbase + V7_NT - 1
    const bool extra = V7_PARTS == 2 && part == 0;       // ... plus tap 8 (two-part form only)
    auto tap_of = [&](int i) -> int { return i < V7_NT ? kbase + i : 8; };
    // Offsets / masks of this thread's taps of tile `t` -- (dy, dx, mask) each.  Only issues the loads (independent LDS.U16 /
    // LDG), one branch on the transport for all of them: a branch per value would put every load in its own basic block,
    // i.e. one trip through the saturated load/store queue per value (measured: 7,000 cycles per tile).
    auto fetch_raw = [&](int t, uint32_t* rv) {
      if (DBG && (p.experiment & 8) && t >= 1) return;   // diagnostics: the values of tile 0 are reused
      int b, ty0, tx0;
      tile_origin(p, tile0 + t * tile_step, b, ty0, tx0);
      const int y = ty0 + row / TC_TW, x = tx0 + row % TC_TW;
      const bool inside = y < p.H && x < p.W;
      // kernel channel c (0..17: offsets dy/dx interleaved, 18..26: mask) -> channel of the tensor the values come from
      auto chan = [&](int c) -> int { return FUSED27 ? (c < 18 ? (c < 9 ? c : c + 9) : c - 9) : c; };
      auto all = [&](auto&& rd) {
#pragma unroll
        for (int i = 0; i < V7_NENT; ++i) {
          const int k = tap_of(i);
          rv[3 * i] = rd(2 * k); rv[3 * i + 1] = rd(2 * k + 1); rv[3 * i + 2] = rd(18 + k);
        }
      };
      if (raw_mode == V7_RAW_TMA) {
        all([&](int c) -> uint32_t { return s.raw[chan(c)][row]; });
      } else if (raw_mode == V7_RAW_ROWS) {
        all([&](int c) -> uint32_t { return raw_flat[row * 27 + chan(c)]; });
      } else if (inside) {
        const TO* po = reinterpret_cast<const TO*>(p.offset) + b * p.f_sn + (long long)y * p.f_sh + (long long)x * p.f_sw;
        const TO* pm = reinterpret_cast<const TO*>(p.mask) + b * p.m_sn + (long long)y * p.m_sh + (long long)x * p.m_sw;
        all([&](int c) -> uint32_t { return (FUSED27 || c < 18) ? ldg_bits<TO>(po + chan(c) * p.f_sc) : ldg_bits<TO>(pm + (c - 18) * p.m_sc); });
      } else {
#pragma unroll
        for (int i = 0; i < V7_NRV; ++i) rv[i] = 0u;
      }
    };
    // The raw buffer is single: it is handed back once every lane holds its values in registers.
    auto release_raw = [&](uint32_t* rv) {
      if (raw_mode == V7_RAW_LDG) return;
#pragma unroll
      for (int i = 0; i < V7_NRV; ++i) asm volatile("" ::"r"(rv[i]) : "memory");      // the loads have returned
      __syncwarp();                                    // every lane's reads are ordered before the release below
      if (lane == 0) mbar_arrive(smem_u32(&s.raw_empty));
    };
    uint32_t rv[V7_NRV];
    for (int it = 0; it < my_tiles; ++it) {
      const int gb = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      int b, ty0, tx0;
      tile_origin(p, tile0 + it * tile_step, b, ty0, tx0);
      const int y = ty0 + row / TC_TW, x = tx0 + row % TC_TW;
      const bool inside = y < p.H && x < p.W;
      // Optimistic and batched: the three barrier polls of this tile and the loads go through the load/store queue together;
      // the loads are simply repeated in the rare case that the values had not landed yet.
      const long long tr0 = dbg ? clock64() : 0;
      const uint32_t ok_raw = raw_mode != V7_RAW_LDG ? mbar_test(smem_u32(&s.raw_full), (uint32_t)it & 1u) : 1u;
      const uint32_t ok_geo = mbar_test(smem_u32(&s.geo_empty[gb]), tphase ^ 1u);
      const uint32_t ok_box = mbar_test(smem_u32(&s.box_full[gb]), tphase);
      fetch_raw(it, rv);
      if (!ok_raw) {
        mbar_wait_d<V6_NS_HELP, DBG>(smem_u32(&s.raw_full), (uint32_t)it & 1u, w1);
        fetch_raw(it, rv);
      }
      release_raw(rv);
      if (dbg) w4 += clock64() - tr0;
      if (!ok_geo) mbar_wait_d<V6_NS_HELP, DBG>(smem_u32(&s.geo_empty[gb]), tphase ^ 1u, w0);   // producers are done with the old entries
      const long long tg0 = dbg ? clock64() : 0;
      const int by0 = ty0 - V6_BOX_TOP, bx0 = tx0 - V6_BOX_LEFT;
      const int base = b * p.H * p.W;
      const float fy0 = (float)(y - 1), fx0 = (float)(x - 1);
      uint4 ent[V7_NENT];
      // Straight-line code for the section's taps (no branch per tap: the dependent chains -- exp, reciprocal, floor, products
      // -- of the taps interleave); samples the box does not serve are rare and patched afterwards.
      auto entry = [&](int i, int k, uint32_t& slow) {
        float mk = bits_to_f32<TO>(rv[3 * i + 2]);
        // the sigmoid result is rounded to the tensor dtype, as torch.sigmoid on that tensor would
        if (FUSED27) mk = to_f32<TO>(from_f32<TO>(__fdividef(1.0f, 1.0f + __expf(-mk))));
        const float dy = bits_to_f32<TO>(rv[3 * i]), dx = bits_to_f32<TO>(rv[3 * i + 1]);
        uint4 e;
        if (!v6_geo_entry(by0, bx0, fy0 + (float)(k / 3), fx0 + (float)(k % 3), dy, dx, mk, e)) slow |= 1u << i;
        ent[i] = e;
      };
      auto patch = [&](int i, int k) {
        float mk = bits_to_f32<TO>(rv[3 * i + 2]);
        if (FUSED27) mk = to_f32<TO>(from_f32<TO>(__fdividef(1.0f, 1.0f + __expf(-mk))));
        ent[i] = v6_geo_entry_slow(p.H, p.W, base, y, x, k, bits_to_f32<TO>(rv[3 * i]), bits_to_f32<TO>(rv[3 * i + 1]), mk);
      };
      const bool skip_geo = DBG && (p.experiment & 4) && it >= 2;      // diagnostics: the entries of tiles 0 / 1 are reused
      if (skip_geo) {
      } else if (inside) {
        uint32_t slow = 0;
#pragma unroll
        for (int i = 0; i < V7_NT; ++i) entry(i, kbase + i, slow);
        if (slow) {
#pragma unroll
          for (int i = 0; i < V7_NT; ++i)
            if (slow & (1u << i)) patch(i, kbase + i);
        }
      } else {
#pragma unroll
        for (int i = 0; i < V7_NT; ++i) ent[i] = make_uint4(V6_SAFE, 0u, 0u, 0u);
      }
      if (!skip_geo) {
#pragma unroll
        for (int i = 0; i < V7_NT; ++i) s.geo[gb][kbase + i][row] = ent[i];
      }
      if (part == 0) {                                   // the first taps are all the first K blocks of the tile need
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s.geo_first[gb]));
      }
      if (V7_NENT > V7_NT) {                             // two-part form: tap 8 goes to part 0
        ent[V7_NENT - 1] = make_uint4(V6_SAFE, 0u, 0u, 0u);
        if (extra) {
          if (inside && !skip_geo) {
            uint32_t slow = 0;
            entry(V7_NENT - 1, 8, slow);
            if (slow) patch(V7_NENT - 1, 8);
          }
          if (!skip_geo) s.geo[gb][8][row] = ent[V7_NENT - 1];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s.geo_full[gb]));
      if (dbg) w2 += clock64() - tg0;
      // ---- tail channels of this thread's taps: box (tail plane) of this tile, then this tile's tail A columns
      if (!ok_box) mbar_wait_d<V6_NS_HELP, DBG>(smem_u32(&s.box_full[gb]), tphase, w3);
      tc_fence_after();
      const long long tt0 = dbg ? clock64() : 0;
      const uint32_t box_tail = smem_u32(&s.box_tail[gb][0]);
      const uint32_t hsel = (uint32_t)(lane & 1) * 8u;
      const uint32_t taddr = tmem_base + lane_base + v7_tail_col(it);
      uint32_t r[2 * V7_NENT];
      uint32_t all_x = 0;
#pragma unroll
      for (int i = 0; i < V7_NENT; ++i) all_x |= ent[i].x;
      if (DBG && (p.experiment & 1)) {                 // diagnostics: no tail gathers
#pragma unroll
        for (int i = 0; i < 2 * V7_NENT; ++i) r[i] = 0u;
      } else if ((int)all_x >= 0) {
        // every sample of this thread is served by the box: straight-line code, all loads in flight together (one trip
        // through the load/store queue instead of one per tap)
        uint2 c[V7_NENT][4];
#pragma unroll
        for (int i = 0; i < V7_NENT; ++i) {
          const uint32_t ad = box_tail + (ent[i].x >> 3) + hsel;
          c[i][0] = lds8o<0>(ad); c[i][1] = lds8o<V6_TAIL_PX>(ad); c[i][2] = lds8o<V6_TAIL_ROW>(ad); c[i][3] = lds8o<V6_TAIL_ROW + V6_TAIL_PX>(ad);
        }
#pragma unroll
        for (int i = 0; i < V7_NENT; ++i) {
          const uint4 t = lerp_chunk(make_uint4(c[i][0].x, c[i][0].y, 0u, 0u), make_uint4(c[i][1].x, c[i][1].y, 0u, 0u),
                                     make_uint4(c[i][2].x, c[i][2].y, 0u, 0u), make_uint4(c[i][3].x, c[i][3].y, 0u, 0u),
                                     make_uint2(ent[i].y, ent[i].z));
          r[2 * i] = t.x; r[2 * i + 1] = t.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < V7_NENT; ++i) {
          const uint2 v = v7_sample_tail(ent[i], box_tail, p.x_tail, p.tail_stride, tail_row, hsel);
          r[2 * i] = v.x; r[2 * i + 1] = v.y;
        }
      }
      // TMEM columns of the tail A block: tap k at 2 k, 2 k + 1 (K elements 4 k .. 4 k + 3); column 18 = K elements 36, 37
      // = 1.0 (the weight image holds bias hi / lo there); columns 19..23 zero.
      const uint32_t ones[6] = {0x3f803f80u, 0u, 0u, 0u, 0u, 0u};
      if (V7_PARTS == 3) {
        tmem_st_32x32b_x4(taddr + (uint32_t)(6 * part), r);
        tmem_st_32x32b_x2(taddr + (uint32_t)(6 * part + 4), r + 4);
        if (part == 2) {
          tmem_st_32x32b_x4(taddr + 18u, ones);
          tmem_st_32x32b_x2(taddr + 22u, ones + 4);
        }
      } else {
        tmem_st_32x32b_x8(taddr + (part == 0 ? 0u : 8u), r);     // taps 0..3 / 4..7
        if (part == 0) {
          const uint32_t r2[8] = {r[2 * (V7_NENT - 1)], r[2 * (V7_NENT - 1) + 1], 0x3f803f80u, 0u, 0u, 0u, 0u, 0u};
          tmem_st_32x32b_x8(taddr + 16u, r2);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&s.tail_full[it & 3]));
        mbar_arrive(smem_u32(&s.box_empty[gb]));
      }
      if (dbg) w5 += clock64() - tt0;
    }
  } else if (warp >= V7_W_EPI && warp < V7_W_EPI + V7_EPI_WARPS) {
    // =========================================================================== epilogue (4 warps)
    const int quad = warp & 3;                           // TMEM lanes [32*quad, 32*quad + 32) belong to this warp
    const int row = quad * 32 + lane;                    // tile row = TMEM lane of this thread
    const int etid = (warp - V7_W_EPI) * 32 + lane;      // 0..127 for the cooperative store
    const uint32_t ostage = smem_u32(&s.ostage[0]);
    const int cs_x = etid >> 3, cs_c = etid & 7;
    const uint32_t cs_src = ostage + (uint32_t)cs_x * 128 + ((uint32_t)(cs_c ^ (cs_x & 7)) << 4);   // + 2048 per tile row (16 % 8 == 0)
    const size_t cs_row = (size_t)p.W * a.o_main_px;
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
      uint8_t* ot = reinterpret_cast<uint8_t*>(p.out_tail) + pixel * a.o_tail_px;
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
        epi7_bar_sync();
        const int cols = min(TC_TW, p.W - tx0);
        const int seg16 = cols * p.O / 8;                  // 16-byte units per tile row (W % 8 == 0)
        const uint32_t pitch = (uint32_t)(TC_TW * p.O * 2);
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + ((size_t)b * p.o_sn + ((size_t)ty0 * p.W + tx0) * p.O) * 2;
        for (int i = 0; i < TC_TH && ty0 + i < p.H; ++i)
          for (int j = etid; j < seg16; j += 128)
            *reinterpret_cast<uint4*>(dst + (size_t)i * p.W * p.O * 2 + 16 * j) = lds16(ostage + i * pitch + 16 * j);
        epi7_bar_sync();
      }
      if (PLANES) {
        if (a.use_tma_store) {
          // main plane: the swizzled staging tile leaves as one TMA store (clipped at the image border by the copy engine)
          fence_proxy_async();                             // this thread's staging writes -> visible to the async proxy
          epi7_bar_sync();
          if (etid == 0) {
            tma_store_4d(&a.tm_out, 0, tx0, ty0, b, ostage);
            tma_store_commit();
            tma_store_wait_read();                         // the staging tile has been read: it may be overwritten
          }
          epi7_bar_sync();
        } else {
          epi7_bar_sync();
          const int xx = tx0 + cs_x;
          uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + ((size_t)(b * p.H + ty0) * p.W + xx) * a.o_main_px + cs_c * 16;
          if (xx < p.W) {
#pragma unroll
            for (int i = 0; i < TC_TH; ++i)
              if (ty0 + i < p.H) *reinterpret_cast<uint4*>(dst + i * cs_row) = lds16(cs_src + i * (TC_TW * 128));
          }
          epi7_bar_sync();
        }
      }
      if (dbg) w3 += clock64() - te0;
    }
  }

  if (dbg && lane == 0) {
    unsigned long long* d = p.debug + ((size_t)blockIdx.x * 32 + warp) * 8;
    d[0] = (unsigned long long)(clock64() - t_begin); d[1] = w0; d[2] = w1; d[3] = w2; d[4] = w3; d[5] = my_tiles; d[6] = w4; d[7] = w5;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == V7_W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, V6_TMEM_COLS);
  }
}
