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
//     by bulk copies one tile ahead (double buffered), so every corner is an LDS.128 with no L1 tag / miss traffic;
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

constexpr int V6_BOX_H = 18, V6_BOX_W = 26, V6_BOX_PX = V6_BOX_H * V6_BOX_W;      // 468 pixels
constexpr int V6_BOX_TOP = 5, V6_BOX_LEFT = 5;                                    // box origin = tile origin - (5, 5)
constexpr int V6_PRODUCER_WARPS = 16;                                             // 4 groups x 4 TMEM sub-partitions
constexpr int V6_W_MMA = 16, V6_W_BOX = 17, V6_W_BLOAD = 18, V6_W_BACK = 20;      // warp 19 idles; warps 20..27: back end
constexpr int V6_BACK_WARPS = 8;                                                  // geometry + epilogue, two per TMEM quarter
constexpr int V6_THREADS = 28 * 32;                                               // 896
constexpr int V6_KBLOCKS = 10;                                                    // 9 main + 1 tail
constexpr int V6_NA = 8, V6_NB = 3;                                               // TMEM A ring / smem B ring depth
constexpr int V6_A_COL0 = 256;                                                    // TMEM columns [256, 512): A ring
constexpr int V6_TMEM_COLS = 512;
constexpr uint32_t V6_INSIDE = 0x80000000u;
// Box index of the tile's own first pixel: always inside the image and always copied.  Zero-weight entries (dead samples,
// rows of a partial tile) point here so that they never multiply 0 by uninitialised shared memory.
constexpr uint32_t V6_SAFE = V6_BOX_TOP * V6_BOX_W + V6_BOX_LEFT;

__host__ __device__ inline void v6_k_to_tap_channel(int kb, int kk, int& tap, int& c) {
  if (kb < 9) { tap = kb; c = ((kk >> 5) << 5) + 8 * ((kk >> 2) & 3) + 4 * ((kk >> 4) & 1) + (kk & 3); }
  else if (kb == 9 && kk < 36) { tap = kk >> 2; c = TC_CMAIN + (kk & 3); }
  else { tap = 0; c = -1; }
}

struct __align__(1024) V6Smem {
  uint8_t b[V6_NB][TC_B_BYTES];                        // weight K blocks (bulk copies, SWIZZLE_128B image)
  uint8_t box_main[2][V6_BOX_PX * TC_CMAIN * 2];       // 2 x 59,904 B
  uint8_t box_tail[2][V6_BOX_PX * TC_CTAIL * 2];       // 2 x  7,488 B
  uint4 geo[2][9][TC_M];                               // x: box index + flags, y/z: 4 bf16 weights, w: global pixel (v4 pixf)
  uint8_t ostage[TC_M * 128];                          // epilogue staging tile (chunk j of row r at j ^ (r & 7))
  unsigned long long a_full[V6_NA], a_empty[V6_NA], b_full[V6_NB], b_empty[V6_NB];
  unsigned long long acc_full[2], acc_empty[2], geo_full[2], geo_empty[2], box_full[2], box_empty[2];
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
__device__ __forceinline__ void sts16(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// A-from-TMEM form of the MMA: A = 128 lanes x 8 columns (16 bf16 of K per lane) at a_tmem, B from shared memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 back-end warps

// Sampling geometry of tap k at output pixel (y, x): Appendix B of SURVEY.md, corners clamped into the image (zero weight
// where torchvision skips a corner), located inside the staged box when all four corners are there.
__device__ __forceinline__ uint4 v6_geo_entry(int H, int W, int base, int by0, int bx0, int y, int x, int k, float dy, float dx,
                                              float mk) {
  float py = (float)(y - 1 + k / 3) + dy;
  float px = (float)(x - 1 + k % 3) + dx;
  const bool live = (py > -1.0f) && (py < (float)H) && (px > -1.0f) && (px < (float)W);
  if (!live) { py = -2.0f; px = -2.0f; mk = 0.0f; }
  const float fy = floorf(py), fx = floorf(px);
  const int y0 = (int)fy, x0 = (int)fx;
  const float lh = py - fy, lw = px - fx, hh = 1.0f - lh, hw = 1.0f - lw;
  const bool r0 = (unsigned)y0 < (unsigned)H, r1 = (unsigned)(y0 + 1) < (unsigned)H;
  const bool c0 = (unsigned)x0 < (unsigned)W, c1 = (unsigned)(x0 + 1) < (unsigned)W;
  const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
  const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
  const int sx = cx1 - cx0, sy = cy1 - cy0;
  uint2 wq;
  store_geo_w(wq, (r0 && c0) ? hh * hw * mk : 0.0f, (r0 && c1) ? hh * lw * mk : 0.0f, (r1 && c0) ? lh * hw * mk : 0.0f,
              (r1 && c1) ? lh * lw * mk : 0.0f);
  const bool dead = ((wq.x | wq.y) & 0x7fff7fffu) == 0u;           // all four weights are +-0: the value is irrelevant
  const int ry = cy0 - by0, rx = cx0 - bx0;
  const bool in = ry >= 0 && ry + sy < V6_BOX_H && rx >= 0 && rx + sx < V6_BOX_W;
  uint4 e;
  e.x = dead ? (V6_INSIDE | V6_SAFE)
             : in ? (V6_INSIDE | (uint32_t)(ry * V6_BOX_W + rx) | ((uint32_t)sx << 16) | ((uint32_t)sy << 17)) : 0u;
  e.y = dead ? 0u : wq.x;
  e.z = dead ? 0u : wq.y;
  e.w = (uint32_t)(base + cy0 * W + cx0) | ((uint32_t)sx << 30) | ((uint32_t)sy << 31);
  return e;
}

// Two 16-byte chunks (c_first, c_second: byte offsets inside the 128-byte pixel) of the modulated bilinear sample `e`.
__device__ __forceinline__ void v6_sample_main(const uint4& e, uint32_t box_main, const uint8_t* x_main, uint32_t main_row,
                                               uint32_t c_first, uint32_t c_second, uint4& F, uint4& S) {
  uint4 f[4], g[4];
  if (e.x & V6_INSIDE) {
    const uint32_t a00 = box_main + (e.x & 0xffffu) * (TC_CMAIN * 2);
    const uint32_t a01 = a00 + ((e.x & 0x10000u) ? TC_CMAIN * 2 : 0u);
    const uint32_t dy = (e.x & 0x20000u) ? V6_BOX_W * TC_CMAIN * 2 : 0u;
    f[0] = lds16(a00 + c_first); f[1] = lds16(a01 + c_first); f[2] = lds16(a00 + dy + c_first); f[3] = lds16(a01 + dy + c_first);
    g[0] = lds16(a00 + c_second); g[1] = lds16(a01 + c_second); g[2] = lds16(a00 + dy + c_second); g[3] = lds16(a01 + dy + c_second);
  } else {
    // rare: a corner of this sample lies outside the staged box -> global memory
    const uint8_t* a00 = x_main + (unsigned long long)(e.w & 0x3fffffffu) * (TC_CMAIN * 2);
    const uint8_t* a01 = a00 + ((e.w & 0x40000000u) ? TC_CMAIN * 2 : 0);
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

// The first four tail channels (8 bytes) of the modulated bilinear sample `e`.
__device__ __forceinline__ uint2 v6_sample_tail(const uint4& e, uint32_t box_tail, const uint8_t* x_tail, uint32_t tail_row) {
  uint2 v[4];
  if (e.x & V6_INSIDE) {
    const uint32_t a00 = box_tail + (e.x & 0xffffu) * (TC_CTAIL * 2);
    const uint32_t a01 = a00 + ((e.x & 0x10000u) ? TC_CTAIL * 2 : 0u);
    const uint32_t dy = (e.x & 0x20000u) ? V6_BOX_W * TC_CTAIL * 2 : 0u;
    v[0] = lds8(a00); v[1] = lds8(a01); v[2] = lds8(a00 + dy); v[3] = lds8(a01 + dy);
  } else {
    const uint8_t* a00 = x_tail + (unsigned long long)(e.w & 0x3fffffffu) * (TC_CTAIL * 2);
    const uint8_t* a01 = a00 + ((e.w & 0x40000000u) ? TC_CTAIL * 2 : 0);
    const uint32_t dy = (e.w & 0x80000000u) ? tail_row : 0u;
    v[0] = __ldg(reinterpret_cast<const uint2*>(a00)); v[1] = __ldg(reinterpret_cast<const uint2*>(a01));
    v[2] = __ldg(reinterpret_cast<const uint2*>(a00 + dy)); v[3] = __ldg(reinterpret_cast<const uint2*>(a01 + dy));
  }
  const uint4 r = lerp_chunk(make_uint4(v[0].x, v[0].y, 0u, 0u), make_uint4(v[1].x, v[1].y, 0u, 0u),
                             make_uint4(v[2].x, v[2].y, 0u, 0u), make_uint4(v[3].x, v[3].y, 0u, 0u), make_uint2(e.y, e.z));
  return make_uint2(r.x, r.y);
}

template <typename TO, typename TOUT>
__global__ void __launch_bounds__(V6_THREADS, 1) dcn_tc6_fwd_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  V6Smem& s = *reinterpret_cast<V6Smem*>(smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < V6_NA; ++i) {
      mbar_init(smem_u32(&s.a_full[i]), 4);                     // the four sub-partition warps of the producing group
      mbar_init(smem_u32(&s.a_empty[i]), 1);                    // one tcgen05.commit
    }
    for (int i = 0; i < V6_NB; ++i) {
      mbar_init(smem_u32(&s.b_full[i]), 1);                     // the loader's expect_tx arrival (+ the bytes)
      mbar_init(smem_u32(&s.b_empty[i]), 1);                    // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s.acc_full[i]), 1);                   // one tcgen05.commit
      mbar_init(smem_u32(&s.acc_empty[i]), V6_BACK_WARPS);      // the back-end warps (epilogue halves)
      mbar_init(smem_u32(&s.geo_full[i]), V6_BACK_WARPS);       // the back-end warps (geometry halves)
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

  if (warp < V6_PRODUCER_WARPS) {
    // =========================================================================== A-operand producers
    // Group gi (4 warps, one per TMEM sub-partition) produces the K blocks n = tile_iter * 10 + kb with n % 4 == gi into
    // A-ring stage n % 8; warp q of a group owns tile rows (= TMEM lanes) [32q, 32q + 32).
    const int group = warp >> 2, q = warp & 3;
    const int g = lane >> 2, u = lane & 3;
    const bool par = (g & 1) != 0;
    const uint32_t c_first = (uint32_t)(u + (par ? 4 : 0)) * 16, c_second = (uint32_t)(u + (par ? 0 : 4)) * 16;
    const uint32_t main_row = TC_CMAIN * 2 * (uint32_t)p.W, tail_row = TC_CTAIL * 2 * (uint32_t)p.W;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    for (int it = 0; it < my_tiles; ++it) {
      const int gb = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(smem_u32(&s.geo_full[gb]), tphase);             // this tile's geometry has been written
      mbar_wait(smem_u32(&s.box_full[gb]), tphase);             // this tile's source box has landed in shared memory
      const uint32_t box_main = smem_u32(&s.box_main[gb][0]), box_tail = smem_u32(&s.box_tail[gb][0]);
      const int n0 = it * V6_KBLOCKS;
      int kb = (group - n0) & 3;
      for (; kb < V6_KBLOCKS; kb += 4) {
        const int n = n0 + kb, sa = n % V6_NA;
        const uint32_t empty_bar = smem_u32(&s.a_empty[sa]), empty_par = (((uint32_t)(n / V6_NA)) & 1u) ^ 1u;
        const uint32_t a_taddr = tmem_base + lane_base + (uint32_t)(V6_A_COL0 + sa * 32);
        if (kb < 9) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint4 e0 = s.geo[gb][kb][q * 32 + h * 16 + g], e1 = s.geo[gb][kb][q * 32 + h * 16 + g + 8];
            uint4 F0, S0, F1, S1;
            v6_sample_main(e0, box_main, p.x_main, main_row, c_first, c_second, F0, S0);
            v6_sample_main(e1, box_main, p.x_main, main_row, c_first, c_second, F1, S1);
            // chunk u (X) and chunk u + 4 (Y) of both pixels: odd g loaded them in the opposite order
            const uint4 X0 = par ? S0 : F0, Y0 = par ? F0 : S0, X1 = par ? S1 : F1, Y1 = par ? F1 : S1;
            const uint32_t r[16] = {X0.x, X0.y, X1.x, X1.y, X0.z, X0.w, X1.z, X1.w,
                                    Y0.x, Y0.y, Y1.x, Y1.y, Y0.z, Y0.w, Y1.z, Y1.w};
            if (h == 0) {                                        // the ring stage is needed only now, after the gathers
              mbar_wait(empty_bar, empty_par);
              tc_fence_after();
            }
            tmem_st_16x256b_x4(a_taddr + ((uint32_t)(h * 16) << 16), r);
          }
        } else {
          // ---- tail block: lane = tile row, four channels of each of the nine taps (K = 36 of 48, rest zero)
          const int row = q * 32 + lane;
          uint32_t r[24];
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const uint2 v = v6_sample_tail(s.geo[gb][k][row], box_tail, p.x_tail, tail_row);
            r[2 * k] = v.x; r[2 * k + 1] = v.y;
          }
          r[18] = 0x3f803f80u;                               // K elements 36, 37 = 1.0: the weight image holds bias hi / lo there
#pragma unroll
          for (int i = 19; i < 24; ++i) r[i] = 0u;
          mbar_wait(empty_bar, empty_par);
          tc_fence_after();
          tmem_st_32x32b_x8(a_taddr, r);
          tmem_st_32x32b_x8(a_taddr + 8, r + 8);
          tmem_st_32x32b_x8(a_taddr + 16, r + 16);
        }
        tmem_st_wait();                                          // TMEM writes complete ...
        tc_fence_before();                                       // ... and ordered before the arrive the MMA lane waits on
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s.a_full[sa]));
      }
      __syncwarp();
      if (lane == 0) {                                           // this warp no longer reads geometry / box buffer gb
        mbar_arrive(smem_u32(&s.geo_empty[gb]));
        mbar_arrive(smem_u32(&s.box_empty[gb]));
      }
    }
  } else if (warp == V6_W_MMA) {
    // =========================================================================== MMA issuer (one lane)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TC_M, TC_N);
      uint32_t acc = 0, acc_phase[2] = {0, 0};
      int n = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(smem_u32(&s.acc_empty[acc]), acc_phase[acc] ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TC_ACC_STRIDE;
        for (int kb = 0; kb < V6_KBLOCKS; ++kb, ++n) {
          const int sa = n % V6_NA, sb = n % V6_NB;
          mbar_wait(smem_u32(&s.b_full[sb]), (uint32_t)(n / V6_NB) & 1u);
          mbar_wait(smem_u32(&s.a_full[sa]), (uint32_t)(n / V6_NA) & 1u);
          tc_fence_after();
          const uint32_t a_tmem = tmem_base + (uint32_t)(V6_A_COL0 + sa * 32);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(&s.b[sb][0]));
          const int nk = (kb == V6_KBLOCKS - 1) ? 3 : 4;
          for (int k = 0; k < nk; ++k)                   // 16 bf16 of K = 8 TMEM columns of A = 32 B of the B swizzle atom
            umma_bf16_ts(d_tmem, a_tmem + 8 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(smem_u32(&s.a_empty[sa]));
          umma_commit(smem_u32(&s.b_empty[sb]));
        }
        umma_commit(smem_u32(&s.acc_full[acc]));
        acc_phase[acc] ^= 1;
        acc ^= 1;
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
        mbar_wait(smem_u32(&s.b_empty[sb]), ((uint32_t)(n / V6_NB) & 1u) ^ 1u);
        const uint32_t bar = smem_u32(&s.b_full[sb]);
        mbar_arrive_expect_tx(bar, TC_B_BYTES);
        bulk_g2s(smem_u32(&s.b[sb][0]), p.wpacked + (size_t)kb * TC_B_BYTES, TC_B_BYTES, bar);
        if (++kb == V6_KBLOCKS) kb = 0;
      }
    }
    __syncwarp();
  } else if (warp == V6_W_BOX) {
    // =========================================================================== source-box copies (one lane per box row)
    for (int it = 0; it < my_tiles; ++it) {
      const int sb = it & 1;
      mbar_wait(smem_u32(&s.box_empty[sb]), ((uint32_t)(it >> 1) & 1u) ^ 1u);   // producers are done with the old box
      int b, ty0, tx0;
      tile_origin(p, tile0 + it * tile_step, b, ty0, tx0);
      const int by0 = ty0 - V6_BOX_TOP, bx0 = tx0 - V6_BOX_LEFT;
      const int ya = max(by0, 0), yb = min(by0 + V6_BOX_H, p.H), xa = max(bx0, 0), xb = min(bx0 + V6_BOX_W, p.W);
      const uint32_t ncol = (uint32_t)(xb - xa), nrow = (uint32_t)(yb - ya);
      const uint32_t bar = smem_u32(&s.box_full[sb]);
      if (lane == 0) mbar_arrive_expect_tx(bar, nrow * ncol * (TC_CMAIN + TC_CTAIL) * 2);
      __syncwarp();
      const int y = by0 + lane;
      if (lane < V6_BOX_H && y >= ya && y < yb) {
        const size_t gpix = (size_t)(b * p.H + y) * p.W + xa;
        const uint32_t bpix = (uint32_t)(lane * V6_BOX_W + (xa - bx0));
        bulk_g2s(smem_u32(&s.box_main[sb][0]) + bpix * (TC_CMAIN * 2), p.x_main + gpix * (TC_CMAIN * 2), ncol * TC_CMAIN * 2, bar);
        bulk_g2s(smem_u32(&s.box_tail[sb][0]) + bpix * (TC_CTAIL * 2), p.x_tail + gpix * (TC_CTAIL * 2), ncol * TC_CTAIL * 2, bar);
      }
      __syncwarp();
    }
  } else if (warp >= V6_W_BACK) {
    // =========================================================================== back end: geometry + epilogue (8 warps)
    // Two warps per TMEM quarter; `half` splits both jobs so that the per-tile dependent chain of one warp is short
    // enough to hide behind the producers: half 0 computes taps 0..4 and drains accumulator columns 0..31, half 1 taps
    // 5..8 and columns 32..79.
    const int quad = warp & 3;                           // TMEM lanes [32*quad, 32*quad + 32) belong to this warp
    const int half = (warp - V6_W_BACK) >> 2;
    const int row = quad * 32 + lane;                    // tile row = TMEM lane = geometry row of this thread
    const int etid = (warp - V6_W_BACK) * 32 + lane;     // 0..255 for the cooperative store
    uint32_t acc = 0, acc_phase[2] = {0, 0};
    const uint32_t ostage = smem_u32(&s.ostage[0]);

    auto make_geometry = [&](int it, auto half_tag) {
      constexpr int K0 = decltype(half_tag)::value ? 5 : 0, NK = decltype(half_tag)::value ? 4 : 5;
      const int gb = it & 1;
      mbar_wait(smem_u32(&s.geo_empty[gb]), ((uint32_t)(it >> 1) & 1u) ^ 1u);   // producers are done with the old contents
      int b, ty0, tx0;
      tile_origin(p, tile0 + it * tile_step, b, ty0, tx0);
      const int y = ty0 + row / TC_TW, x = tx0 + row % TC_TW;
      const int by0 = ty0 - V6_BOX_TOP, bx0 = tx0 - V6_BOX_LEFT;
      if (y < p.H && x < p.W) {
        // all offset / mask values of this thread are requested before the first one is used (one DRAM round trip)
        const int f_sc = (int)p.f_sc, m_sc = (int)p.m_sc;
        const TO* off = reinterpret_cast<const TO*>(p.offset) + b * p.f_sn + y * p.f_sh + x * p.f_sw;
        const TO* msk = reinterpret_cast<const TO*>(p.mask) + b * p.m_sn + y * p.m_sh + x * p.m_sw;
        const int base = b * p.H * p.W;
        TO rdy[NK], rdx[NK], rmk[NK];
#pragma unroll
        for (int i = 0; i < NK; ++i) {
          const int k = K0 + i, j0 = 2 * k, j1 = 2 * k + 1;
          if (p.fused27) {
            // ema_vfi.py:57-59 folded in: thirds 0 and 2 of the 27 channels are the offsets (tap k uses channels 2k and
            // 2k+1 of their concatenation), the middle third is the pre-sigmoid mask
            rdy[i] = __ldg(off + (j0 < 9 ? j0 : j0 + 9) * f_sc);
            rdx[i] = __ldg(off + (j1 < 9 ? j1 : j1 + 9) * f_sc);
            rmk[i] = __ldg(msk + (9 + k) * m_sc);
          } else {
            rdy[i] = __ldg(off + j0 * f_sc);
            rdx[i] = __ldg(off + j1 * f_sc);
            rmk[i] = __ldg(msk + k * m_sc);
          }
        }
#pragma unroll
        for (int i = 0; i < NK; ++i) {
          float mk = to_f32<TO>(rmk[i]);
          // the sigmoid result is rounded to the tensor dtype, as torch.sigmoid on that tensor would
          if (p.fused27) mk = to_f32<TO>(from_f32<TO>(1.0f / (1.0f + __expf(-mk))));
          s.geo[gb][K0 + i][row] =
              v6_geo_entry(p.H, p.W, base, by0, bx0, y, x, K0 + i, to_f32<TO>(rdy[i]), to_f32<TO>(rdx[i]), mk);
        }
      } else {
#pragma unroll
        for (int i = 0; i < NK; ++i) s.geo[gb][K0 + i][row] = make_uint4(V6_INSIDE | V6_SAFE, 0u, 0u, 0u);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s.geo_full[gb]));
    };
    auto geometry = [&](int it) {
      if (half) make_geometry(it, std::integral_constant<int, 1>{});
      else make_geometry(it, std::integral_constant<int, 0>{});
    };

    if (my_tiles > 0) geometry(0);
    for (int it = 0; it < my_tiles; ++it) {
      if (it + 1 < my_tiles) geometry(it + 1);           // overlaps the producers' work on tile `it`
      int b, ty0, tx0;
      tile_origin(p, tile0 + it * tile_step, b, ty0, tx0);
      mbar_wait(smem_u32(&s.acc_full[acc]), acc_phase[acc]);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TC_ACC_STRIDE;
      const int y = ty0 + row / TC_TW, x = tx0 + row % TC_TW;
      const bool inside = y < p.H && x < p.W;
      const size_t pixel = (size_t)(b * p.H + y) * p.W + x;
      __nv_bfloat16* ot = reinterpret_cast<__nv_bfloat16*>(p.out_tail) + pixel * TC_CTAIL;
      TOUT* os = reinterpret_cast<TOUT*>(p.out) + b * p.o_sn + y * p.o_sh + x * p.o_sw;
      const int c16_lo = half ? 2 : 0, c16_hi = half ? TC_N / 16 : 2;
      for (int c16 = c16_lo; c16 < c16_hi; ++c16) {
        uint32_t d[16];
        tmem_ld16(taddr + c16 * 16, d);
        tmem_ld_wait();
        if (c16 == c16_hi - 1) {                         // last TMEM read of this accumulator by this warp: hand it back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&s.acc_empty[acc]));
        }
        // the bias is already in the accumulator (K elements 36/37 of the tail block)
        if (p.out_tail) {
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
            else if (inside) *reinterpret_cast<uint4*>(ot) = w4;     // 16 B records of neighbouring pixels coalesce
          }
        } else if (inside) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c16 * 16 + i;
            if (c < p.O) os[c * p.o_sc] = from_f32<TOUT>(__uint_as_float(d[i]));
          }
        }
      }
      if (p.out_tail) {
        // main plane: the staged tile leaves as full 128-byte lines (8 lanes per pixel)
        epi_bar_sync();
        uint8_t* om = reinterpret_cast<uint8_t*>(p.out);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = i * 256 + etid, r = idx >> 3, c = idx & 7;
          const int yy = ty0 + r / TC_TW, xx = tx0 + r % TC_TW;
          const uint4 v = lds16(ostage + (uint32_t)r * 128 + ((uint32_t)(c ^ (r & 7)) << 4));
          if (yy < p.H && xx < p.W)
            *reinterpret_cast<uint4*>(om + ((size_t)(b * p.H + yy) * p.W + xx) * (TC_CMAIN * 2) + c * 16) = v;
        }
        epi_bar_sync();                                   // the staging tile may be overwritten by the next tile
      }
      acc_phase[acc] ^= 1;
      acc ^= 1;
    }
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
  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(TC_M, TC_N);
    const uint64_t bdesc = umma_desc_sw128(smem_u32(sb));
    for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_base, tmem_base + a_col + 8 * k, bdesc + 2 * k, idesc, k != 0);
    for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_base, tmem_base + a_col + 32 + 8 * k, bdesc + 2 * k, idesc, 1);
    umma_commit(smem_u32(bar));
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
