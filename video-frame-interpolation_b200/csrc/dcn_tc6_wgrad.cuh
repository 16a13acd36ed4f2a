// dcn_tc6_wgrad.cuh -- weight / bias gradient of the DCNv2 layer on tcgen05 (bf16 operands, fp32 accumulate in TMEM).
// Included by dcn_tc.cu after dcn_tc6.cuh, inside namespace vfi::<anonymous>.
//
// Replaces the grad_weight / grad_bias half of torchvision::_deform_conv2d_backward (reached from
// /root/reference/train.py:125 through ema_vfi.py:60) on the tensor-core path:
//
//     gW[o, c, k] = sum_p gout[p, o] * col[p, (k, c)]          col = the modulated bilinear samples the forward gathers
//
// as one GEMM per tap whose reduction dimension is the PIXEL:
//
//     D_k[128 o (TMEM lanes), 64 c] (fp32)  +=  A[128 o, 128 px] (TMEM: gout^T of the tile)  x  B_k[128 px, 64 c] (smem)
//
//   * B_k is what the producers compute anyway (one 128-byte row of 64 channels per pixel, 16-byte chunk j of row r at
//     j ^ (r & 7)).  Read as a B operand whose N (= channel) dimension is contiguous this is exactly the canonical
//     SWIZZLE_128B MN-major atom ((8,n),(8,k)):((1,LBO),(8,SBO)) with SBO = 1024 B: b_major (bit 16 of the instruction
//     descriptor) = 1, and a UMMA_K = 16 step advances the start address by two 8-row groups (2048 B).  No transpose.
//   * A = gout^T: lane = output channel o, 32-bit column = a pixel pair of the tile; four loader warps (one per TMEM
//     quarter) read 8 x 32 contiguous bytes per channel and write them with tcgen05.st.32x32b, double buffered.
//   * The accumulators PERSIST in TMEM over all tiles of the CTA and leave once, by fp32 atomics: 5 blocks x 64 columns.
//     Nine taps + the tail channels do not fit 512 columns next to A, so the layer takes two launches: pass 0 = taps 0..4,
//     pass 1 = taps 5..8 + the tail block (four tail channels of each of the nine taps, N = 48).
//   * Box staging, geometry, cp.async of the offsets: the forward's roles, unchanged (v6_role_box / v6_role_geometry).
//     Five producer groups, one block each per tile.

constexpr int WG_STAGES = 3;                                   // col stages in shared memory (16 KB each)
constexpr int WG_NDONE = 12;                                   // completion barriers: 24 blocks (~5 tiles) of aliasing distance
constexpr int WG_BLOCKS = 5;                                   // blocks (taps, or the tail block) per tile and pass
constexpr int WG_GOUT_COL0 = WG_BLOCKS * 64;                   // TMEM columns [0, 320): accumulators; [320, 448): gout^T x 2
constexpr int WG_W_MMA = 20, WG_W_BOX = 21;                    // warps 22, 23 idle
constexpr int WG_W_GOUT = 24;                                  // warps 24..27: gout^T loaders, then the final epilogue
constexpr int WG_W_GEO = 28;                                   // warps 28..31: tap geometry

struct __align__(1024) WgSmem {
  uint8_t stage[WG_STAGES][TC_A_BYTES];                        // col_k tile: 128 pixel rows x 128 B (MN-major B operand)
  uint8_t box_main[2][V6_BOX_PX * V6_MAIN_PX];
  uint8_t box_tail[2][V6_BOX_PX * V6_TAIL_PX];
  uint4 geo[2][9][TC_M];
  uint16_t raw[27][TC_M];
  unsigned long long stage_full[WG_STAGES], stage_empty[WG_NDONE];   // "block n consumed": ring of its own, see V6_NDONE
  unsigned long long box_full[2], box_empty[2], geo_first[2], geo_full[2], geo_empty[2], gout_full[2], gout_empty[2], acc_done;
  uint32_t tmem_base;
};

// Shared-memory descriptor of an MN-major SWIZZLE_128B operand one atom (64 elements) wide: 8-row K groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);       // start address
  d |= (uint64_t)(1024 >> 4) << 16;              // leading byte offset: stride between 64-element atoms along MN (one atom: unused)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: between 8-row groups along K
  d |= (uint64_t)1 << 46;                        // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

struct WgParams {
  TcParams t;                                    // planes, offsets / mask, geometry of the tiling (out / weights unused)
  const void* gout; long long g_sn, g_sc, g_sh, g_sw;  // grad_out [B,O,H,W], any strides (NCHW or channels_last)
  int g_vec;                                         // 1: 16-bit grad_out with unit pixel stride and 16-byte aligned rows
  float* gw; float* gb;                          // [O,C,3,3] / [O] fp32, accumulated into
  int C, pass;
};

// TO: 16-bit dtype of the offset / mask tensors; TG: dtype of grad_out (bf16 or f32, rounded to bf16 for the tensor core).
template <typename TO, typename TG, bool FUSED27>
__global__ void __launch_bounds__(V6_THREADS, 1) dcn_tc6_wgrad_kernel(const WgParams q) {
  const TcParams& p = q.t;
  extern __shared__ uint8_t smem_raw[];
  WgSmem& s = *reinterpret_cast<WgSmem*>(smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < WG_STAGES; ++i) mbar_init(smem_u32(&s.stage_full[i]), 4);    // the four warps of the producing group
    for (int i = 0; i < WG_NDONE; ++i) mbar_init(smem_u32(&s.stage_empty[i]), 1);    // one tcgen05.commit
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s.box_full[i]), 1);
      mbar_init(smem_u32(&s.box_empty[i]), V6_PRODUCER_WARPS);
      mbar_init(smem_u32(&s.geo_first[i]), V6_GEO_WARPS);
      mbar_init(smem_u32(&s.geo_full[i]), V6_GEO_WARPS);
      mbar_init(smem_u32(&s.geo_empty[i]), V6_PRODUCER_WARPS);
      mbar_init(smem_u32(&s.gout_full[i]), 4);                  // four loader warps
      mbar_init(smem_u32(&s.gout_empty[i]), 1);                 // one tcgen05.commit
    }
    mbar_init(smem_u32(&s.acc_done), 1);
    fence_barrier_init();
  }
  if (warp == WG_W_MMA) tmem_alloc(smem_u32(&s.tmem_base), V6_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;
  const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile0 = (int)blockIdx.x, tile_step = (int)gridDim.x;
  long long w0 = 0, w2 = 0;
  const int tap0 = q.pass * WG_BLOCKS;                          // pass 0: taps 0..4; pass 1: taps 5..8, then the tail block

  if (warp < V6_PRODUCER_WARPS) {
    // =========================================================================== col producers
    // Group g produces block g of every tile into stage (5 it + g) % 3; warp w of a group owns tile rows [32w, 32w + 32).
    // Main block: 8 lanes per pixel row, lane j owns 16-byte chunk j (8 channels) of all four corners.
    const int group = warp >> 2, wq = warp & 3;
    const int rsub = lane >> 3, j = lane & 7;
    const int tap = tap0 + group;                               // 9 = the tail block
    const uint32_t main_row = V6_MAIN_PX * (uint32_t)p.W, tail_row = V6_TAIL_PX * (uint32_t)p.W;
    for (int it = 0; it < my_tiles; ++it) {
      const int gb = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait_ns<V6_NS_PROD>(smem_u32(&s.geo_full[gb]), tphase);
      mbar_wait_ns<V6_NS_PROD>(smem_u32(&s.box_full[gb]), tphase);
      const uint32_t box_main = smem_u32(&s.box_main[gb][0]) + (uint32_t)j * 16, box_tail = smem_u32(&s.box_tail[gb][0]);
      const int n = it * WG_BLOCKS + group, st = n % WG_STAGES;
      const uint32_t stage = smem_u32(&s.stage[st][0]);
      if (n >= WG_STAGES) {                                     // the stage was last used by block n - 3: wait until it is consumed
        const int m = n - WG_STAGES;
        mbar_wait_ns<V6_NS_STAGE>(smem_u32(&s.stage_empty[m % WG_NDONE]), (uint32_t)(m / WG_NDONE) & 1u);
      }
      if (tap < 9) {
#pragma unroll 2
        for (int i = 0; i < 8; ++i) {
          const int row = wq * 32 + i * 4 + rsub;
          const uint4 e = s.geo[gb][tap][row];
          uint4 v[4];
          if ((int)e.x >= 0) {
            const uint32_t a = box_main + e.x;
            v[0] = lds16o<0>(a); v[1] = lds16o<V6_MAIN_PX>(a); v[2] = lds16o<V6_MAIN_ROW>(a); v[3] = lds16o<V6_MAIN_ROW + V6_MAIN_PX>(a);
          } else {
            const uint8_t* a00 = p.x_main + (unsigned long long)(e.w & 0x3fffffffu) * V6_MAIN_PX + j * 16;
            const uint8_t* a01 = a00 + ((e.w & 0x40000000u) ? V6_MAIN_PX : 0);
            const uint32_t dy = (e.w & 0x80000000u) ? main_row : 0u;
            v[0] = __ldg(reinterpret_cast<const uint4*>(a00)); v[1] = __ldg(reinterpret_cast<const uint4*>(a01));
            v[2] = __ldg(reinterpret_cast<const uint4*>(a00 + dy)); v[3] = __ldg(reinterpret_cast<const uint4*>(a01 + dy));
          }
          sts16(stage + (uint32_t)row * 128 + ((uint32_t)(j ^ (row & 7)) << 4), lerp_chunk(v[0], v[1], v[2], v[3], make_uint2(e.y, e.z)));
        }
      } else {
        // tail block: lane = tile row; tap k's four tail channels at bytes [8k, 8k + 8) of the row, zeros up to byte 96
        const int row = wq * 32 + lane;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
          if (2 * c < 9) lo = v6_sample_tail(s.geo[gb][2 * c][row], box_tail, p.x_tail, tail_row, (uint32_t)(lane & 1) * 8u);
          if (2 * c + 1 < 9) hi = v6_sample_tail(s.geo[gb][2 * c + 1][row], box_tail, p.x_tail, tail_row, (uint32_t)(lane & 1) * 8u);
          sts16(stage + (uint32_t)row * 128 + ((uint32_t)(c ^ (row & 7)) << 4), make_uint4(lo.x, lo.y, hi.x, hi.y));
        }
      }
      fence_proxy_async();                                      // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&s.stage_full[st]));
        mbar_arrive(smem_u32(&s.geo_empty[gb]));                // this warp no longer reads geometry / box buffer gb
        mbar_arrive(smem_u32(&s.box_empty[gb]));
      }
    }
  } else if (warp == WG_W_MMA) {
    // =========================================================================== MMA issuer (whole warp, one elected lane)
    constexpr uint32_t idesc64 = umma_idesc_bf16(TC_M, 64) | (1u << 16);     // B is MN-major
    constexpr uint32_t idesc48 = umma_idesc_bf16(TC_M, 48) | (1u << 16);
    const uint32_t stage0 = smem_u32(&s.stage[0][0]);
    int n = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t gbuf = (uint32_t)it & 1u;
      mbar_wait_ns<V6_NS_MMA>(smem_u32(&s.gout_full[gbuf]), ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      const uint32_t a_tmem = tmem_base + (uint32_t)(WG_GOUT_COL0 + gbuf * 64);
#pragma unroll 1
      for (int blk = 0; blk < WG_BLOCKS; ++blk, ++n) {
        const int st = n % WG_STAGES;
        mbar_wait_ns<V6_NS_MMA>(smem_u32(&s.stage_full[st]), (uint32_t)(n / WG_STAGES) & 1u);
        tc_fence_after();
        const uint64_t bdesc = umma_desc_sw128_mn(stage0 + (uint32_t)st * TC_A_BYTES);
        const uint32_t d_tmem = tmem_base + (uint32_t)(blk * 64);
        const uint32_t idesc = (tap0 + blk == 9) ? idesc48 : idesc64;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)                          // 16 pixels = 8 TMEM columns of A = two 8-row groups of B
          umma_bf16_ts(d_tmem, a_tmem + 8 * kk, bdesc + (uint64_t)(kk * (2048 >> 4)), idesc, (it | kk) != 0);
        umma_commit_elect(smem_u32(&s.stage_empty[n % WG_NDONE]));
      }
      umma_commit_elect(smem_u32(&s.gout_empty[gbuf]));
    }
    umma_commit_elect(smem_u32(&s.acc_done));
    __syncwarp();
  } else if (warp == WG_W_BOX) {
    v6_role_box<false>(s, p, lane, my_tiles, tile0, tile_step, w0);
  } else if (warp >= WG_W_GEO) {
    v6_role_geometry<TO, FUSED27, false>(s, p, (warp - WG_W_GEO) * 32 + lane, lane, my_tiles, tile0, tile_step, w0, w2);
  } else if (warp >= WG_W_GOUT) {
    // =========================================================================== gout^T loaders, then the epilogue
    const int quad = warp & 3, o = quad * 32 + lane;           // TMEM lane = output channel
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const bool live = o < p.O;
    float bsum = 0.0f;
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t gbuf = (uint32_t)it & 1u;
      int b, ty0, tx0;
      tile_origin(p, tile0 + it * tile_step, b, ty0, tx0);
      const TG* src = reinterpret_cast<const TG*>(q.gout) + b * q.g_sn + (long long)o * q.g_sc + (long long)ty0 * q.g_sh + (long long)tx0 * q.g_sw;
      const int rows = min(TC_TH, p.H - ty0), cols = min(TC_TW, p.W - tx0);
      mbar_wait_ns<V6_NS_HELP>(smem_u32(&s.gout_empty[gbuf]), (((uint32_t)it >> 1) & 1u) ^ 1u);   // the MMAs of tile it - 2 are done
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + (uint32_t)(WG_GOUT_COL0 + gbuf * 64);
#pragma unroll
      for (int y4 = 0; y4 < TC_TH; y4 += 4) {                   // four tile rows (32 registers) in flight at a time
        uint32_t r[4][8];
#pragma unroll
        for (int yy = 0; yy < 4; ++yy) {
          const int y = y4 + yy;
          const TG* row = src + (long long)y * q.g_sh;
          if (live && y < rows) {
            if (sizeof(TG) == 2 && q.g_vec) {                   // 16 pixels = 32 contiguous bytes (W % 8 == 0: cols is 8 or 16)
              const uint4 lo = __ldg(reinterpret_cast<const uint4*>(row));
              const uint4 hi = cols > 8 ? __ldg(reinterpret_cast<const uint4*>(row + 8)) : make_uint4(0u, 0u, 0u, 0u);
              r[yy][0] = lo.x; r[yy][1] = lo.y; r[yy][2] = lo.z; r[yy][3] = lo.w;
              r[yy][4] = hi.x; r[yy][5] = hi.y; r[yy][6] = hi.z; r[yy][7] = hi.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                // channels_last grad_out lands here: lanes are consecutive channels, so a warp reads 64 contiguous bytes
                const float f0 = 2 * i < cols ? to_f32<TG>(__ldg(row + (2 * i) * q.g_sw)) : 0.0f;
                const float f1 = 2 * i + 1 < cols ? to_f32<TG>(__ldg(row + (2 * i + 1) * q.g_sw)) : 0.0f;
                const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
                r[yy][i] = *reinterpret_cast<const uint32_t*>(&h);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[yy][i] = 0u;
          }
        }
#pragma unroll
        for (int yy = 0; yy < 4; ++yy) {
#pragma unroll
          for (int i = 0; i < 8; ++i) bsum += __uint_as_float(r[yy][i] << 16) + __uint_as_float(r[yy][i] & 0xffff0000u);
          tmem_st_32x32b_x8(taddr + 8 * (y4 + yy), r[yy]);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s.gout_full[gbuf]));
    }
    // ---- epilogue: the accumulators leave TMEM once, by fp32 atomics (grad_weight may live in a flat gradient bucket)
    mbar_wait_ns<V6_NS_HELP>(smem_u32(&s.acc_done), 0u);
    tc_fence_after();
    if (my_tiles > 0) {
      for (int blk = 0; blk < WG_BLOCKS; ++blk) {
        const int tap = tap0 + blk;
#pragma unroll 1
        for (int c16 = 0; c16 < 4; ++c16) {
          if (tap == 9 && c16 == 3) break;                      // the tail block has 48 columns
          uint32_t d[16];
          tmem_ld16(tmem_base + lane_base + (uint32_t)(blk * 64 + c16 * 16), d);
          tmem_ld_wait();
          if (!live || !q.gw) continue;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = c16 * 16 + i;
            int k = tap, c = col;
            if (tap == 9) { k = col >> 2; c = TC_CMAIN + (col & 3); if (k >= 9) continue; }
            if (c < q.C) atomicAdd(q.gw + ((size_t)o * q.C + c) * 9 + k, __uint_as_float(d[i]));
          }
        }
      }
      if (q.pass == 0 && q.gb && live) atomicAdd(q.gb + o, bsum);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WG_W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, V6_TMEM_COLS);
  }
}
