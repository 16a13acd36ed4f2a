// dcn_gcol.cuh -- the dense half of the DCNv2 data backward on tcgen05.  Included by dcn_tc.cu inside namespace vfi::<anonymous>.
//
// torchvision::_deform_conv2d_backward (reached from /root/reference/train.py:125 through ema_vfi.py:60) starts with the column
// gradient  gcol[p, (k, c)] = sum_o gout[p, o] * W[o, c, k]  -- a plain GEMM [P, 67] x [67, 603] with no gather in it.  Round 1
// made it with cuBLAS (torch.matmul) around eager pack / zero-fill copies; this kernel replaces that on the tensor-core training
// path:
//
//     D_k[128 px (TMEM lanes), 80 c] (fp32)  =  A[128 px, 80 o] (TMEM, bf16)  x  B_k[80 c, 80 o]^T (smem, SWIZZLE_128B, K-major)
//
//   * A = the grad_out tile, read where it lies (any strides; rounded to bf16) by four loader warps -- thread = pixel = TMEM
//     lane -- and written with tcgen05.st.32x32b (double buffered across tiles).  No packed copy of grad_out, no zero fill.
//   * B = W transposed per tap, RESIDENT in shared memory for the whole (persistent) kernel: nine 64-K atoms (o = 0..63) and
//     three tail atoms holding the K = 16 slices o = 64..79 of four taps each -- 12 x 10 KB, loaded once per CTA.  Five
//     tcgen05.mma (M128 N80 K16) per tap, accumulators double buffered in TMEM.
//   * Epilogue: tcgen05.ld (thread = pixel, 72 channels of one tap) -> bf16 -> a [128][72] staging tile -> ONE TMA store per
//     (tile, tap) into gcol[P][648] (clipped at the last tile), double buffered so the store overlaps the next tap.
//
// The consumer (dcn_bwd_cols_kernel) is unchanged; folding it in behind this front half is the next step (gcol would then
// never reach HBM).

constexpr int GC_TAPS = 9, GC_ATOMS = 12;                      // 9 main atoms + 3 tail atoms (taps 0-3 / 4-7 / 8)
constexpr int GC_LD = 72;                                      // gcol columns per tap
constexpr int GC_W_LOAD = 0, GC_W_MMA = 4, GC_W_EPI = 5;       // warps 0-3 loaders, 4 MMA + weight load, 5-8 epilogue
constexpr int GC_THREADS = 9 * 32;
constexpr int GC_A_COL0 = 256, GC_A_STRIDE = 64;               // TMEM: D at 0 / 128, A (40 columns) at 256 / 320

struct __align__(1024) GcSmem {
  uint8_t b[GC_ATOMS][TC_B_BYTES];                             // 122,880
  uint8_t stage[2][TC_M * GC_LD * 2];                          // 2 x 18,432: [128 px][72] bf16, TMA store source
  unsigned long long b_full, a_full[2], a_empty[2], d_full[2], d_empty[2];
  uint32_t tmem_base;
};

struct GcArgs {
  const void* gout; long long g_sn, g_sc, g_sh, g_sw;          // grad_out [B,O,H,W], any strides
  const uint8_t* wimg;                                         // [12][80][128 B] swizzled (pack_gcol_weight_kernel)
  int B, H, W, O;
  long long P;
  int num_tiles;
  alignas(64) CUtensorMap tm_gcol;                             // {648, P} bf16, box {72, 128}
};

// Weight image of the column-gradient GEMM: atom a < 9 = tap a, row c, K element kk = output channel o = kk (0..63);
// atom 9 + g: row c, K element kk = slice kk / 16 (tap 4 g + kk / 16) of o = 64 + kk % 16.  Zero beyond C / O / tap 8.
template <typename TW>
__global__ void pack_gcol_weight_kernel(const TW* __restrict__ w, int O, int C, uint8_t* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= GC_ATOMS * TC_N * 64) return;
  const int a = idx / (TC_N * 64), c = (idx / 64) % TC_N, kk = idx % 64;
  int tap, o;
  if (a < GC_TAPS) { tap = a; o = kk; }
  else { tap = 4 * (a - GC_TAPS) + (kk >> 4); o = 64 + (kk & 15); }
  float v = 0.0f;
  if (tap < GC_TAPS && o < O && c < C) v = to_f32<TW>(w[((size_t)o * C + c) * 9 + tap]);
  const size_t off = (size_t)a * TC_B_BYTES + (size_t)c * 128 + ((((kk >> 3) ^ (c & 7))) << 4) + (kk & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(v);
}

template <typename TG>
__global__ void __launch_bounds__(GC_THREADS, 1) dcn_gcol_gemm_kernel(const __grid_constant__ GcArgs q) {
  extern __shared__ uint8_t smem_raw[];
  GcSmem& s = *reinterpret_cast<GcSmem*>(smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023));
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    mbar_init(smem_u32(&s.b_full), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s.a_full[i]), 4);
      mbar_init(smem_u32(&s.a_empty[i]), 1);
      mbar_init(smem_u32(&s.d_full[i]), 1);
      mbar_init(smem_u32(&s.d_empty[i]), 4);
    }
    fence_barrier_init();
  }
  if (warp == GC_W_MMA) {
    tmem_alloc(smem_u32(&s.tmem_base), 512);
    if (lane == 0) tma_prefetch_desc(&q.tm_gcol);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;
  const int my_tiles = (q.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile0 = (int)blockIdx.x, tile_step = (int)gridDim.x;

  if (warp < GC_W_MMA) {
    // =========================================================================== grad_out tile -> TMEM (thread = pixel)
    const int quad = warp & 3, row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const long long HW = (long long)q.H * q.W;
    for (int it = 0; it < my_tiles; ++it) {
      const int ab = it & 1;
      const long long pp = (long long)(tile0 + it * tile_step) * TC_M + row;
      uint32_t r[40];
#pragma unroll
      for (int j = 0; j < 40; ++j) r[j] = 0u;
      if (pp < q.P) {
        const long long b = pp / HW, rem = pp - b * HW;
        const int y = (int)(rem / q.W), x = (int)(rem - (long long)y * q.W);
        const TG* src = reinterpret_cast<const TG*>(q.gout) + b * q.g_sn + (long long)y * q.g_sh + (long long)x * q.g_sw;
#pragma unroll
        for (int j = 0; j < 34; ++j) {                            // 68 channel slots: o = 2 j, 2 j + 1
          const float f0 = 2 * j < q.O ? to_f32<TG>(__ldg(src + (long long)(2 * j) * q.g_sc)) : 0.0f;
          const float f1 = 2 * j + 1 < q.O ? to_f32<TG>(__ldg(src + (long long)(2 * j + 1) * q.g_sc)) : 0.0f;
          const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
          r[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
      }
      mbar_wait_ns<V6_NS_HELP>(smem_u32(&s.a_empty[ab]), (((uint32_t)it >> 1) & 1u) ^ 1u);     // the MMAs of tile it - 2 are done
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + (uint32_t)(GC_A_COL0 + ab * GC_A_STRIDE);
#pragma unroll
      for (int j = 0; j < 5; ++j) tmem_st_32x32b_x8(taddr + 8 * j, r + 8 * j);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s.a_full[ab]));
    }
  } else if (warp == GC_W_MMA) {
    // =========================================================================== weights (once), then the MMA issue loop
    if (lane == 0) {
      mbar_arrive_expect_tx(smem_u32(&s.b_full), (uint32_t)GC_ATOMS * TC_B_BYTES);
      for (int a = 0; a < GC_ATOMS; ++a) bulk_g2s(smem_u32(&s.b[a][0]), q.wimg + (size_t)a * TC_B_BYTES, TC_B_BYTES, smem_u32(&s.b_full));
    }
    __syncwarp();
    mbar_wait_ns<64>(smem_u32(&s.b_full), 0u);
    constexpr uint32_t idesc = umma_idesc_bf16(TC_M, TC_N);
    const uint32_t b_smem = smem_u32(&s.b[0][0]);
    uint32_t nd = 0;                                              // accumulator uses so far
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t ab = (uint32_t)it & 1u;
      mbar_wait_ns<V6_NS_MMA>(smem_u32(&s.a_full[ab]), ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      const uint32_t a_tmem = tmem_base + (uint32_t)(GC_A_COL0 + ab * GC_A_STRIDE);
#pragma unroll 1
      for (int k = 0; k < GC_TAPS; ++k, ++nd) {
        const uint32_t db = nd & 1u;
        mbar_wait_ns<V6_NS_MMA>(smem_u32(&s.d_empty[db]), ((nd >> 1) & 1u) ^ 1u);              // the epilogue has read this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + db * TC_ACC_STRIDE;
        const uint64_t bmain = umma_desc_sw128(b_smem + (uint32_t)k * TC_B_BYTES);
        const uint64_t btail = umma_desc_sw128(b_smem + (uint32_t)(GC_TAPS + (k >> 2)) * TC_B_BYTES) + (uint64_t)(2 * (k & 3));
        umma_bf16_ts(d_tmem, a_tmem, bmain, idesc, 0);
        umma_bf16_ts(d_tmem, a_tmem + 8, bmain + 2, idesc, 1);
        umma_bf16_ts(d_tmem, a_tmem + 16, bmain + 4, idesc, 1);
        umma_bf16_ts(d_tmem, a_tmem + 24, bmain + 6, idesc, 1);
        umma_bf16_ts(d_tmem, a_tmem + 32, btail, idesc, 1);                                    // o = 64..79
        umma_commit_elect(smem_u32(&s.d_full[db]));
      }
      umma_commit_elect(smem_u32(&s.a_empty[ab]));
    }
    __syncwarp();
  } else {
    // =========================================================================== epilogue: TMEM -> bf16 -> staging -> TMA store
    const int quad = warp & 3, row = quad * 32 + lane;
    const int etid = (warp - GC_W_EPI) * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint32_t nd = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int p0 = (tile0 + it * tile_step) * TC_M;
      for (int k = 0; k < GC_TAPS; ++k, ++nd) {
        const uint32_t db = nd & 1u;
        mbar_wait_ns<V6_NS_HELP>(smem_u32(&s.d_full[db]), (nd >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_base + db * TC_ACC_STRIDE;
        const uint32_t st = smem_u32(&s.stage[db][0]) + (uint32_t)row * (GC_LD * 2);
        // staging buffer db was the source of the TMA store issued two taps ago: thread 0 has waited for it (below) before
        // the barrier that ended the previous tap
#pragma unroll
        for (int c16 = 0; c16 < 5; ++c16) {
          uint32_t d[16];
          tmem_ld16(taddr + c16 * 16, d);
          tmem_ld_wait();
          if (c16 == 4) {                                  // last read of this accumulator: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s.d_empty[db]));
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c0 = c16 * 16 + h * 8;
            if (c0 >= GC_LD) break;
            uint4 w4;
            uint32_t* w = reinterpret_cast<uint32_t*>(&w4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __nv_bfloat162 hv = __floats2bfloat162_rn(__uint_as_float(d[h * 8 + 2 * i]), __uint_as_float(d[h * 8 + 2 * i + 1]));
              w[i] = *reinterpret_cast<const uint32_t*>(&hv);
            }
            sts16(st + (uint32_t)c0 * 2, w4);
          }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) {
          tma_store_4d(&q.tm_gcol, k * GC_LD, p0, 0, 0, smem_u32(&s.stage[db][0]));
          tma_store_commit();
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the store from the OTHER buffer has been read
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
    if (etid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == GC_W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
