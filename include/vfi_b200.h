/*
 * vfi_b200.h -- C ABI of libvfi_b200.so: the B200 (sm_100a) implementation of the warp + DeformConv2d hot path
 * of 424635328/video-frame-interpolation.
 *
 * Every entry point replaces one edge of the reference's L2->L1 boundary.  The reference
 * reaches these ops through Python, so "the reference's FFI for this path" is a ctypes binding; the one shipped in
 * video-frame-interpolation_b200/_lib.py is the stub a reference maintainer would add.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers owned by the caller (in practice the PyTorch caching allocator),
 *     including the workspace.  The library allocates nothing on the data path and keeps no reference to any buffer.
 *   - Work is enqueued on the stream passed in; no call synchronises the device or the stream.
 *   - Every function returns a vfi_status; 0 is success.  Nothing throws across the boundary.  A human-readable
 *     description of the last failure on the calling thread is available from vfi_last_error().
 *   - Tensors are described by vfi_tensor: logical NCHW extents plus element strides, so NCHW-contiguous,
 *     channels-last and channel-padded channels-last buffers all go through the same call.  The kernels pick a
 *     vectorised / TMA / tcgen05 specialisation when the strides allow and a strided path otherwise.
 *   - There is NO CPU fallback and no multi-architecture dispatch: the device must be compute capability 10.x.
 */
#ifndef VFI_B200_H_
#define VFI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFI_B200_ABI_VERSION 1

typedef void* vfi_stream_t; /* cudaStream_t */

typedef enum vfi_status {
  VFI_OK = 0,
  VFI_ERR_INVALID = 1,      /* bad argument: null pointer, shape mismatch, negative extent ...           */
  VFI_ERR_UNSUPPORTED = 2,  /* well-formed but outside the geometry this library implements               */
  VFI_ERR_CUDA = 3,         /* a CUDA runtime / driver call failed; message carries cudaGetErrorString     */
  VFI_ERR_WORKSPACE = 4,    /* workspace missing or smaller than vfi_dcn_workspace_bytes() asked for       */
  VFI_ERR_DEVICE = 5        /* current device is not sm_100 (B200)                                         */
} vfi_status;

typedef enum vfi_dtype { VFI_F32 = 0, VFI_BF16 = 1, VFI_F16 = 2 } vfi_dtype;

typedef struct vfi_tensor {
  void* data;            /* device pointer to element (0,0,0,0)            */
  int32_t dtype;         /* vfi_dtype                                      */
  int32_t reserved;      /* must be 0                                      */
  int64_t n, c, h, w;    /* logical extents                                */
  int64_t sn, sc, sh, sw;/* strides in elements                            */
} vfi_tensor;

/* Arithmetic used by the DCN contraction. */
typedef enum vfi_dcn_math {
  VFI_DCN_MATH_AUTO = 0,  /* fp32 tensors -> FP32, bf16/f16 tensors -> BF16_TC                                      */
  VFI_DCN_MATH_FP32 = 1,  /* fp32 gather + FFMA contraction: the parity mode (max-abs 1e-5 vs torchvision fp32)     */
  VFI_DCN_MATH_BF16_TC = 2,/* bf16 operands, fp32 accumulate in TMEM via tcgen05.mma (C <= 72, O <= 80); the bilinear  */
                          /* blend of the four corners runs in packed bf16 FMAs (HFMA2.BF16)                          */
  VFI_DCN_MATH_BF16_TC_HQ = 3 /* REMOVED in round 2 (fp32 four-corner blend on the v4 kernel): VFI_ERR_UNSUPPORTED      */
} vfi_dcn_math;

/* ---- library ----------------------------------------------------------------------------------------------- */
int vfi_abi_version(void);
const char* vfi_version_string(void);
const char* vfi_last_error(void);            /* thread-local; never NULL                                       */
int vfi_check_device(void);                  /* VFI_OK when the current CUDA device is sm_100                  */
/* number of kernels this library has launched on the calling thread since the last vfi_reset_launch_count() */
int64_t vfi_launch_count(void);
void vfi_reset_launch_count(void);

/* Flags of the warp entry points.  The reference line `2.0 * v / max(size-1, 1)` (ema_vfi.py:165-166) has two
 * bit-level meanings: a true IEEE division when the reference runs on CPU (BASELINE config 1, the golden vectors),
 * and a multiplication by the fp32 reciprocal when it runs on CUDA (aten's tensor/scalar fast path).  Default = IEEE. */
#define VFI_WARP_DIV_IEEE 0
#define VFI_WARP_DIV_RECIPROCAL 1
/* vfi_warp_fwd only: `out` is the [B,3,H,W] view of a DCN tail plane (bf16 channels-last records of 8 elements = 16 bytes per
 * pixel, 16-byte aligned) and the kernel writes WHOLE records, [c0 c1 c2 0 | c0 c1 c2 0]: elements 3..7 of every pixel are
 * overwritten.  This is what removes the torch.cat of ema_vfi.py:134 (the warp writes where DCN layer 1 gathers from).
 * Without the flag only the C channels described by `out` are written, whatever its strides are.  The flag with an `out`
 * that is not such a view is VFI_ERR_INVALID. */
#define VFI_WARP_OUT_TAIL_RECORD 2
/* vfi_warp_fwd only.  By default three planar channels with unit pixel strides take the STAGED kernel: per 32 x 32 tile the
 * source window is copied into shared memory by the copy engine (TMA, zero fill outside the frame) and gathered from there;
 * tiles whose corners span more than the window (large incoherent flow) gather through L1 inside the same launch.  Results do
 * not depend on the route.  VFI_WARP_NO_STAGING forces the L1 kernel of round 1 for the whole frame (A/B runs, parity tests);
 * VFI_WARP_COUNT_TILES makes the launch count its staged / L1 tiles (read with vfi_warp_tile_counts; costs one atomic per tile). */
#define VFI_WARP_NO_STAGING 4
#define VFI_WARP_COUNT_TILES 8

/* ---- warp: replaces EMA_VFI.warp, /root/reference/src/models/ema_vfi.py:149-171 ----------------------------- */
/* out[b,c,y,x] = bilinear(src[b,c], x + flow[b,0,y,x], y + flow[b,1,y,x]), zeros outside, align_corners=True,
 * with the reference's normalise/un-normalise round trip replayed in fp32.
 * src/out: [B,C,H,W] same dtype (f32|bf16|f16); flow: [B,2,H,W] f32 or the dtype of src. */
int vfi_warp_fwd(const vfi_tensor* src, const vfi_tensor* flow, const vfi_tensor* out, int32_t flags,
                 vfi_stream_t stream);

/* Diagnostic: tiles the staged warp kernel served from its shared-memory window / through L1 since the last reset, over all
 * launches that carried VFI_WARP_COUNT_TILES on the current device.  Synchronises the device. */
int vfi_warp_tile_counts(uint64_t* staged, uint64_t* direct, int32_t reset);

/* Autograd of the above (aten::grid_sampler_2d_backward chained through ema_vfi.py:165-166).
 * grad_flow [B,2,H,W] f32 is always written.  grad_src may be NULL (the model path: frame2 needs no grad); when
 * given it must be f32, zero-filled by the caller, and receives atomic scatter-adds. */
int vfi_warp_bwd(const vfi_tensor* grad_out, const vfi_tensor* src, const vfi_tensor* flow,
                 const vfi_tensor* grad_flow, const vfi_tensor* grad_src, int32_t flags, vfi_stream_t stream);

/* North-star extension with no reference counterpart: one pass computing
 * out = m * warp(src_a, flow_a) + (1 - m) * warp(src_b, flow_b), m: [B,1,H,W]. */
int vfi_warp_blend_fwd(const vfi_tensor* src_a, const vfi_tensor* flow_a, const vfi_tensor* src_b,
                       const vfi_tensor* flow_b, const vfi_tensor* m, const vfi_tensor* out, int32_t flags,
                       vfi_stream_t stream);

/* ---- DCNv2: replaces torchvision.ops.deform_conv2d as called from ema_vfi.py:60 ---------------------------- */
/* Geometry is the reference's: 3x3 kernel, stride 1, padding 1, dilation 1, groups 1, one offset group, mask on
 * (ema_vfi.py:45-51).  x [B,C,H,W]; offset [B,18,H,W] (2k = row shift, 2k+1 = column shift of tap k = 3i+j);
 * mask [B,9,H,W]; weight [O,C,3,3] contiguous; bias [O] or NULL; out [B,O,H,W]. */
size_t vfi_dcn_workspace_bytes(int64_t B, int64_t C, int64_t O, int64_t H, int64_t W, int32_t math);

/* Tensor-core path operand formats.
 *  - Activation "planes": channels-last bf16 in two dense buffers, main [B,H,W,64] (128 B per pixel: one aligned cache
 *    line per bilinear corner) and tail [B,H,W,8] (16 B per pixel: channels 64..71, zero beyond C; when there are at most four tail channels -- the
 *    reference's C = 67 -- bytes 8..15 of every record must MIRROR bytes 0..7, which every kernel of this library that
 *    writes planes does: the gather may then read either half and spreads its 8-byte reads over all shared-memory banks).
 *  - Weight image (the v6 / v7 kernels' K order, vfi_dcn_k_order): 11 K blocks x [80 rows (o, zero padded)] x 128 B; K block
 *    t < 9 holds the 64 main channels of tap t in the order the producers' tcgen05.st.16x256b mapping implies, block 9 four
 *    tail channels of each of the nine taps (K elements 0..35) and the two bias slots 36 / 37 (the kernels add the bias on the
 *    tensor core; vfi_dcn_pack_weight itself writes no bias), block 10 is zero; each row's eight 16-byte chunks are already
 *    permuted for the SWIZZLE_128B canonical layout (chunk j of row r stored at j ^ (r & 7)).
 *    vfi_dcn_packed_weight_bytes() = 112,640 bytes, 16-byte aligned. */
size_t vfi_dcn_packed_weight_bytes(void);
int vfi_dcn_pack_weight(const void* weight, int32_t weight_dtype, int64_t O, int64_t C, void* packed,
                        vfi_stream_t stream);

/* Host-only query (no GPU needed): K element kk of block kb of the weight image multiplies weight[:, channel, tap] (channel
 * -1: zero padding or the two bias slots 36/37 of block 9).  variant must be 6: the image the v6 / v7 kernels build in their
 * workspace (and vfi_dcn_pack_weight writes), whose main blocks follow the thread <-> column mapping of tcgen05.st.16x256b.
 * (variant 4, the round-1 v4 kernel's image, was removed with that kernel: VFI_ERR_INVALID.) */
int vfi_dcn_k_order(int32_t variant, int32_t kb, int32_t kk, int32_t* tap, int32_t* channel);

/* Converts an activation tensor of any supported layout/dtype (C <= 72) into planes: main_plane B*H*W*64 bf16,
 * tail_plane B*H*W*8 bf16. */
int vfi_dcn_pack_input(const vfi_tensor* x, void* main_plane, void* tail_plane, vfi_stream_t stream);

int vfi_dcn_fwd(const vfi_tensor* x, const vfi_tensor* offset, const vfi_tensor* mask, const void* weight,
                int32_t weight_dtype, const void* bias, int32_t bias_dtype, const vfi_tensor* out, int64_t O,
                int32_t math, void* workspace, size_t workspace_bytes, vfi_stream_t stream);

/* Hot-path form of the forward (tensor-core math only).  Two pieces of glue of the reference are folded in:
 *  - the offset/mask split of ema_vfi.py:57-59: conv27 is the raw [B,27,H,W] offset_conv output; offsets are its
 *    thirds 0 and 2, the mask is sigmoid(third 1), computed in the kernel's geometry stage;
 *  - the torch.cat of ema_vfi.py:134: the input is given as planes, x_main [B,64,H,W] + x_tail [B,<=8,H,W] (dense
 *    channels-last bf16, pixel strides 64 and 8 elements, pad channels of the tail zero, upper record half mirrored for <= 4 tail channels) -- feat and the warped frame
 *    as they exist before the cat.  x_tail may be NULL, in which case x_main is handled as vfi_dcn_fwd handles x.
 * Output: planes when out_tail != NULL (out [B,64,H,W], out_tail [B,O-64,H,W]; what the next layer reads directly),
 * otherwise any strided [B,O,H,W] tensor. */
int vfi_dcn_fwd_fused(const vfi_tensor* x_main, const vfi_tensor* x_tail, const vfi_tensor* conv27, const void* weight,
                      int32_t weight_dtype, const void* bias, int32_t bias_dtype, const vfi_tensor* out,
                      const vfi_tensor* out_tail, int64_t O, int32_t math, void* workspace, size_t workspace_bytes,
                      vfi_stream_t stream);

/* Gradients (torchvision::_deform_conv2d_backward).  grad_out must have x's dtype.  All gradient tensors are f32.  grad_x must be zero-filled by
 * the caller (atomic scatter-add target); grad_weight / grad_bias likewise (accumulated with atomics so that a
 * caller may point them into a flat gradient bucket).  Any of the three outputs of bwd_data may be NULL. */
int vfi_dcn_bwd_data(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* offset,
                     const vfi_tensor* mask, const void* weight, int32_t weight_dtype, int64_t O,
                     const vfi_tensor* grad_x, const vfi_tensor* grad_offset, const vfi_tensor* grad_mask,
                     void* workspace, size_t workspace_bytes, vfi_stream_t stream);
int vfi_dcn_bwd_weight(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* offset,
                       const vfi_tensor* mask, int64_t O, float* grad_weight, float* grad_bias,
                       vfi_stream_t stream);

/* grad_weight / grad_bias on the tensor cores (bf16 operands, fp32 accumulation in tensor memory): the same contraction as
 * vfi_dcn_bwd_weight with the activations, the modulated samples and grad_out rounded to bf16 -- the training companion of
 * VFI_DCN_MATH_BF16_TC (tolerance: max|delta| / max|ref| <= 1e-2).  x is any [B,C,H,W] tensor with C <= 68 (packed to planes
 * in the workspace, vfi_dcn_workspace_bytes(..., VFI_DCN_MATH_BF16_TC)); offset / mask 16-bit with unit pixel stride and
 * W % 8 == 0; grad_out bf16 or f32, any strides (NCHW or channels_last).  Accumulates into grad_weight [O,C,3,3] / grad_bias [O] (either may be NULL) with fp32
 * atomics, so the caller zero-fills them or points them into a flat gradient bucket.  Anything outside these limits
 * returns VFI_ERR_UNSUPPORTED and the caller uses vfi_dcn_bwd_weight. */
int vfi_dcn_bwd_weight_tc(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* offset, const vfi_tensor* mask,
                          int64_t O, float* grad_weight, float* grad_bias, void* workspace, size_t workspace_bytes,
                          vfi_stream_t stream);

/* grad_x / grad_offset / grad_mask, column-gradient form.  torchvision::_deform_conv2d_backward (reference call site
 * src/models/ema_vfi.py:60 via autograd) splits into a dense product and a position-dependent part:
 *   gcol[p, k*72 + c] = sum_o grad_out[p, o] * weight[o, c, k]      (plain GEMM [P,72] x [72,648]: vfi_dcn_gcol below for bf16 columns, the caller's BLAS for true-fp32 ones; rows of
 *                                                                   gcol_ld >= 648 elements, gcol_ld % 4 == 0, columns
 *                                                                   c >= C of every tap zero; gcol_dtype VFI_BF16 or VFI_F32)
 * and this call, which gathers the four corner rows of x per (pixel, tap), reduces <gcol, corner> to grad_mask [B,9,H,W] and
 * grad_offset [B,18,H,W] (f32, any strides) and adds gcol * mask * corner weight into grad_x_rows: a CHANNELS-LAST f32
 * accumulator [B*H*W][grad_x_ld] (grad_x_ld >= 68, % 4 == 0, 16-byte aligned, zero-filled by the caller; column c = channel
 * c), i.e. grad_x = grad_x_rows[:, :C] viewed as [B,H,W,C].  x: any [B,C,H,W] tensor, C <= 68, packed to channels-last planes
 * of gcol's dtype in the workspace (vfi_dcn_bwd_data_cols_workspace_bytes; bf16 gcol rounds x to bf16, f32 gcol keeps fp32
 * arithmetic throughout: tolerance of vfi_dcn_bwd_data); offset / mask: any dtype and strides.  Any of the three outputs may
 * be NULL.  Semantics (liveness, corner validity, ungated offset gradient) are vfi_dcn_bwd_data's. */
/* The column gradient itself on the tensor cores (bf16 operands, fp32 accumulation in tensor memory, bf16 result): replaces the
 * grad_columns GEMM torchvision::_deform_conv2d_backward makes with its BLAS.  grad_out: [B,O,H,W], any dtype and strides, read
 * where it lies (O <= 68); weight: [O,C,3,3] contiguous (C <= 72); gcol: [B*H*W][648] bf16, 16-byte aligned, gcol_ld = 648,
 * column k * 72 + c (columns c >= C of every tap are written as zeros).  Workspace: vfi_dcn_gcol_workspace_bytes(), 256-byte
 * aligned. */
size_t vfi_dcn_gcol_workspace_bytes(void);
int vfi_dcn_gcol(const vfi_tensor* grad_out, const void* weight, int32_t weight_dtype, int64_t C, void* gcol, int64_t gcol_ld,
                 void* workspace, size_t workspace_bytes, vfi_stream_t stream);
size_t vfi_dcn_bwd_data_cols_workspace_bytes(int64_t B, int64_t H, int64_t W, int32_t gcol_dtype);
int vfi_dcn_bwd_data_cols(const void* gcol, int32_t gcol_dtype, int64_t gcol_ld, const vfi_tensor* x, const vfi_tensor* offset,
                          const vfi_tensor* mask, float* grad_x_rows, int64_t grad_x_ld, const vfi_tensor* grad_offset,
                          const vfi_tensor* grad_mask, void* workspace, size_t workspace_bytes, vfi_stream_t stream);

/* Backward of vfi_dcn_fwd_fused (the training companion of the hot-path form; round 2).  The glue of ema_vfi.py:57-59 stays
 * folded in: offsets and the mask come from the raw 27-channel offset_conv output conv27 (16-bit; sigmoid computed in the
 * kernel exactly as the forward does), and the data gradient goes back to THAT tensor.
 *  - vfi_dcn_bwd_weight_tc_fused: vfi_dcn_bwd_weight_tc with (offset, mask) replaced by conv27 (NCHW with 16-byte aligned rows,
 *    or dense channels-last [B,H,W,27]); same workspace, same limits, same accumulation into grad_weight / grad_bias.
 *  - vfi_dcn_bwd_data_cols_fused: vfi_dcn_bwd_data_cols for bf16 columns with x given as planes, read where they lie (x_main
 *    [B,64,H,W] + x_tail [B,<=4,H,W], bf16 channels-last views with dense rows and a pixel stride that is a multiple of 16
 *    bytes: two dense plane buffers or the channel ranges 0..63 / 64.. of one [B,H,W,72] record buffer -- no packing pass, no
 *    workspace); grad_conv27: f32 [B,27,H,W], any strides, receives d loss / d conv27 (channels 0..8 and 18..26: the offset
 *    gradient in the reference's cat(o1, o2) order; channels 9..17: grad_mask * m * (1 - m)).  grad_x_rows as in
 *    vfi_dcn_bwd_data_cols.  Either output may be NULL. */
int vfi_dcn_bwd_weight_tc_fused(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* conv27, int64_t O,
                                float* grad_weight, float* grad_bias, void* workspace, size_t workspace_bytes, vfi_stream_t stream);
int vfi_dcn_bwd_data_cols_fused(const void* gcol, int64_t gcol_ld, const vfi_tensor* x_main, const vfi_tensor* x_tail,
                                const vfi_tensor* conv27, float* grad_x_rows, int64_t grad_x_ld, const vfi_tensor* grad_conv27,
                                vfi_stream_t stream);

/* ---- diagnostics ------------------------------------------------------------------------------------------ */
/* D[128,80] (f32, row-major) = A[128,K] * B[80,K]^T with A, B row-major bf16 in device memory, K % 64 == 0.  One CTA,
 * serialised; exercises exactly the shared-memory descriptors, swizzle, tcgen05.mma/commit/ld and TMEM allocation the
 * DCN kernel uses, so a wrong DCN result can be attributed to the tensor-core plumbing or to the gather. */
int vfi_selftest_umma(const void* a_bf16, const void* b_bf16, float* d, int32_t K, vfi_stream_t stream);
/* D[128,80] = 2 * A[128,64] * B[80,64]^T with the A operand written to TENSOR MEMORY (once by tcgen05.st.16x256b with the
 * thread <-> (lane, column) mapping the v6 DCN producers use, once by tcgen05.st.32x32b) and consumed by the A-from-TMEM
 * form of tcgen05.mma.  raw [128][32] u32 receives the TMEM image of the first copy (K elements 2c, 2c+1 of row r at
 * raw[r][c]) so a wrong mapping can be decoded on the host. */
int vfi_selftest_umma_ts(const void* a_bf16, const void* b_bf16, float* d, uint32_t* raw, vfi_stream_t stream);
/* With VFI_DCN_DEBUG=1 in the environment the staged tcgen05 DCN kernel records, per CTA and warp, the cycles spent in
 * each of its pipeline waits; this copies `count` u64 counters ([cta][32 warps][8]) of the last launch to host memory
 * (synchronises the device). */
int vfi_debug_read(uint64_t* host_dst, size_t count);
/* Pipeline watchdog of the tcgen05 kernels (builds with -DVFI_WATCHDOG only; the default build traps after ~5 s instead and
 * this call then always returns 0): a warp that waits on an mbarrier for more than ~2 s records itself, raises a
 * flag that makes every other waiter leave as well, and the kernel ends instead of hanging the GPU (its results are
 * invalid).  Returns 1 if that happened since the last call (synchronises the device, clears the flag); info36 (may be NULL)
 * receives [block << 32 | warp, shared-memory address of the barrier, parity waited for, number of waiters that gave up] and,
 * for each of the 32 warps of that block, the barrier it was waiting on (address | parity << 32, 0 = not waiting). */
int vfi_debug_abort_info(uint64_t* info36);

#ifdef __cplusplus
}
#endif
#endif /* VFI_B200_H_ */
