// dcn_tc.cu -- DCNv2 forward as a tcgen05 / TMEM implicit GEMM (bf16 operands, fp32 accumulate).  PLACEHOLDER:
// the operand packers are real; the tensor-core kernel lands in the next commit.
#include "common.cuh"

namespace vfi {

constexpr int TC_N = 80;        // UMMA N: 67 output channels padded to a multiple of 16
constexpr int TC_CPAD = 72;     // channels per tap in the K dimension (67 padded to a multiple of 8)
constexpr int TC_K = 656;       // 9 * 72 = 648 padded to a multiple of UMMA_K = 16

bool dcn_tc_available() { return false; }
size_t dcn_tc_packed_weight_bytes() { return (size_t)TC_N * TC_K * 2; }
size_t dcn_tc_workspace_bytes(long long B, long long H, long long W) {
  size_t w = ((dcn_tc_packed_weight_bytes() + 255) / 256) * 256;
  size_t bias = 512;
  size_t xin = (((size_t)B * H * W * TC_CPAD * 2) + 255) / 256 * 256;
  return w + bias + xin;
}

namespace {
template <typename TW>
__global__ void pack_weight_kernel(const TW* __restrict__ w, int O, int C, __nv_bfloat16* __restrict__ packed) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= TC_N * TC_K) return;
  int o = idx / TC_K, q = idx % TC_K;
  int k = q / TC_CPAD, c = q % TC_CPAD;
  float v = 0.0f;
  if (o < O && k < 9 && c < C) v = to_f32<TW>(w[((size_t)o * C + c) * 9 + k]);
  packed[idx] = __float2bfloat16_rn(v);
}

template <typename TX>
__global__ void pack_input_kernel(const TX* __restrict__ x, long long sn, long long sc, long long sh, long long sw, int B,
                                  int C, int H, int W, __nv_bfloat16* __restrict__ packed) {
  // one thread per (pixel, 8-channel chunk); simple strided reads (v1), 16-byte writes
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * H * W * (TC_CPAD / 8);
  if (idx >= total) return;
  int chunk = (int)(idx % (TC_CPAD / 8));
  long long pix = idx / (TC_CPAD / 8);
  int xx = (int)(pix % W);
  long long t = pix / W;
  int y = (int)(t % H);
  int b = (int)(t / H);
  const TX* src = x + b * sn + y * sh + xx * sw;
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int c = chunk * 8 + i;
    v[i] = __float2bfloat16_rn(c < C ? to_f32<TX>(__ldg(src + c * sc)) : 0.0f);
  }
  *reinterpret_cast<uint4*>(packed + pix * TC_CPAD + chunk * 8) = *reinterpret_cast<uint4*>(v);
}
}  // namespace

int dcn_tc_pack_weight(const void* weight, int weight_dtype, long long O, long long C, void* packed, cudaStream_t st) {
  VFI_REQUIRE(weight && packed, VFI_ERR_INVALID, "vfi_dcn_pack_weight: null pointer");
  VFI_REQUIRE(O > 0 && O <= TC_N && C > 0 && C <= TC_CPAD, VFI_ERR_UNSUPPORTED,
              "vfi_dcn_pack_weight: tensor-core path supports C <= %d, O <= %d", TC_CPAD, TC_N);
  VFI_DISPATCH(weight_dtype, TW, {
    pack_weight_kernel<TW><<<ceil_div(TC_N * TC_K, 256), 256, 0, st>>>(reinterpret_cast<const TW*>(weight), (int)O, (int)C,
                                                                       reinterpret_cast<__nv_bfloat16*>(packed));
  });
  VFI_LAUNCH_CHECK("pack_weight_kernel");
  return VFI_OK;
}

int dcn_tc_pack_input(const vfi_tensor* x, void* packed, cudaStream_t st) {
  VFI_REQUIRE(x && x->data && packed, VFI_ERR_INVALID, "vfi_dcn_pack_input: null pointer");
  VFI_REQUIRE(x->c > 0 && x->c <= TC_CPAD, VFI_ERR_UNSUPPORTED, "vfi_dcn_pack_input: C must be <= %d", TC_CPAD);
  VFI_REQUIRE(aligned(packed, 16), VFI_ERR_INVALID, "vfi_dcn_pack_input: destination must be 16-byte aligned");
  long long total = (long long)x->n * x->h * x->w * (TC_CPAD / 8);
  if (total == 0) return VFI_OK;
  VFI_DISPATCH(x->dtype, TX, {
    pack_input_kernel<TX><<<ceil_div(total, 256), 256, 0, st>>>(reinterpret_cast<const TX*>(x->data), x->sn, x->sc, x->sh,
                                                               x->sw, (int)x->n, (int)x->c, (int)x->h, (int)x->w,
                                                               reinterpret_cast<__nv_bfloat16*>(packed));
  });
  VFI_LAUNCH_CHECK("pack_input_kernel");
  return VFI_OK;
}

int dcn_tc_fwd(const vfi_tensor*, const vfi_tensor*, const vfi_tensor*, const void*, int, const void*, int,
               const vfi_tensor*, long long, void*, size_t, cudaStream_t) {
  set_error("vfi_dcn_fwd: the tcgen05 path is not built into this library yet");
  return VFI_ERR_UNSUPPORTED;
}

}  // namespace vfi
