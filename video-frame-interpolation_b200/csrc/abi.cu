// abi.cu -- library-level entry points of libvfi_b200.so: version, errors, device check, DCN dispatch.
#include <cstring>

#include "common.cuh"

namespace vfi {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

// dcn_simt.cu
size_t dcn_simt_workspace_bytes();
int dcn_simt_fwd(const vfi_tensor* x, const vfi_tensor* offset, const vfi_tensor* mask, const void* weight,
                 int weight_dtype, const void* bias, int bias_dtype, const vfi_tensor* out, long long O, void* workspace,
                 size_t workspace_bytes, cudaStream_t st);
// dcn_tc.cu
bool dcn_tc_available();
size_t dcn_tc_workspace_bytes(long long B, long long H, long long W);
size_t dcn_tc_packed_weight_bytes();
int dcn_tc_pack_weight(const void* weight, int weight_dtype, const void* bias, int bias_dtype, long long O, long long C,
                       void* packed, float* bias_out, cudaStream_t st, int variant);
int umma_ts_selftest(const void* A, const void* Bm, float* D, uint32_t* raw, cudaStream_t st);
int dcn_tc_k_order(int variant, int kb, int kk, int* tap, int* channel);
int dcn_tc_abort_info(unsigned long long* info);
int dcn_tc_bwd_weight(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* offset, const vfi_tensor* mask,
                      long long O, float* gw, float* gb, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t dcn_tc_bwd_data_cols_workspace_bytes(long long B, long long H, long long W, int gcol_dtype);
int dcn_tc_bwd_data_cols(const void* gcol, int gcol_dtype, long long gcol_ld, const vfi_tensor* x, const vfi_tensor* offset,
                         const vfi_tensor* mask, float* gx_rows, long long gx_ld, const vfi_tensor* grad_offset,
                         const vfi_tensor* grad_mask, void* workspace, size_t workspace_bytes, cudaStream_t st);
int dcn_tc_bwd_weight_fused(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* conv27, long long O, float* gw, float* gb,
                            void* workspace, size_t workspace_bytes, cudaStream_t st);
int dcn_tc_bwd_data_cols_fused(const void* gcol, long long gcol_ld, const vfi_tensor* x_main, const vfi_tensor* x_tail,
                               const vfi_tensor* conv27, float* gx_rows, long long gx_ld, const vfi_tensor* grad_conv27,
                               cudaStream_t st);
int umma_selftest(const void* A, const void* Bm, float* D, int K, cudaStream_t st);
size_t dcn_tc_gcol_workspace_bytes();
int dcn_tc_gcol(const vfi_tensor* grad_out, const void* weight, int weight_dtype, long long C, void* gcol, long long gcol_ld,
                void* workspace, size_t workspace_bytes, cudaStream_t st);
unsigned long long* dcn_tc_debug_buffer();
int dcn_tc_pack_input(const vfi_tensor* x, void* main_plane, void* tail_plane, cudaStream_t st);
int dcn_tc_fwd(const vfi_tensor* x, const vfi_tensor* offset, const vfi_tensor* mask, const void* weight,
               int weight_dtype, const void* bias, int bias_dtype, const vfi_tensor* out, long long O, bool hq,
               void* workspace, size_t workspace_bytes, cudaStream_t st);
int dcn_tc_fwd_fused(const vfi_tensor* x_main, const vfi_tensor* x_tail, const vfi_tensor* conv27, const void* weight,
                     int weight_dtype, const void* bias, int bias_dtype, const vfi_tensor* out, const vfi_tensor* out_tail,
                     long long O, bool hq, void* workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace vfi

using namespace vfi;

extern "C" int vfi_abi_version(void) { return VFI_B200_ABI_VERSION; }
extern "C" const char* vfi_version_string(void) { return "vfi_b200 0.1 (sm_100a; warp + DCNv2 hot path)"; }
extern "C" const char* vfi_last_error(void) { return g_err; }
extern "C" int64_t vfi_launch_count(void) { return g_launches; }
extern "C" void vfi_reset_launch_count(void) { g_launches = 0; }

extern "C" int vfi_check_device(void) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no usable CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    return VFI_ERR_DEVICE;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    set_error("device %d is sm_%d%d; libvfi_b200 is built for sm_100a (B200) only", dev, major, minor);
    return VFI_ERR_DEVICE;
  }
  return VFI_OK;
}

static int resolve_math(const vfi_tensor* x, int32_t math) {
  if (math != VFI_DCN_MATH_AUTO) return math;
  if (!dcn_tc_available()) return VFI_DCN_MATH_FP32;   // fp32 arithmetic on whatever dtype the tensors have
  return (x && x->dtype == VFI_F32) ? VFI_DCN_MATH_FP32 : VFI_DCN_MATH_BF16_TC;
}

extern "C" size_t vfi_dcn_workspace_bytes(int64_t B, int64_t C, int64_t O, int64_t H, int64_t W, int32_t math) {
  (void)C; (void)O;
  size_t simt = dcn_simt_workspace_bytes();
  if (math == VFI_DCN_MATH_FP32) return simt;
  size_t tc = dcn_tc_workspace_bytes(B, H, W);
  return tc > simt ? tc : simt;
}

extern "C" size_t vfi_dcn_packed_weight_bytes(void) { return dcn_tc_packed_weight_bytes(); }

extern "C" int vfi_dcn_pack_weight(const void* weight, int32_t weight_dtype, int64_t O, int64_t C, void* packed,
                                   vfi_stream_t stream) {
  return dcn_tc_pack_weight(weight, weight_dtype, nullptr, VFI_F32, O, C, packed, nullptr, (cudaStream_t)stream, 4);
}

extern "C" int vfi_dcn_pack_input(const vfi_tensor* x, void* main_plane, void* tail_plane, vfi_stream_t stream) {
  return dcn_tc_pack_input(x, main_plane, tail_plane, (cudaStream_t)stream);
}

extern "C" int vfi_dcn_fwd(const vfi_tensor* x, const vfi_tensor* offset, const vfi_tensor* mask, const void* weight,
                           int32_t weight_dtype, const void* bias, int32_t bias_dtype, const vfi_tensor* out, int64_t O,
                           int32_t math, void* workspace, size_t workspace_bytes, vfi_stream_t stream) {
  VFI_REQUIRE(x, VFI_ERR_INVALID, "vfi_dcn_fwd: null tensor descriptor");
  int m = resolve_math(x, math);
  if (m == VFI_DCN_MATH_FP32)
    return dcn_simt_fwd(x, offset, mask, weight, weight_dtype, bias, bias_dtype, out, O, workspace, workspace_bytes,
                        (cudaStream_t)stream);
  if (m == VFI_DCN_MATH_BF16_TC || m == VFI_DCN_MATH_BF16_TC_HQ)
    return dcn_tc_fwd(x, offset, mask, weight, weight_dtype, bias, bias_dtype, out, O, m == VFI_DCN_MATH_BF16_TC_HQ,
                      workspace, workspace_bytes, (cudaStream_t)stream);
  set_error("vfi_dcn_fwd: unknown math mode %d", (int)math);
  return VFI_ERR_INVALID;
}

extern "C" int vfi_selftest_umma(const void* a_bf16, const void* b_bf16, float* d, int32_t K, vfi_stream_t stream) {
  return umma_selftest(a_bf16, b_bf16, d, K, (cudaStream_t)stream);
}

extern "C" int vfi_dcn_bwd_weight_tc(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* offset,
                                     const vfi_tensor* mask, int64_t O, float* grad_weight, float* grad_bias, void* workspace,
                                     size_t workspace_bytes, vfi_stream_t stream) {
  return dcn_tc_bwd_weight(grad_out, x, offset, mask, O, grad_weight, grad_bias, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t vfi_dcn_bwd_data_cols_workspace_bytes(int64_t B, int64_t H, int64_t W, int32_t gcol_dtype) {
  return dcn_tc_bwd_data_cols_workspace_bytes(B, H, W, gcol_dtype);
}

extern "C" int vfi_dcn_bwd_data_cols(const void* gcol, int32_t gcol_dtype, int64_t gcol_ld, const vfi_tensor* x,
                                     const vfi_tensor* offset, const vfi_tensor* mask, float* grad_x_rows, int64_t grad_x_ld,
                                     const vfi_tensor* grad_offset, const vfi_tensor* grad_mask, void* workspace,
                                     size_t workspace_bytes, vfi_stream_t stream) {
  return dcn_tc_bwd_data_cols(gcol, gcol_dtype, gcol_ld, x, offset, mask, grad_x_rows, grad_x_ld, grad_offset, grad_mask,
                              workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int vfi_dcn_bwd_weight_tc_fused(const vfi_tensor* grad_out, const vfi_tensor* x, const vfi_tensor* conv27, int64_t O,
                                           float* grad_weight, float* grad_bias, void* workspace, size_t workspace_bytes,
                                           vfi_stream_t stream) {
  return dcn_tc_bwd_weight_fused(grad_out, x, conv27, O, grad_weight, grad_bias, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int vfi_dcn_bwd_data_cols_fused(const void* gcol, int64_t gcol_ld, const vfi_tensor* x_main, const vfi_tensor* x_tail,
                                           const vfi_tensor* conv27, float* grad_x_rows, int64_t grad_x_ld,
                                           const vfi_tensor* grad_conv27, vfi_stream_t stream) {
  return dcn_tc_bwd_data_cols_fused(gcol, gcol_ld, x_main, x_tail, conv27, grad_x_rows, grad_x_ld, grad_conv27, (cudaStream_t)stream);
}

extern "C" size_t vfi_dcn_gcol_workspace_bytes(void) { return dcn_tc_gcol_workspace_bytes(); }

extern "C" int vfi_dcn_gcol(const vfi_tensor* grad_out, const void* weight, int32_t weight_dtype, int64_t C, void* gcol,
                            int64_t gcol_ld, void* workspace, size_t workspace_bytes, vfi_stream_t stream) {
  return dcn_tc_gcol(grad_out, weight, weight_dtype, C, gcol, gcol_ld, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int vfi_debug_abort_info(uint64_t* info36) { return dcn_tc_abort_info(reinterpret_cast<unsigned long long*>(info36)); }

extern "C" int vfi_dcn_k_order(int32_t variant, int32_t kb, int32_t kk, int32_t* tap, int32_t* channel) {
  return dcn_tc_k_order(variant, kb, kk, tap, channel);
}

extern "C" int vfi_selftest_umma_ts(const void* a_bf16, const void* b_bf16, float* d, uint32_t* raw, vfi_stream_t stream) {
  return umma_ts_selftest(a_bf16, b_bf16, d, raw, (cudaStream_t)stream);
}

extern "C" int vfi_dcn_fwd_fused(const vfi_tensor* x_main, const vfi_tensor* x_tail, const vfi_tensor* conv27,
                                 const void* weight, int32_t weight_dtype, const void* bias, int32_t bias_dtype,
                                 const vfi_tensor* out, const vfi_tensor* out_tail, int64_t O, int32_t math,
                                 void* workspace, size_t workspace_bytes, vfi_stream_t stream) {
  VFI_REQUIRE(math == VFI_DCN_MATH_AUTO || math == VFI_DCN_MATH_BF16_TC || math == VFI_DCN_MATH_BF16_TC_HQ,
              VFI_ERR_UNSUPPORTED, "vfi_dcn_fwd_fused: only the tensor-core math modes are implemented in fused form");
  return dcn_tc_fwd_fused(x_main, x_tail, conv27, weight, weight_dtype, bias, bias_dtype, out, out_tail, O,
                          math == VFI_DCN_MATH_BF16_TC_HQ, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int vfi_debug_read(uint64_t* host_dst, size_t count) {
  unsigned long long* b = dcn_tc_debug_buffer();
  VFI_REQUIRE(b && host_dst, VFI_ERR_UNSUPPORTED, "vfi_debug_read: set VFI_DCN_DEBUG=1 before the first DCN launch");
  VFI_CUDA(cudaMemcpy(host_dst, b, (count < 256 * 32 * 8 ? count : 256 * 32 * 8) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return VFI_OK;
}
