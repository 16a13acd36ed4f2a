// common.cuh -- shared host/device helpers for libvfi_b200 (sm_100a only).
#pragma once
#include <cuda.h>        // CUtensorMap (types only: the encoder is resolved at run time, see tensor_map_encoder)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/vfi_b200.h"

namespace vfi {

// ---------------------------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);   // defined in abi.cu (thread-local buffer)
void count_launch(int n = 1);           // defined in abi.cu

#define VFI_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::vfi::set_error(__VA_ARGS__);      \
      return (code);                      \
    }                                     \
  } while (0)

#define VFI_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      ::vfi::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return VFI_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

// Launch-error check that does not synchronise.
#define VFI_LAUNCH_CHECK(name)                                                              \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess) {                                                               \
      ::vfi::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));           \
      return VFI_ERR_CUDA;                                                                  \
    }                                                                                       \
    ::vfi::count_launch();                                                                  \
  } while (0)

// ---------------------------------------------------------------------------------------------- tensors
// Device-side view of a vfi_tensor (typed pointer left to the kernel template).
struct TView {
  const void* p;
  long long sn, sc, sh, sw;
};

inline TView view(const vfi_tensor* t) { return TView{t->data, t->sn, t->sc, t->sh, t->sw}; }

inline bool same_shape(const vfi_tensor* a, const vfi_tensor* b) {
  return a->n == b->n && a->c == b->c && a->h == b->h && a->w == b->w;
}
inline bool is_nchw_contig(const vfi_tensor* t) {
  return t->sw == 1 && t->sh == t->w && t->sc == t->h * t->w && t->sn == t->c * t->h * t->w;
}
inline size_t dtype_size(int dt) { return dt == VFI_F32 ? 4 : 2; }
inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---------------------------------------------------------------------------------------------- scalar I/O
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

template <typename T> __device__ __forceinline__ float ldg_f32(const T* p) { return to_f32<T>(__ldg(p)); }

// Dispatch a vfi_dtype to a C++ type.  Usage: VFI_DISPATCH(dt, T, { kernel<T><<<...>>>(...); })
#define VFI_DISPATCH(dt, T, ...)                                   \
  switch (dt) {                                                    \
    case VFI_F32: { using T = float; __VA_ARGS__; break; }         \
    case VFI_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }\
    case VFI_F16: { using T = __half; __VA_ARGS__; break; }        \
    default: ::vfi::set_error("unknown dtype %d", (int)(dt)); return VFI_ERR_INVALID; \
  }

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------- TMA tensor maps (host)
// cuTensorMapEncodeTiled, resolved at run time (the library has no link-time dependency on libcuda: it loads in the CPU-only
// build container).  Returns null when the driver does not export it.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

}  // namespace vfi
