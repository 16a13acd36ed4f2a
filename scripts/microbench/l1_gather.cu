// Microbenchmark: bytes/cycle/SM of the load flavours the DCN producer could use, all data L1- or smem-resident.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l1_gather l1_gather.cu && ./l1_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 2000;

// mode 0: LDG.128, a warp reads 4 aligned 128 B lines (8 lanes per line)     -> 512 B / instr
// mode 1: LDG.32,  a warp reads 1 aligned 128 B line                         -> 128 B / instr
// mode 2: LDS.128, a warp reads 4 rows of 128 B from shared memory           -> 512 B / instr
// mode 3: LDS.32,  a warp reads 1 row of 128 B                               -> 128 B / instr
// mode 4: LDG.64,  a warp reads 2 aligned lines                              -> 256 B / instr
template <int MODE>
__global__ void __launch_bounds__(1024) bench(const uint8_t* __restrict__ g, unsigned long long* out, int lines) {
  extern __shared__ uint8_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x * 16; i < lines * 128; i += blockDim.x * 16) *reinterpret_cast<uint4*>(sm + i) = make_uint4(i, 1, 2, 3);
  __syncthreads();
  uint32_t acc = 0;
  uint32_t line = (warp * 37 + blockIdx.x * 11) & (lines - 1);
  // warm L1
  for (int i = threadIdx.x; i < lines * 8; i += blockDim.x) acc += reinterpret_cast<const uint4*>(g)[i].x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 16
  for (int it = 0; it < ITERS; ++it) {
    line = (line + 29) & (lines - 1);         // cheap walk over the lines (lines is a power of two), same for the whole warp
    if (MODE == 0) {
      uint32_t l = (line + (lane >> 3) * 13) & (lines - 1);
      uint4 v = __ldg(reinterpret_cast<const uint4*>(g + (size_t)l * 128 + (lane & 7) * 16));
      acc += v.x ^ v.y ^ v.z ^ v.w;
    } else if (MODE == 1) {
      acc += __ldg(reinterpret_cast<const uint32_t*>(g + (size_t)line * 128 + lane * 4));
    } else if (MODE == 2) {
      uint32_t l = (line + (lane >> 3) * 13) & (lines - 1);
      uint4 v = *reinterpret_cast<const uint4*>(sm + l * 128 + (lane & 7) * 16);
      acc += v.x ^ v.y ^ v.z ^ v.w;
    } else if (MODE == 3) {
      acc += *reinterpret_cast<const uint32_t*>(sm + line * 128 + lane * 4);
    } else {
      uint32_t l = (line + (lane >> 4) * 13) & (lines - 1);
      uint2 v = __ldg(reinterpret_cast<const uint2*>(g + (size_t)l * 128 + (lane & 15) * 8));
      acc += v.x ^ v.y;
    }
  }
  long long t1 = clock64();
  if (acc == 0x12345678) out[1] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
}

template <int MODE>
void run(const char* name, int bytes_per_instr, const uint8_t* g, unsigned long long* out, int lines) {
  int smem = lines * 128;
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int warps : {8, 16, 32}) {
    bench<MODE><<<148, warps * 32, smem>>>(g, out, lines);
    cudaDeviceSynchronize();
    bench<MODE><<<148, warps * 32, smem>>>(g, out, lines);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    double cyc_per_instr = (double)h[0] / ITERS / warps;    // SM-level cycles per warp-instruction
    printf("%-28s warps/SM=%2d  %6.2f cyc/warp-instr  %6.1f B/cyc/SM  (%s)\n", name, warps, cyc_per_instr,
           bytes_per_instr / cyc_per_instr, cudaGetErrorString(e));
  }
}

int main() {
  const int lines = 256;   // 32 KB working set: fits L1 next to 32 KB of smem
  uint8_t* g; unsigned long long* out;
  cudaMalloc(&g, 1 << 20); cudaMemset(g, 1, 1 << 20); cudaMalloc(&out, 16);
  run<0>("LDG.128 4 lines/instr", 512, g, out, lines);
  run<4>("LDG.64  2 lines/instr", 256, g, out, lines);
  run<1>("LDG.32  1 line/instr", 128, g, out, lines);
  run<2>("LDS.128 4 rows/instr", 512, g, out, lines);
  run<3>("LDS.32  1 row/instr", 128, g, out, lines);
  return 0;
}
