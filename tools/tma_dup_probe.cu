// Can a TMA tensor map REPEAT pixels (global stride 0 on extra dimensions)?  If so, nearest-neighbour x2 upsampling
// folds into the consumer conv's A-operand loads: map dims (c, dup_x = 2, w/2, dup_y = 2, n*h/2) with strides
// (-, 0, pitch, 0, row pitch) and box (64, 2, TW/2, 2, TH/2) lands in smem in exactly the hi-res pixel order.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_dup_probe tools/tma_dup_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, __half* out, int bytes) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u, bar = base, data = base + 1024;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(data), "l"((uint64_t)&tm), "r"(bar), "r"(0), "r"(0), "r"(1), "r"(0), "r"(1) : "memory");
    uint32_t ok = 0, spins = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
      if (++spins > (1u << 24)) __trap();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<__half*>(sm + (data - smem_u32(sm)))[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const int H2 = 4, W2 = 4, C = 64;
  std::vector<__half> h(H2 * W2 * C);
  for (int y = 0; y < H2; ++y) for (int x = 0; x < W2; ++x) for (int c = 0; c < C; ++c) h[(y * W2 + x) * C + c] = __float2half(y * 10 + x + c * 0.001f);
  __half *d, *o; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMalloc(&o, 65536)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  for (int swz = 0; swz < 2; ++swz) {
    CUtensorMap tm;
    cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)W2, 2, (cuuint64_t)H2};
    cuuint64_t st[4] = {0, (cuuint64_t)C * 2, 0, (cuuint64_t)C * 2 * W2};
    cuuint32_t box[5] = {64, 2, 2, 2, 2}, es[5] = {1, 1, 1, 1, 1};   // hi-res tile: 4 rows x 4 cols starting at lo-res (x=1, y=1)
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle %d: encode with zero strides -> CUresult %d\n", swz, (int)r);
    if (r != CUDA_SUCCESS) continue;
    const int bytes = 16 * 128;
    k<<<1, 128, 8192>>>(tm, o, bytes);
    CK(cudaDeviceSynchronize());
    std::vector<__half> res(bytes / 2);
    CK(cudaMemcpy(res.data(), o, bytes, cudaMemcpyDeviceToHost));
    printf("first element of each 128-B smem row (expect hi-res rows y=2..5, x=2..5 -> lo-res y*10+x = 11 11 12 12 / 11 11 12 12 / 21 21 22 22 / 21 21 22 22):\n");
    for (int row = 0; row < 16; ++row) printf("%5.1f%s", __half2float(res[row * 64 + (swz ? ((0 ^ (row & 7)) * 8) : 0)]), row % 4 == 3 ? "\n" : " ");
  }
  return 0;
}
