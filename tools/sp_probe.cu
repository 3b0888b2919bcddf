// tcgen05.mma.sp (2:4 structured-sparse A operand, kind::f16) probe for B200.
//   part 1  DECODES the sparsity-metadata layout in tensor memory empirically: B is the 32x32 identity, the compressed
//           A rows hold 1..16, so D[m][n] (fp32) IS the decompressed logical A row m -- every stored value shows up at
//           the logical K position the metadata sends it to.  Several metadata patterns / columns / selector values.
//   part 2  times dense (K = 16) vs sparse (K = 32 logical) MMAs for several N, single CTA and CTA pair, to see whether the
//           sparse instruction runs at the dense instruction's cycle count (2x the math) and where shared-memory operand
//           reads bound it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/sp_probe tools/sp_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../coco-dataset-based-light-weight-fast-object-detection-model_b200/csrc/yx_ptx.cuh"

using namespace yx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// idesc: D = f32, A = B = f16, K-major, M = 128 (256 for the pair), N = n; bit 2 = sparse, bits 0-1 = metadata selector
__device__ __forceinline__ uint32_t idesc_sp(uint32_t m, uint32_t n, uint32_t id2) {
  return (id2 & 3u) | (1u << 2) | (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ void umma_sp(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t e_tmem,
                                        uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %7, 0;\n\t"
               "tcgen05.mma.sp.cta_group::1.kind::f16 [%0], da, db, [%5], %6, p;\n\t}"
               ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(e_tmem), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_sp_2sm(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t e_tmem,
                                            uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %7, 0;\n\t"
               "tcgen05.mma.sp.cta_group::2.kind::f16 [%0], da, db, [%5], %6, p;\n\t}"
               ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(e_tmem), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, "
      "%27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// byte offset of fp16 element (row r, k) inside a 128-byte-swizzled tile of 128-byte rows
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t k) {
  return r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2;
}

constexpr int kPats = 6;
__constant__ uint32_t c_pats[kPats] = {0x4, 0x8, 0xC, 0x9, 0xD, 0xE};  // (0,1) (0,2) (0,3) (1,2) (1,3) (2,3): idx0 | idx1 << 2

// trial t: what metadata goes where and how the MMA addresses it
struct Trial { int kind, col_off, id2; };

__global__ void __launch_bounds__(128, 1) decode_kernel(float* out, int n_trials, const Trial* trials) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base, slot = base + 16, sA = base + 1024, sB = sA + 16 * 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  // A: compressed row m = 1..16 in its first 32 bytes (16 stored fp16 = 32 logical K), rest 0; B = identity 32x32 (K-major)
  for (uint32_t i = threadIdx.x; i < 16 * 1024 / 4; i += 128) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sA + i * 4), "r"(0u) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sB + i * 4), "r"(0u) : "memory");
  }
  __syncthreads();
  for (int j = 0; j < 16; ++j) {
    const __half v = __float2half((float)(j + 1));
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(sA + sw128(row, j)), "h"(*reinterpret_cast<const unsigned short*>(&v)) : "memory");
  }
  if (row < 32) {
    const __half one = __float2half(1.0f);
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(sB + sw128(row, row)), "h"(*reinterpret_cast<const unsigned short*>(&one)) : "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t ecol = 256;  // metadata columns 256..263
  uint32_t parity = 0;
  for (int t = 0; t < n_trials; ++t) {
    const Trial tr = trials[t];
    uint32_t w[4];
    switch (tr.kind) {
      case 0: w[0] = 0x44444444u; w[1] = 0xEEEEEEEEu; w[2] = 0x88888888u; w[3] = 0x99999999u; break;
      case 1: w[0] = 0xEEEEEEEEu; w[1] = 0x44444444u; w[2] = 0x99999999u; w[3] = 0x88888888u; break;
      case 2:  // nibble j of lane L = pats[(L + j) % 6]
        w[0] = 0;
        for (int j = 0; j < 8; ++j) w[0] |= c_pats[(row + j) % kPats] << (4 * j);
        w[1] = w[2] = w[3] = 0x44444444u;
        break;
      default:  // low half = pats[L % 6] x4, high half = pats[(L / 6) % 6] x4
        w[0] = c_pats[row % kPats] * 0x1111u | (c_pats[(row / kPats) % kPats] * 0x1111u) << 16;
        w[1] = w[2] = w[3] = 0x44444444u;
        break;
    }
    for (int c = 0; c < 4; ++c) tmem_st1(tmem + lane_base + ecol + c, w[c]);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      if (elect_one()) {
        umma_sp(tmem, sdesc_lo(sA), sdesc_hi(1024), sdesc_lo(sB), sdesc_hi(1024), tmem + ecol + tr.col_off, idesc_sp(128, 32, tr.id2), 0u);
        umma_commit(bar);
      }
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    tc_fence_after();
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + lane_base, v);
    tmem_ld_wait();
    for (int n = 0; n < 32; ++n) out[((size_t)t * 128 + row) * 32 + n] = __uint_as_float(v[n]);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------- part 2: timing
struct TP { int n, iters, sparse, pair; long long* cyc; };

template <bool PAIR>
__global__ void __launch_bounds__(128, 1) time_kernel(TP p) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base, slot = base + 16, sA = base + 1024, sB = sA + 4 * 16 * 1024;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { if (PAIR) tmem_alloc_2sm(slot, 512); else tmem_alloc(slot, 512); }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  // valid metadata (pattern (0,1) everywhere) in columns 256..271 of every lane
  for (int c = 0; c < 16; ++c) tmem_st1(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c, 0x44444444u);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  if (warp == 1 && rank == 0) {
    const uint32_t m = PAIR ? 256u : 128u;
    const uint32_t idesc_d = (1u << 4) | (((uint32_t)p.n >> 3) << 17) | ((m >> 4) << 24);
    const uint32_t a_hi = sdesc_hi(1024), b_hi = sdesc_hi(1024);
    const long long t0 = clock64();
    uint32_t st = 0;
    for (int it = 0; it < p.iters; ++it) {
      const uint32_t a_lo = sdesc_lo(sA + st * 16 * 1024), b_lo = sdesc_lo(sB + st * 32 * 1024);
      if (elect_one()) {
        if (p.sparse) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // A row of 128 B = 128 logical K = 4 sparse MMAs; B rows: two 64-channel tiles
            const uint32_t bk = b_lo + (ks >> 1) * (PAIR ? 1024 : 1024) + (ks & 1) * 4;
            if (PAIR) umma_sp_2sm(tmem, a_lo + 2 * ks, a_hi, bk, b_hi, tmem + 256 + (ks & 2), idesc_sp(m, p.n, ks & 1), 1u);
            else umma_sp(tmem, a_lo + 2 * ks, a_hi, bk, b_hi, tmem + 256 + (ks & 2), idesc_sp(m, p.n, ks & 1), 1u);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (PAIR) umma_f16_ss_lohi_2sm(tmem, a_lo + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc_d, 1u);
            else umma_f16_ss_lohi(tmem, a_lo + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc_d, 1u);
          }
        }
      }
      if (++st == 2) st = 0;
    }
    if (elect_one()) { if (PAIR) umma_commit_2sm(bar); else umma_commit(bar); }
    mbar_wait(bar, 0);
    if ((threadIdx.x & 31) == 0) p.cyc[blockIdx.x] = clock64() - t0;
  } else if (PAIR && warp == 1) {
    mbar_wait(bar, 0);  // the multicast commit also arrives here
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 0) { if (PAIR) tmem_dealloc_2sm(tmem, 512); else tmem_dealloc(tmem, 512); }
}

// usage: sp_probe decode <kind> <col_off> <id2> [n] | sp_probe time <pair> <sparse> <n>
// (one experiment per process: an illegal-instruction fault poisons the context)
int main(int argc, char** argv) {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  if (argc >= 5 && argv[1][0] == 'd') {
    Trial tr = {atoi(argv[2]), atoi(argv[3]), atoi(argv[4])};
    Trial* dtr;
    float* dout;
    CK(cudaMalloc(&dtr, sizeof tr));
    CK(cudaMemcpy(dtr, &tr, sizeof tr, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dout, (size_t)128 * 32 * 4));
    CK(cudaMemset(dout, 0xFF, (size_t)128 * 32 * 4));
    CK(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    decode_kernel<<<1, 128, 40 * 1024>>>(dout, 1, dtr);
    CK(cudaDeviceSynchronize());
    std::vector<float> h((size_t)128 * 32);
    CK(cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost));
    printf("== decode: kind %d  e-column offset %d  id2 %d   (row: 32 logical K positions; value v = stored element v-1)\n", tr.kind,
           tr.col_off, tr.id2);
    const int nrows = tr.kind >= 2 ? 128 : 20;
    for (int r = 0; r < nrows; ++r) {
      printf("r%03d:", r);
      for (int n = 0; n < 32; ++n) printf("%s%2d", (n % 4 == 0) ? " |" : " ", (int)h[(size_t)r * 32 + n]);
      printf("\n");
    }
    return 0;
  }
  if (argc >= 5 && argv[1][0] == 't') {
    const int pair = atoi(argv[2]), sparse = atoi(argv[3]), n = atoi(argv[4]);
    long long* dc;
    CK(cudaMalloc(&dc, 8192 * 8));
    CK(cudaFuncSetAttribute(time_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(time_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TP p;
    p.n = n; p.iters = 1000; p.sparse = sparse; p.pair = pair; p.cyc = dc;
    cudaLaunchConfig_t cfg = {};
    const int grid = pair ? (sms / 2) * 2 : sms;
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = 2048 + 2 * 16 * 1024 + 2 * 32 * 1024 + 32 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = pair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (pair) CK(cudaLaunchKernelEx(&cfg, time_kernel<true>, p)); else CK(cudaLaunchKernelEx(&cfg, time_kernel<false>, p));
    CK(cudaDeviceSynchronize());
    std::vector<long long> c(grid);
    CK(cudaMemcpy(c.data(), dc, grid * 8, cudaMemcpyDeviceToHost));
    double tot = 0; int cnt = 0;
    for (int i = 0; i < grid; i += pair ? 2 : 1) { tot += c[i]; ++cnt; }
    printf("time: pair %d sparse %d N %3d | %8.1f cycles per MMA (dense K=16, sparse K=32 logical)\n", pair, sparse, n,
           tot / cnt / (p.iters * 4.0));
    return 0;
  }
  printf("usage: sp_probe decode <kind> <col_off> <id2> | sp_probe time <pair> <sparse> <n>\n");
  return 2;
}
