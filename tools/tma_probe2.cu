// TMA issue-rate probe, v2: what bounds cp.async.bulk.tensor throughput from ONE CTA?
//   - box rows swept 8..256 (is the cost per instruction, per row or per byte?)
//   - 1/2/4 producer warps per CTA, each with its own ring (is the cap per issuing thread?)
//   - "nowait" mode: issue without any consumer handshake (raw issue cost of the instruction)
//   - L2 promotion NONE/128/256, swizzle none/128B, 2-D vs 4-D maps
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_probe2 tools/tma_probe2.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

struct Params {
  CUtensorMap tm;
  int box_bytes, slot_bytes, stages, iters, nprod, nowait, rank2;
  int bw, bh, tiles_w, tiles_h, n_img;
  long long* cycles;   // [grid][2]: total, issue-only
  const CUtensorMap* gdesc;  // descriptor copy in global memory (nullptr = use the kernel-parameter copy)
  int prefetch;
};

// warps 0..nprod-1: producers (lane 0), warps nprod..2*nprod-1: their consumers
__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ Params p) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ring = warp % p.nprod;
  const uint32_t bar_full = base + ring * 128, bar_empty = bar_full + 64;
  const uint32_t data = base + 1024 + ring * p.stages * p.slot_bytes;
  if (threadIdx.x == 0) {
    for (int r = 0; r < p.nprod; ++r)
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(base + r * 128 + 8 * s, p.nowait ? p.iters / p.stages : 1);  // nowait: one phase = all loads of the slot
        mbar_init(base + r * 128 + 64 + 8 * s, 1);
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const CUtensorMap* tm = p.gdesc ? p.gdesc : &p.tm;
  if (p.prefetch && threadIdx.x == 0) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
  const int n_tiles = p.tiles_w * p.tiles_h * p.n_img;
  long long t0 = clock64(), t_issue = 0;
  if (warp < p.nprod) {
    if (lane == 0) {
      int tile = (blockIdx.x * p.nprod + ring) % n_tiles; uint32_t s = 0, ph = 0;
      for (int it = 0; it < p.iters; ++it) {
        const int img = tile / (p.tiles_w * p.tiles_h), r = tile % (p.tiles_w * p.tiles_h);
        if (!p.nowait) mbar_wait(bar_empty + 8 * s, ph ^ 1);
        mbar_expect_tx(bar_full + 8 * s, p.box_bytes);
        if (p.rank2) tma_load_2d(data + s * p.slot_bytes, tm, bar_full + 8 * s, 0, tile * p.bw * p.bh);
        else tma_load_4d(data + s * p.slot_bytes, tm, bar_full + 8 * s, 0, (r % p.tiles_w) * p.bw, (r / p.tiles_w) * p.bh, img);
        if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
        tile += gridDim.x * p.nprod; if (tile >= n_tiles) tile -= n_tiles;
      }
      t_issue = clock64() - t0;
      if (p.nowait) {  // drain: every slot's barrier completes iters/stages phases
        for (int s2 = 0; s2 < p.stages; ++s2) mbar_wait(bar_full + 8 * s2, 0);
      }
    }
  } else if (warp < 2 * p.nprod && lane == 0 && !p.nowait) {
    uint32_t s = 0, ph = 0;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(bar_full + 8 * s, ph);
      mbar_arrive(bar_empty + 8 * s);
      if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) { p.cycles[blockIdx.x * 2] = clock64() - t0; p.cycles[blockIdx.x * 2 + 1] = t_issue; }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t buf_bytes = (size_t)64 << 20;  // L2-resident working set
  uint8_t* buf; CK(cudaMalloc(&buf, buf_bytes)); CK(cudaMemset(buf, 1, buf_bytes));
  long long* dcyc; CK(cudaMalloc(&dcyc, 4096 * 8));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  printf("%-34s %5s %4s %6s | %10s %10s %9s\n", "config", "nprod", "stg", "nowait", "cyc/instr", "issue/instr", "B/clk/SM");
  struct Cfg { const char* name; int bw, bh, rank2, swz, promo; };
  const Cfg cfgs[] = {
      {"4D box 16x4  (64 rows)", 16, 4, 0, 1, 2},  {"4D box 16x8  (128 rows)", 16, 8, 0, 1, 2},
      {"2D box 64x128 (contiguous 16KB)", 16, 8, 1, 1, 2},
  };
  const int W = 320, H = 320, C = 64;
  const int n_img = (int)(buf_bytes / ((size_t)W * H * C * 2));
  for (const Cfg& c : cfgs) {
    Params p;
    CUresult r;
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUtensorMapL2promotion promo = c.promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : c.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    const CUtensorMapSwizzle swz = c.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    if (c.rank2) {
      cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)W * H * n_img};
      cuuint64_t st[1] = {(cuuint64_t)C * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)(c.bw * c.bh)};
      r = enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
      cuuint64_t st[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
      cuuint32_t box[4] = {64, (cuuint32_t)c.bw, (cuuint32_t)c.bh, 1};
      r = enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, buf, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("%-34s encode failed %d\n", c.name, (int)r); continue; }
    p.box_bytes = c.bw * c.bh * 128; p.slot_bytes = (p.box_bytes + 1023) / 1024 * 1024;
    p.bw = c.bw; p.bh = c.bh; p.tiles_w = W / c.bw; p.tiles_h = H / c.bh; p.n_img = n_img; p.rank2 = c.rank2; p.cycles = dcyc;
    CUtensorMap* gd; CK(cudaMalloc(&gd, 256)); CK(cudaMemcpy(gd, &p.tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    for (int loc : {0, 1, 2})
    for (int nprod : {1, 2, 3})
      for (int nowait : {0, 1}) {
        p.gdesc = loc ? gd : nullptr; p.prefetch = loc == 2;
        const int stages = std::min(8, (200 * 1024 / nprod) / p.slot_bytes);
        if (stages < 2) continue;
        p.stages = stages; p.nprod = nprod; p.nowait = nowait; p.iters = 32 * stages;
        const int smem = 2048 + nprod * stages * p.slot_bytes;
        probe<<<sms, 256, smem>>>(p);
        probe<<<sms, 256, smem>>>(p);
        CK(cudaDeviceSynchronize());
        std::vector<long long> cyc(sms * 2);
        CK(cudaMemcpy(cyc.data(), dcyc, sms * 16, cudaMemcpyDeviceToHost));
        double tot = 0, iss = 0; for (int i = 0; i < sms; ++i) { tot += cyc[2 * i]; iss += cyc[2 * i + 1]; }
        tot /= sms; iss /= sms;
        printf("%-34s %s %5d %4d %6d | %10.1f %10.1f %9.1f\n", c.name, loc == 0 ? "param " : loc == 1 ? "global" : "gl+pre", nprod, stages, nowait, tot / p.iters, iss / p.iters,
               (double)p.iters * nprod * p.box_bytes / tot);
      }
  }
  return 0;
}
