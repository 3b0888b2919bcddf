// Microbenchmark: steady-state TMA tile-load / tile-store throughput per SM on B200 for the box shapes the
// conv kernels use (full 128-B rows, 96-B rows, overlapping row-packed rows, halo boxes, stride-2 parity views).
// Answers "what limits the small-channel layers": bytes, rows (requests) or alignment.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu -lcuda && tools/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
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
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"((uint64_t)m), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct Params {
  CUtensorMap tm;
  int box_bytes, slot_bytes, stages, iters;
  int tiles_w, tiles_h, n_img;  // tile grid
  int tw, th, cx0, cy0;         // tile pitch in the map's coordinates, origin shift (halo: -1)
  int chunks;                   // boxes along the channel dim per tile (each 64 elements)
  int store;                    // 0 = loads, 1 = stores
  long long* cycles;
};

__global__ void __launch_bounds__(64, 2) probe(const __grid_constant__ Params p) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar_full = base, bar_empty = base + 64, data = base + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.tiles_w * p.tiles_h * p.n_img;
  long long t0 = clock64();
  if (p.store) {
    if (warp == 0 && lane == 0) {
      int tile = blockIdx.x;
      for (int it = 0; it < p.iters; ++it) {
        const int img = tile / (p.tiles_w * p.tiles_h), r = tile % (p.tiles_w * p.tiles_h);
        for (int c = 0; c < p.chunks; ++c)
          tma_store_4d(&p.tm, data + (it % p.stages) * p.slot_bytes, c * 64, (r % p.tiles_w) * p.tw, (r / p.tiles_w) * p.th, img);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        tile += gridDim.x; if (tile >= n_tiles) tile -= n_tiles;
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp == 0) {
    if (lane == 0) {
      int tile = blockIdx.x; uint32_t s = 0, ph = 0;
      for (int it = 0; it < p.iters; ++it) {
        const int img = tile / (p.tiles_w * p.tiles_h), r = tile % (p.tiles_w * p.tiles_h);
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          mbar_expect_tx(bar_full + 8 * s, p.box_bytes);
          tma_load_4d(data + s * p.slot_bytes, &p.tm, bar_full + 8 * s, c * 64, (r % p.tiles_w) * p.tw + p.cx0, (r / p.tiles_w) * p.th + p.cy0, img);
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
        }
        tile += gridDim.x; if (tile >= n_tiles) tile -= n_tiles;
      }
    }
  } else if (lane == 0) {
    uint32_t s = 0, ph = 0;
    for (int it = 0; it < p.iters * p.chunks; ++it) {
      mbar_wait(bar_full + 8 * s, ph);
      mbar_arrive(bar_empty + 8 * s);
      if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Case {
  const char* name;
  int C, pitch;      // channels (map dim 0) and pixel pitch in elements
  int W, H;          // map dims 1, 2 (pixels)
  int wstride_mul;   // row pitch = pitch * W * wstride_mul (2 for parity views)
  int bw, bh;        // box pixels
  int tw, th, c0;    // tile pitch, origin shift
  int store;
};

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const size_t buf_bytes = (size_t)3 << 30;
  uint8_t* buf; CK(cudaMalloc(&buf, buf_bytes)); CK(cudaMemset(buf, 1, buf_bytes));
  long long* dcyc; CK(cudaMalloc(&dcyc, 1024 * 8));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const Case cases[] = {
      {"full rows C64 p64 box16x8", 64, 64, 320, 320, 1, 16, 8, 16, 8, 0, 0},
      {"full rows C192 p192 3 chunks", 192, 192, 160, 160, 1, 16, 8, 16, 8, 0, 0},
      {"C48 p48 (96B rows)", 48, 48, 320, 320, 1, 16, 8, 16, 8, 0, 0},
      {"C48 p64 (96B rows, aligned)", 48, 64, 320, 320, 1, 16, 8, 16, 8, 0, 0},
      {"C96 p96 2 chunks (128+64B)", 96, 96, 320, 320, 1, 16, 8, 16, 8, 0, 0},
      {"C32 p32 (64B rows)", 32, 32, 320, 320, 1, 16, 8, 16, 8, 0, 0},
      {"rowpack C64 p16 (overlapping 128B rows)", 64, 16, 640, 640, 1, 16, 8, 16, 8, 0, 0},
      {"halo C64 p64 box10x18", 64, 64, 320, 320, 1, 10, 18, 8, 16, -1, 0},
      {"halo C48 p48 box10x18", 48, 48, 320, 320, 1, 10, 18, 8, 16, -1, 0},
      {"halo C192 p192 box10x18 3 chunks", 192, 192, 160, 160, 1, 10, 18, 8, 16, -1, 0},
      {"parity C48 p96 (s2 view of 48ch)", 48, 96, 320, 320, 2, 16, 8, 16, 8, 0, 0},
      {"parity C64 p128", 64, 128, 320, 320, 2, 16, 8, 16, 8, 0, 0},
      {"parity C96 p192 2 chunks", 96, 192, 160, 160, 2, 16, 8, 16, 8, 0, 0},
      {"wide box C64 p64 box32x8 (256 rows)", 64, 64, 320, 320, 1, 32, 8, 32, 8, 0, 0},
      {"STORE C64 p64", 64, 64, 320, 320, 1, 16, 8, 16, 8, 0, 1},
      {"STORE C48 p48", 48, 48, 320, 320, 1, 16, 8, 16, 8, 0, 1},
      {"STORE C48 p64", 48, 64, 320, 320, 1, 16, 8, 16, 8, 0, 1},
      {"STORE C96 p96 2 chunks", 96, 96, 320, 320, 1, 16, 8, 16, 8, 0, 1},
      {"STORE C192 p192 3 chunks", 192, 192, 160, 160, 1, 16, 8, 16, 8, 0, 1},
      {"STORE C80 p80 2 chunks (head cls)", 80, 80, 160, 160, 1, 16, 8, 16, 8, 0, 1},
  };
  printf("SMs %d clock %d kHz\n", sms, khz);
  printf("%-44s %5s %4s %6s | %9s %9s %9s | %8s\n", "case", "ctas", "stg", "set", "B/clk/SM", "rows/clk", "usefulB/c", "GB/s");
  for (const Case& c : cases) {
    const size_t row_bytes = (size_t)c.pitch * 2 * c.W * c.wstride_mul;
    const size_t img_bytes = row_bytes * c.H;
    for (int big = 0; big < 2; ++big) {                       // 0: L2-resident working set, 1: DRAM-sized
      const int n_img = (int)std::max<size_t>(1, (big ? ((size_t)2 << 30) : ((size_t)48 << 20)) / img_bytes);
      cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)n_img};
      cuuint64_t st[3] = {(cuuint64_t)c.pitch * 2 * c.wstride_mul, row_bytes, img_bytes};
      if (c.pitch == 16) { dims[1] = c.W - 4; }              // row-packed view: pixel pitch 32 B, 64 "channels"
      cuuint32_t box[4] = {64, (cuuint32_t)c.bw, (cuuint32_t)c.bh, 1}, es[4] = {1, 1, 1, 1};
      Params p;
      CUresult r = enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, buf, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("%-44s encode failed %d\n", c.name, (int)r); continue; }
      const int chunks = (c.C + 63) / 64;
      p.box_bytes = c.bw * c.bh * 128; p.slot_bytes = (p.box_bytes + 1023) / 1024 * 1024;
      p.tiles_w = (int)dims[1] / c.tw; p.tiles_h = c.H / c.th; p.n_img = n_img;
      p.tw = c.tw; p.th = c.th; p.cx0 = c.c0; p.cy0 = c.c0; p.chunks = chunks; p.store = c.store; p.cycles = dcyc;
      for (int ctas = 1; ctas <= 2; ++ctas)
        for (int stages : {2, 4, 8}) {
          if (c.store && stages != 4) continue;
          const int smem = 2048 + stages * p.slot_bytes;
          if (smem * ctas > 220 * 1024) continue;
          p.stages = stages;
          p.iters = 400;
          const int grid = sms * ctas;
          // force the requested residency: pad smem so exactly `ctas` CTAs fit per SM
          const int smem_launch = std::max(smem, ctas == 1 ? 120 * 1024 : 80 * 1024);
          probe<<<grid, 64, smem_launch>>>(p);  // warm
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          probe<<<grid, 64, smem_launch>>>(p);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          std::vector<long long> cyc(grid);
          CK(cudaMemcpy(cyc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost));
          double avg = 0; for (long long v : cyc) avg += v; avg /= grid;
          const double boxes = (double)p.iters * chunks * ctas;  // per SM
          const double rows = boxes * c.bw * c.bh;
          const double useful = (double)p.iters * ctas * c.bw * c.bh * std::min(c.C, 64 * chunks) * 2.0;
          printf("%-44s %5d %4d %6s | %9.1f %9.3f %9.1f | %8.0f\n", c.name, ctas, stages, big ? "DRAM" : "L2", boxes * p.box_bytes / avg,
                 rows / avg, useful / avg, useful * sms / (ms * 1e-3) / 1e9);
        }
    }
  }
  return 0;
}
