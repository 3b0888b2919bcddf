// tcgen05.mma issue/execute-rate probe on B200: cycles per UTCHMMA (M=128, K=16, fp16 SS-mode, cta_group::1) as a
// function of N, of the A-descriptor geometry (1024-B aligned atoms vs the halo kernel's 128-B aligned starts with
// SBO = 1280 B) and of how many accumulators are alternated.  Operands are whatever is in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_probe tools/mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void umma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
               ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t a) { return ((a & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t sdesc_hi(uint32_t sbo) { return (sbo >> 4) | (1u << 14) | (2u << 29); }

struct P { int n, iters, a_sbo, a_off, n_acc, stages, ksteps, fence, spin_warps, ld_warps, st_warps; long long* cyc; };

__global__ void __launch_bounds__(512, 1) probe(P p) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base, never = base + 8, done = base + 32, slot = base + 16, sA = base + 1024, sB = sA + 4 * 24 * 1024;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(never, 1); asm volatile("st.shared.u32 [%0], %1;" ::"r"(done), "r"(0)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (((uint32_t)p.n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hi = sdesc_hi(p.a_sbo), b_hi = sdesc_hi(1024);
    const long long t0 = clock64();
    uint32_t st = 0, acc = 0;
    for (int it = 0; it < p.iters; ++it) {
      const uint32_t stg = p.ksteps >= 9 ? 0u : st;   // (the tap patterns span more than one 24 KB stage)
      const uint32_t a_lo = sdesc_lo(sA + stg * 24 * 1024 + p.a_off), b_lo = sdesc_lo(sB + stg * 24 * 1024);
      const uint32_t d = tmem + acc * (p.n_acc > 2 ? 128 : 256);
      if (p.fence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        if (p.ksteps == 9) {   // "taps" pattern of the stem: nine K = 16 MMAs with nine different A / B tile bases, straight-line
#pragma unroll
          for (int t = 0; t < 9; ++t) umma(d, a_lo + ((t / 3) * 10 + (t % 3)) * 8, a_hi, b_lo + t * 384, b_hi, idesc, 1u);
        } else if (p.ksteps == 18) {   // the same with two accumulators per tap (two stacked halves)
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            umma(d, a_lo + ((t / 3) * 10 + (t % 3)) * 8, a_hi, b_lo + t * 384, b_hi, idesc, 1u);
            umma(d + 64, a_lo + 1280 + ((t / 3) * 10 + (t % 3)) * 8, a_hi, b_lo + t * 384, b_hi, idesc, 1u);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) if (ks < p.ksteps) umma(d, a_lo + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, 1u);
        }
      }
      if (++st == (uint32_t)p.stages) st = 0;
      if (++acc == (uint32_t)p.n_acc) acc = 0;
    }
    const long long t_issue = clock64() - t0;
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    mbar_wait(bar, 0);
    if ((threadIdx.x & 31) == 0) { p.cyc[blockIdx.x * 2] = clock64() - t0; p.cyc[blockIdx.x * 2 + 1] = t_issue; }
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(done), "r"(1) : "memory");
  } else if (warp >= 4 && warp < 4 + p.ld_warps) {
    // "epilogue": tcgen05.ld of the upper TMEM half in a loop until the MMA warp is done
    uint32_t flag = 0, sink = 0;
    while (!flag) {
      for (int c = 0; c < 8; ++c) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem + 256 + c * 16 + (((uint32_t)(warp & 3) * 32) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) sink ^= v[i];
        if (p.st_warps) asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(sA + 200 * 1024 + (threadIdx.x & 127) * 16), "r"(sink) : "memory");
      }
      asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(flag) : "r"(done));
    }
    if (sink == 0x12345) p.cyc[4000] = sink;
  } else if (warp >= 8 && warp < 8 + p.spin_warps) {
    // blocked producers / epilogue warps: all lanes poll an mbarrier that never completes, until the MMA warp is done
    uint32_t flag = 0;
    while (!flag) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(never), "r"(0) : "memory");
      asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(flag) : "r"(done));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main(int argc, char** argv) {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d; CK(cudaMalloc(&d, 8192 * 8));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  printf("%5s %6s %6s %5s %5s %5s %5s | %12s %12s %10s\n", "N", "ksteps", "fence", "spin", "ld", "st", "geom", "cyc/MMA", "issue/MMA", "ideal N/2");
  if (argc > 1) {   // dependent-chain sweep: small N, 1 or 4 K steps per accumulator visit, 1 / 2 / 4 accumulators alternated
    for (int n : {48, 96, 192})
      for (int ksteps : {1, 4, 9, 18})
        for (int n_acc : {1, 2, 4}) {
          if (n_acc * 256 > 512 && n > 64) continue;
          P p; p.n = n; p.iters = 4000; p.n_acc = n_acc; p.stages = 4; p.cyc = d;
          p.ksteps = ksteps; p.fence = 0; p.spin_warps = 0; p.ld_warps = 0; p.st_warps = 0; p.a_sbo = 1280; p.a_off = 11 * 128;
          probe<<<sms, 512, 2048 + 8 * 24 * 1024 + 8 * 1024>>>(p);
          CK(cudaDeviceSynchronize());
          std::vector<long long> c(sms * 2);
          CK(cudaMemcpy(c.data(), d, sms * 16, cudaMemcpyDeviceToHost));
          double tot = 0, iss = 0; for (int i = 0; i < sms; ++i) { tot += c[2 * i]; iss += c[2 * i + 1]; }
          printf("N %3d  ksteps/visit %d  accumulators %d : %7.1f cycles per MMA (issue %6.1f)\n", n, ksteps, n_acc,
                 tot / sms / (p.iters * (double)ksteps), iss / sms / (p.iters * (double)ksteps));
        }
    return 0;
  }
  struct M { int ksteps, fence, spin, ld, st; };
  const M modes[] = {{4, 0, 0, 0, 0}, {3, 1, 0, 0, 0}, {4, 1, 4, 0, 0}, {4, 1, 8, 0, 0}, {4, 1, 0, 4, 0}, {4, 1, 0, 4, 1}, {4, 1, 4, 4, 1}, {3, 1, 8, 4, 1}};
  for (int n : {48, 96, 192})
    for (const M& m : modes)
      for (int geom = 0; geom < 3; geom += 2) {
        {
          const int n_acc = 1, stages = 4;
          P p; p.n = n; p.iters = 2000; p.n_acc = n_acc; p.stages = stages; p.cyc = d;
          p.ksteps = m.ksteps; p.fence = m.fence; p.spin_warps = m.spin; p.ld_warps = m.ld; p.st_warps = m.st;
          p.a_sbo = geom == 0 ? 1024 : 1280; p.a_off = geom == 2 ? 11 * 128 : 0;  // geom 2: tap (1,1) of a halo tile
          probe<<<sms, 512, 2048 + 8 * 24 * 1024 + 8 * 1024>>>(p);  // 4 A + 4 B stages of 24 KB (+ slack for N = 256)
          CK(cudaDeviceSynchronize());
          std::vector<long long> c(sms * 2);
          CK(cudaMemcpy(c.data(), d, sms * 16, cudaMemcpyDeviceToHost));
          double tot = 0, iss = 0; for (int i = 0; i < sms; ++i) { tot += c[2 * i]; iss += c[2 * i + 1]; }
          printf("%5d %6d %6d %5d %5d %5d %5s | %12.1f %12.1f %10.1f\n", n, m.ksteps, m.fence, m.spin, m.ld, m.st, geom == 0 ? "align" : "halo",
                 tot / sms / (p.iters * (double)m.ksteps), iss / sms / (p.iters * (double)m.ksteps), n / 2.0);
        }
      }
  return 0;
}
