// 2:4 structured-sparse weights for the sparse tensor-core conv variant (tcgen05.mma.sp, yx_conv.cu SP = true).
//
// The reference ships pruned checkpoints as sparse COO tensors that are densified at load time
// (choijhanyangackr/main.py:52-55); masks come from 01_mask_generator.py:18-46.  When EVERY group of four consecutive
// input channels of a conv's folded weight holds at most two non-zeros (a 2:4-compliant mask), the layer can run with the
// weights as the SPARSE A operand of the MMA: half of the weight bytes, twice the K per instruction.
//
// This file checks compliance and packs, on the device, from the engine's KRSC fp16 blob [cout_pad][taps][cin_pad]:
//   * compressed weights  [cout_pad128][taps][cin_pad / 2] fp16: the two kept values of every group, in position order
//     (cout rows beyond cout_pad are zero, so a 128-row M tile never reads outside the tensor);
//   * metadata            [n_mt][cols_per_tile][128] uint32, one TMEM column per (M tile, tap, K = 32 step): column index
//     tap * (cin_pad / 32) + step (step = 2 * chunk + ks), lane / nibble placement as documented in yx_ptx.cuh.
#include <cuda_fp16.h>
#include <stdint.h>

#include <algorithm>

#include "yx_internal.h"

namespace yx {

// positions (ascending) of the two kept values of one group; groups with fewer than two non-zeros are completed with
// zero-valued positions (any valid ascending pair encodes the same logical row)
__device__ __forceinline__ bool pick_pair(const __half* w4, int* i0, int* i1) {
  int nz[4], n = 0;
  for (int i = 0; i < 4; ++i)
    if (__half2float(w4[i]) != 0.0f || (__half_as_ushort(w4[i]) & 0x7fffu) != 0) nz[n++] = i;   // NaN counts as non-zero
  if (n > 2) return false;
  if (n == 2) { *i0 = nz[0]; *i1 = nz[1]; }
  else if (n == 1) { if (nz[0] < 3) { *i0 = nz[0]; *i1 = 3; } else { *i0 = 2; *i1 = 3; } }
  else { *i0 = 0; *i1 = 1; }
  return true;
}

// one thread per (cout row of the padded-to-128 tensor, tap, group of four input channels)
__global__ void sparse_pack_kernel(const __half* __restrict__ w, int cout_pad, int taps, int cin_pad, int cout128,
                                   __half* __restrict__ wc, int* __restrict__ bad) {
  const int groups = cin_pad >> 2;
  const int64_t total = (int64_t)cout128 * taps * groups;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const int64_t rt = i / groups;   // row * taps + tap
    const int row = (int)(rt / taps);
    __half w4[4];
    if (row < cout_pad) {
      const __half* src = w + rt * cin_pad + 4 * g;
      for (int k = 0; k < 4; ++k) w4[k] = src[k];
    } else {
      for (int k = 0; k < 4; ++k) w4[k] = __ushort_as_half(0);
    }
    int i0 = 0, i1 = 1;
    if (!pick_pair(w4, &i0, &i1)) { atomicAdd(bad, 1); continue; }
    __half* dst = wc + rt * (cin_pad >> 1) + 2 * g;
    dst[0] = w4[i0];
    dst[1] = w4[i1];
  }
}

// one thread per metadata word: (M tile, column, lane)
__global__ void sparse_meta_kernel(const __half* __restrict__ w, int cout_pad, int taps, int cin_pad, int n_mt,
                                   uint32_t* __restrict__ meta) {
  const int spt = cin_pad >> 5;   // K = 32 steps per tap
  const int cols = taps * spt;
  const int64_t total = (int64_t)n_mt * cols * 128;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int L = (int)(i & 127);
    const int col = (int)((i >> 7) % cols), mt = (int)((i >> 7) / cols);
    const int tap = col / spt, step = col % spt;
    const int kh = (L >> 3) & 1;
    uint32_t word = 0;
    for (int half = 0; half < 2; ++half) {
      const int m = (L & 7) + 16 * (L >> 4) + 8 * half;   // row of the 128-row M tile
      const int row = mt * 128 + m;
      for (int g = 0; g < 4; ++g) {
        const int k0 = step * 32 + kh * 16 + 4 * g;   // first input channel of the group
        int i0 = 0, i1 = 1;
        if (row < cout_pad && k0 + 3 < cin_pad) {
          __half w4[4];
          const __half* src = w + ((int64_t)row * taps + tap) * cin_pad + k0;
          for (int k = 0; k < 4; ++k) w4[k] = src[k];
          pick_pair(w4, &i0, &i1);   // non-compliant groups were already reported by sparse_pack_kernel
        }
        word |= (uint32_t)(i0 | (i1 << 2)) << (4 * (g + 4 * half));
      }
    }
    meta[i] = word;
  }
}

// Packs one conv's weights.  *compliant = 0 when some group of four holds more than two non-zeros (nothing usable is
// written then).  Host-synchronising (engine creation time).
int sparse_pack(const void* weights_krsc, int cout_pad, int taps, int cin_pad, void* wc_out, void* meta_out, int* scratch_dev,
                int* compliant, cudaStream_t st) {
  YX_REQUIRE(cin_pad % 32 == 0, "sparse pack: cin_pad must be a multiple of 32");
  const int cout128 = round_up(cout_pad, 128), n_mt = cout128 / 128;
  YX_CUDA(cudaMemsetAsync(scratch_dev, 0, 4, st));
  const int64_t n1 = (int64_t)cout128 * taps * (cin_pad >> 2);
  sparse_pack_kernel<<<(int)std::min<int64_t>((n1 + 255) / 256, 148 * 8), 256, 0, st>>>(
      static_cast<const __half*>(weights_krsc), cout_pad, taps, cin_pad, cout128, static_cast<__half*>(wc_out), scratch_dev);
  YX_CUDA(cudaGetLastError());
  const int64_t n2 = (int64_t)n_mt * taps * (cin_pad / 32) * 128;
  sparse_meta_kernel<<<(int)std::min<int64_t>((n2 + 255) / 256, 148 * 8), 256, 0, st>>>(
      static_cast<const __half*>(weights_krsc), cout_pad, taps, cin_pad, n_mt, static_cast<uint32_t*>(meta_out));
  YX_CUDA(cudaGetLastError());
  int bad = 0;
  YX_CUDA(cudaMemcpyAsync(&bad, scratch_dev, 4, cudaMemcpyDeviceToHost, st));
  YX_CUDA(cudaStreamSynchronize(st));
  *compliant = bad == 0 ? 1 : 0;
  return YX_OK;
}

}  // namespace yx
