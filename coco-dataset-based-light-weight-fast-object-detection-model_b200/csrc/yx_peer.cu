// CUDA IPC plumbing for the fused detection gather (include/yolox_b200.h, "multi-GPU" section): export the handle of a
// caller-owned device allocation, map a peer's allocation into this process.  The data path itself is the tail of
// nms_kernel in yx_post.cu.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "yx_internal.h"

using namespace yx;

static_assert(sizeof(cudaIpcMemHandle_t) == YX_IPC_HANDLE_BYTES, "handle size");

extern "C" int yx_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out) {
  YX_REQUIRE(dev_ptr && handle_out && offset_out, "null argument");
  // the handle names the whole cudaMalloc block: find its base (the caching allocator hands out interior pointers)
  typedef CUresult (*range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static range_fn get_range = nullptr;
  if (!get_range) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !ptr) {
      set_error("cuMemGetAddressRange entry point unavailable");
      return YX_ERR_CUDA;
    }
    get_range = reinterpret_cast<range_fn>(ptr);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (get_range(&base, &size, (CUdeviceptr)(uintptr_t)dev_ptr) != CUDA_SUCCESS) {
    set_error("cuMemGetAddressRange failed (not a device pointer?)");
    return YX_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  YX_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = (int64_t)((uintptr_t)dev_ptr - (uintptr_t)base);
  return YX_OK;
}

extern "C" int yx_ipc_open(const void* handle, void** base_out) {
  YX_REQUIRE(handle && base_out, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  YX_CUDA(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
  return YX_OK;
}

extern "C" int yx_ipc_close(void* base) {
  YX_REQUIRE(base != nullptr, "null argument");
  YX_CUDA(cudaIpcCloseMemHandle(base));
  return YX_OK;
}
