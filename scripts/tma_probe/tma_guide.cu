// CUDA programming guide TMA example (libcu++ wrappers) morphing toward the failing raw-PTX use.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdlib.h>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int bytes, int* out) {
  __shared__ alignas(128) int smem_buffer[8192];
  #pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = smem_buffer[i];
}
int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaFree(0));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)p;
  int W = (variant & 4) ? 192 : 256, H = 256;
  int bw = (variant & 2) ? 64 : 32, bh = (variant & 2) ? 16 : 32;
  if (variant & 16) { bw = 100; bh = 20; }
  int* src; CK(cudaMalloc(&src, W * H * 4));
  int* h = (int*)malloc(W * H * 4); for (int i = 0; i < W * H; ++i) h[i] = i;
  CK(cudaMemcpy(src, h, W * H * 4, cudaMemcpyHostToDevice));
  CUtensorMap tm{};
  cuuint64_t size[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t stride[1] = {(cuuint64_t)W * 4};
  cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, (variant & 1) ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_INT32, 2, src, size, stride, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   (variant & 8) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d W %d box %dx%d encode rc=%d\n", variant, W, bw, bh, (int)r);
  int bytes = bw * bh * 4;
  int* out; CK(cudaMalloc(&out, bytes));
  int cx = (variant & 32) ? 5 : 64, cy = (variant & 32) ? 7 : 32;
  kernel<<<1, 128>>>(tm, cx, cy, bytes, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  guide kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) { int ho[4]; CK(cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost)); printf("  first: %d %d (expect %d)\n", ho[0], ho[1], cy * W + cx); }
  return 0;
}
