// Minimal TMA probe: variants selected by argv[1].
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Args { CUtensorMap tmap[4]; uint32_t* out; const CUtensorMap* gmap; int rank; int c0, c1, c2; int bytes; int idx; int use_g; };

__global__ void probe(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(a.bytes) : "memory");
        const CUtensorMap* m = a.use_g ? a.gmap : &a.tmap[a.idx];
        if (a.rank == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(a.c0), "r"(a.c1), "r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(a.c0), "r"(a.c1), "r"(a.c2), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    } while (!done);
    for (int i = threadIdx.x; i < a.bytes / 4; i += blockDim.x) a.out[i] = reinterpret_cast<uint32_t*>(smem)[i];
}

int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int W4 = 192, H = 144, F = 3; const size_t pitch = W4 * 4;
    uint32_t* src; CK(cudaMalloc(&src, pitch * H * F));
    uint32_t* hsrc = (uint32_t*)malloc(pitch * H * F);
    for (int i = 0; i < W4 * H * F; ++i) hsrc[i] = i;
    CK(cudaMemcpy(src, hsrc, pitch * H * F, cudaMemcpyHostToDevice));
    Args a; memset(&a, 0, sizeof(a));
    int bw = 100, bh = 20;
    int rank = (variant & 1) ? 2 : 3;
    CUtensorMapDataType dt = (variant & 2) ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    int esz = (variant & 2) ? 1 : 4;
    if (variant & 4) { bw = 64; bh = 16; }
    if (variant & 2) bw = (variant & 4) ? 256 : 208;
    cuuint64_t dims[3] = {(cuuint64_t)(W4 * 4 / esz), H, F};
    cuuint64_t strides[2] = {pitch, pitch * H};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&a.tmap[1], dt, rank, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d rank %d esz %d box %dx%d encode rc=%d\n", variant, rank, esz, bw, bh, (int)r);
    a.rank = rank; a.idx = 1; a.c0 = (variant & 8) ? -3 : 5; a.c1 = (variant & 8) ? -2 : 7; a.c2 = 1;
    a.bytes = bw * esz * bh;
    CK(cudaMalloc(&a.out, a.bytes));
    { CUtensorMap* g; CK(cudaMalloc(&g, 128)); CK(cudaMemcpy(g, &a.tmap[1], 128, cudaMemcpyHostToDevice)); a.gmap = g; a.use_g = (variant & 16) ? 1 : 0; }
    { int drv=0, rt=0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0); printf("  driver %d runtime %d cc %d.%d %s\n", drv, rt, pr.major, pr.minor, pr.name); }
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    probe<<<1, 128, 65536>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        uint32_t* h = (uint32_t*)malloc(a.bytes);
        CK(cudaMemcpy(h, a.out, a.bytes, cudaMemcpyDeviceToHost));
        printf("  first words: %u %u %u %u (expect word index of (c0,c1,frame))\n", h[0], h[1], h[2], h[3]);
    }
    return 0;
}
