// Memory-path probe for the tiled compositing kernel: a persistent-CTA tile copy with the same
// structure (TMA box loads into a ring, eight warps streaming rows out with 16-byte stores, cells
// taken round-robin, frames innermost in blocks) but no resampling, to see what the access pattern
// itself can sustain for different cell / box shapes and frame blocks.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tile_copy_probe tile_copy_probe.cu
//   ./tile_copy_probe cell_w_bytes cell_h box_w_bytes box_h frame_block frames stages [ctas_per_sm] [store_mode]
// store_mode 0: LDS.128 + STG.128 (streaming), 1: no stores at all (loads only)
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

struct Args {
    CUtensorMap tmap;
    uint8_t* dst;
    long long pitch, fstride;
    int cells_x, cells_y, cw, ch, bw, bh, fb, frames, stages, box_bytes, store_mode, sx_step, sy_step16;
};

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

extern __shared__ __align__(128) uint8_t smem[];

__global__ void __launch_bounds__(256, 2) probe(const __grid_constant__ Args a) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = smem_u32(smem);
    const uint32_t full = base + a.stages * a.box_bytes, empty = full + 64;
    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full + 8 * s), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty + 8 * s), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int n_cells = a.cells_x * a.cells_y;
    const int n_blocks = (a.frames + a.fb - 1) / a.fb;
    // unit sequence of this CTA: for blk, for cell = blockIdx.x + k * grid, for f in block
    // issuer state
    int i_blk = 0, i_cell = blockIdx.x, i_f = 0, i_slot = 0;
    uint32_t i_phase = 0;
    bool i_active = i_cell < n_cells;
    auto issue = [&]() {
        if (!i_active) return;
        mbar_wait(empty + 8 * i_slot, i_phase ^ 1);
        const int cx = i_cell % a.cells_x, cy = i_cell / a.cells_x;
        const int frame = i_blk * a.fb + i_f;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * i_slot), "r"(a.box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(base + i_slot * a.box_bytes), "l"(reinterpret_cast<uint64_t>(&a.tmap)), "r"((cx * a.sx_step / 16) * 4), "r"((cy * a.sy_step16) >> 4),
                       "r"(frame), "r"(full + 8 * i_slot) : "memory");
        if (++i_slot == a.stages) { i_slot = 0; i_phase ^= 1; }
        const int nf = min(a.fb, a.frames - i_blk * a.fb);
        if (++i_f == nf) {
            i_f = 0;
            i_cell += gridDim.x;
            if (i_cell >= n_cells) {
                i_cell = blockIdx.x;
                if (++i_blk >= n_blocks) i_active = false;
            }
        }
    };
    if (tid == 0) for (int i = 0; i < a.stages - 2; ++i) issue();
    int slot = 0;
    uint32_t phase = 0;
    for (int blk = 0; blk < n_blocks; ++blk) {
        const int nf = min(a.fb, a.frames - blk * a.fb);
        for (int cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
            const int cx = cell % a.cells_x, cy = cell / a.cells_x;
            for (int f = 0; f < nf; ++f) {
                if (tid == 0) issue();
                mbar_wait(full + 8 * slot, phase);
                const uint32_t box = base + slot * a.box_bytes;
                uint8_t* out = a.dst + (long long)(blk * a.fb + f) * a.fstride + (long long)(cy * a.ch) * a.pitch + cx * a.cw;
                for (int r = warp; r < a.ch; r += 8) {
                    for (int c = lane * 16; c < a.cw; c += 512) {
                        uint4 v;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(box + r * a.bw + c));
                        if (a.store_mode == 0) __stcs(reinterpret_cast<uint4*>(out + (long long)r * a.pitch + c), v);
                        else if (v.x == 0x12345678u) out[0] = 1;
                    }
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty + 8 * slot) : "memory");
                if (++slot == a.stages) { slot = 0; phase ^= 1; }
            }
        }
    }
}

int main(int argc, char** argv) {
    if (argc < 8) { printf("usage: cw ch bw bh fb frames stages [ctas_per_sm] [store_mode]\n"); return 1; }
    Args a; memset(&a, 0, sizeof(a));
    a.cw = atoi(argv[1]); a.ch = atoi(argv[2]); a.bw = atoi(argv[3]); a.bh = atoi(argv[4]);
    a.fb = atoi(argv[5]); a.frames = atoi(argv[6]); a.stages = atoi(argv[7]);
    const int per_sm = argc > 8 ? atoi(argv[8]) : 2;
    a.store_mode = argc > 9 ? atoi(argv[9]) : 0;
    a.sx_step = argc > 10 ? atoi(argv[10]) : a.cw;          // source bytes between the boxes of neighbouring cells
    a.sy_step16 = argc > 11 ? atoi(argv[11]) : a.ch * 16;   // source rows (x16) between vertically neighbouring cells
    const int W = 22400 / a.cw * a.cw, H = 1120 / a.ch * a.ch;   // one "panorama" of ~25 MB
    a.cells_x = W / a.cw; a.cells_y = H / a.ch;
    a.pitch = 22400 + 512; a.fstride = a.pitch * (H + 64);
    a.box_bytes = (a.bw * a.bh + 127) & ~127;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    uint8_t *src, *dst;
    CK(cudaMalloc(&src, a.fstride * a.frames)); CK(cudaMalloc(&dst, a.fstride * a.frames));
    CK(cudaMemset(src, 1, a.fstride * a.frames));
    a.dst = dst;
    cuuint64_t dims[3] = {(cuuint64_t)(a.pitch / 4), (cuuint64_t)(H + 64), (cuuint64_t)a.frames};
    cuuint64_t strides[2] = {(cuuint64_t)a.pitch, (cuuint64_t)a.fstride};
    cuuint32_t box[3] = {(cuuint32_t)(a.bw / 4), (cuuint32_t)a.bh, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&a.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const size_t smem_pad = argc > 12 ? (size_t)atoi(argv[12]) * 1024 : 0;   // extra dynamic shared memory (KB)
    const size_t smem_bytes = (size_t)a.stages * a.box_bytes + 256 + smem_pad;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    const int grid = 148 * per_sm;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) probe<<<grid, 256, smem_bytes>>>(a);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int i = 0; i < reps; ++i) probe<<<grid, 256, smem_bytes>>>(a);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    const double useful = (double)W * H * a.frames;
    const double loaded = (double)a.bw * a.bh * a.cells_x * a.cells_y * a.frames;
    printf("sx %d sy16 %d ", a.sx_step, a.sy_step16); printf("cell %4dx%-3d box %4dx%-3d fb %3d frames %3d stages %d ctas/sm %d store %d smem %6zu : %.3f ms  copy %.0f GB/s (r+w useful)  box-load %.0f GB/s\n",
           a.cw, a.ch, a.bw, a.bh, a.fb, a.frames, a.stages, per_sm, a.store_mode, smem_bytes, ms,
           (a.store_mode == 0 ? 2.0 : 1.0) * useful / ms / 1e6, loaded / ms / 1e6);
    return 0;
}
