// Shared-memory data-pipe probe (experiments, not part of the library): how many cycles does an SM need
// per warp-level load instruction for the access patterns the resampling loop could use?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_probe lds_probe.cu
// Every mode runs 2 CTAs x 8 warps per SM (the tiled kernel's residency), each warp ITERS trips of a loop whose
// body is UNROLL independent "pixels"; the report is SM cycles per pixel-row (one source row of one warp-pixel).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

// mode: see main()
template <int MODE>
__global__ void __launch_bounds__(256, 2) probe(uint32_t* out, int stride_b, int phase, int walk) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8192; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + warp * 2048;
    const uint32_t b = (uint32_t)(phase + lane * stride_b / 100);   // byte offset of this lane's tap window
    uint32_t accs[4] = {0u, 0u, 0u, 0u};
    for (int it = 0; it < ITERS; ++it) {
        const uint32_t rowbase = base + (((uint32_t)(it * walk) & 3u) << 9);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t o = rowbase + b + u * 2048 * 0 + 0;   // same row for the four "pixels": only the pipe matters
            uint32_t& acc = accs[u];
            if (MODE == 0) {   // three aligned words (current kernel)
                const uint32_t a = o & ~3u;
                acc += lds32(a) + lds32(a + 4) + lds32(a + 8);
            } else if (MODE == 1) {   // two aligned 8-byte loads
                const uint32_t a = o & ~7u;
                const uint2 x = lds64(a), y = lds64(a + 8);
                acc += x.x + x.y + y.x + y.y;
            } else if (MODE == 2) {   // one aligned 16-byte load
                const uint32_t a = o & ~15u;
                const uint4 x = lds128(a);
                acc += x.x + x.y + x.z + x.w;
            } else if (MODE == 3) {   // two aligned 16-byte loads
                const uint32_t a = o & ~15u;
                const uint4 x = lds128(a), y = lds128(a + 16);
                acc += x.x + x.y + x.z + x.w + y.x + y.y + y.z + y.w;
            } else if (MODE == 4) {   // one word + two indexed shuffles
                const uint32_t a = o & ~3u;
                const uint32_t w = lds32(a);
                const uint32_t s1 = __shfl_sync(0xffffffffu, w, (lane + 1 + (lane & 1)) & 31);
                const uint32_t s2 = __shfl_sync(0xffffffffu, w, (lane + 2 + (lane & 1)) & 31);
                acc += w + s1 + s2;
            } else if (MODE == 5) {   // one aligned 8-byte load
                const uint32_t a = o & ~7u;
                const uint2 x = lds64(a);
                acc += x.x + x.y;
            } else if (MODE == 6) {   // one word
                acc += lds32(o & ~3u);
            } else if (MODE == 7) {   // two shuffles only
                const uint32_t s1 = __shfl_sync(0xffffffffu, b + it * 7 + u, (lane + 1 + (lane & 1)) & 31);
                const uint32_t s2 = __shfl_sync(0xffffffffu, b * 3 + it + u, (lane + 2 + (lane & 1)) & 31);
                acc += s1 + s2;
            }
        }
    }
    const uint32_t acc = accs[0] ^ accs[1] ^ accs[2] ^ accs[3];
    if (acc == 0x12345u) out[blockIdx.x] = acc;
}

template <int MODE>
static void run(const char* what, int stride_b, int phase, int n_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    uint32_t* d;
    cudaMalloc(&d, 4096 * 4);
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    probe<MODE><<<2 * n_sm, 256, 32768>>>(d, stride_b, phase, 1);
    cudaEventRecord(e0);
    probe<MODE><<<2 * n_sm, 256, 32768>>>(d, stride_b, phase, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    // per SM: 16 warps x ITERS x 4 pixel-rows
    const double cyc = ms * 1e-3 * 1.965e9 / (16.0 * ITERS * 4.0);
    printf("%-44s stride %4d/100 B phase %d : %.3f ms  %.2f SM cycles per warp pixel-row  (%s)\n", what, stride_b, phase, ms, cyc,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    const int strides[] = {300, 315, 600, 630, 1200, 1260};
    for (int stride : strides) {
        for (int phase = 0; phase < 4; phase += 3) {
            run<0>("3 x LDS.32", stride, phase, n_sm);
            run<1>("2 x LDS.64", stride, phase, n_sm);
            run<2>("1 x LDS.128", stride, phase, n_sm);
            run<3>("2 x LDS.128", stride, phase, n_sm);
            run<4>("1 x LDS.32 + 2 x SHFL.IDX", stride, phase, n_sm);
            run<5>("1 x LDS.64", stride, phase, n_sm);
            run<6>("1 x LDS.32", stride, phase, n_sm);
            run<7>("2 x SHFL.IDX", stride, phase, n_sm);
        }
    }
    return 0;
}
