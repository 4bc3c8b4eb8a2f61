// Probe (experiments, not part of the library): strided window upload host -> device while a contiguous download
// runs the other way - copy engine (cudaMemcpy2DAsync) against SMs reading mapped pinned memory directly.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o zero_copy_probe zero_copy_probe.cu
// Geometry of the sequence pipeline on 8 x 1080p: frames of 1080 rows x 5760 bytes, the visible window is bytes
// [2304, 5760) of every row, 8 cameras x 16 frames per round.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// one warp per row segment: 16-byte loads from the mapped host pointer, 16-byte stores to the device frame
__global__ void upload_rows(const uint8_t* __restrict__ host, uint8_t* __restrict__ dev, int rows, int row_bytes, int b0,
                            int nbytes) {
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n16 = nbytes >> 4;
    for (int r = warp; r < rows; r += warps) {
        const uint4* s = reinterpret_cast<const uint4*>(host + (size_t)r * row_bytes + b0);
        uint4* d = reinterpret_cast<uint4*>(dev + (size_t)r * row_bytes + b0);
        for (int i = lane; i < n16; i += 32) d[i] = __ldcs(s + i);
    }
}

int main(int argc, char** argv) {
    const int H = 1080, ROW = 5760, F = 16, CAMS = 8, B0 = 2304, NB = 3456;
    const size_t frames = (size_t)CAMS * F, rows = frames * H, bytes = rows * ROW;
    const size_t out_bytes = (size_t)F * 33480000;
    uint8_t *h_src, *d_src, *d_out, *h_out;
    CK(cudaHostAlloc(&h_src, bytes, cudaHostAllocMapped));
    CK(cudaHostAlloc(&h_out, out_bytes, cudaHostAllocDefault));
    CK(cudaMalloc(&d_src, bytes));
    CK(cudaMalloc(&d_out, out_bytes));
    for (size_t i = 0; i < bytes; i += 4096) h_src[i] = (uint8_t)i;
    uint8_t* h_map;
    CK(cudaHostGetDevicePointer(&h_map, h_src, 0));
    cudaStream_t s_up, s_dn;
    CK(cudaStreamCreate(&s_up));
    CK(cudaStreamCreate(&s_dn));
    cudaEvent_t e[4];
    for (auto& x : e) CK(cudaEventCreate(&x));
    const int reps = 6;
    for (int mode = -2; mode < 4; ++mode) {         // -2 / -1 copy engine with 3-D copies (64-row bands / whole height), 0 one 2-D copy per camera, 1..3 SM kernel
        for (int duplex = 0; duplex < 2; ++duplex) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e[0], s_up));
            CK(cudaEventRecord(e[2], s_dn));
            for (int rep = 0; rep < reps; ++rep) {
                if (mode < 0) {
                    const int band = mode == -2 ? 64 : H;
                    for (size_t f = 0; f < frames; f += F)
                        for (int y0 = 0; y0 < H; y0 += band) {
                            cudaMemcpy3DParms p = {};
                            p.srcPtr = make_cudaPitchedPtr(h_src + f * H * ROW, ROW, ROW, H);
                            p.dstPtr = make_cudaPitchedPtr(d_src + f * H * ROW, ROW, ROW, H);
                            p.srcPos = make_cudaPos(B0, y0, 0);
                            p.dstPos = make_cudaPos(B0, y0, 0);
                            p.extent = make_cudaExtent(NB, (size_t)(y0 + band <= H ? band : H - y0), F);
                            p.kind = cudaMemcpyHostToDevice;
                            CK(cudaMemcpy3DAsync(&p, s_up));
                        }
                } else if (mode == 0) {
                    for (size_t f = 0; f < frames; f += F)      // one 2-D copy per camera: F frames are contiguous rows
                        CK(cudaMemcpy2DAsync(d_src + f * H * ROW + B0, ROW, h_src + f * H * ROW + B0, ROW, NB, (size_t)F * H,
                                             cudaMemcpyHostToDevice, s_up));
                } else {
                    const int ctas = 148 << (mode - 1);
                    upload_rows<<<ctas, 256, 0, s_up>>>(h_map, d_src, (int)rows, ROW, B0, NB);
                }
                if (duplex) CK(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, s_dn));
            }
            CK(cudaEventRecord(e[1], s_up));
            CK(cudaEventRecord(e[3], s_dn));
            CK(cudaDeviceSynchronize());
            CK(cudaGetLastError());
            float up = 0, dn = 0;
            CK(cudaEventElapsedTime(&up, e[0], e[1]));
            CK(cudaEventElapsedTime(&dn, e[2], e[3]));
            printf("%-32s duplex %d : H2D windows %.1f GB/s", mode == -2 ? "copy engine, 3-D, 64-row bands" : mode == -1 ? "copy engine, 3-D, whole height" : mode == 0 ? "copy engine, one 2-D copy/cam" : mode == 1 ? "SM kernel, 148 CTAs" : mode == 2 ? "SM kernel, 296 CTAs" : "SM kernel, 592 CTAs",
                   duplex, (double)reps * rows * NB / up / 1e6);
            if (duplex) printf("   D2H %.1f GB/s", (double)reps * out_bytes / dn / 1e6);
            printf("\n");
        }
    }
    return 0;
}
