// Shape fix-up of the compositing path: cv2.resize(img, (W, H), interpolation=INTER_LINEAR) for
// uint8 frames (StitcherClass.py:226-233), bit-exact with OpenCV's 8-bit linear resize.
//
// Recipe (OpenCV resize.cpp, restated in oracle/resize_model.py):
//   scale = 1 / (n_dst / n_src)                                     -- float64, computed on the host
//   f = float32((d + 0.5) * scale - 0.5); s = floor(f); f -= s      -- float32 after the cast
//   columns only: s < 0 -> (0, f = 0);  s >= n_src - 1 -> (n_src - 1, f = 0)
//   w0 = rint(float32(1 - f) * 2048), w1 = rint(f * 2048)
//   rows: taps clip(s, 0, n_src - 1) and clip(s + 1, 0, n_src - 1), weights as computed
//   h_r = p[r][sx] * a0 + p[r][sx + 1] * a1
//   out = (((b0 * (h_0 >> 4)) >> 16) + ((b1 * (h_1 >> 4)) >> 16) + 2) >> 2
//   exact 2 x 2 decimation goes through OpenCV's area kernel: (p00 + p01 + p10 + p11 + 2) >> 2
// The float operations use the round-to-nearest intrinsics, which are never contracted.
#include "mcs_common.h"

struct ResizeArgs {
    const uint8_t* src;
    uint8_t* dst;
    long long src_pitch, src_frame_stride, dst_pitch, dst_frame_stride;
    double scale_x, scale_y;
    int src_w, src_h, dst_w, dst_h;
    int area2;   // exact 2 x 2 decimation
};

__device__ __forceinline__ void linear_coef(int d, double scale, int n_src, bool clamp_edges, int& s, int& w0,
                                            int& w1) {
    float f = __double2float_rn(__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5));
    s = __float2int_rd(f);
    f = __fsub_rn(f, (float)s);
    if (clamp_edges) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }
    }
    w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    w1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

// One thread per 4 consecutive output pixels of one row (block 32 x 8: a warp writes 128 pixels of
// a row), frames along grid z.  Taps come straight from global memory through L1.
template <int C>
__global__ void __launch_bounds__(256)
mcs_resize_linear_kernel(const __grid_constant__ ResizeArgs a) {
    const int x_first = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (y >= a.dst_h || x_first >= a.dst_w) return;
    const uint8_t* src = a.src + (long long)blockIdx.z * a.src_frame_stride;
    const int n_px = min(4, a.dst_w - x_first);
    uint8_t px[4 * C];

    if (a.area2) {
        const uint8_t* r0 = src + (long long)(2 * y) * a.src_pitch;
        const uint8_t* r1 = r0 + a.src_pitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < n_px) {
                const int o = 2 * (x_first + i) * C;
#pragma unroll
                for (int c = 0; c < C; ++c)
                    px[i * C + c] = (uint8_t)((__ldg(r0 + o + c) + __ldg(r0 + o + C + c) + __ldg(r1 + o + c) +
                                               __ldg(r1 + o + C + c) + 2) >> 2);
            }
        }
    } else {
        int sy, b0, b1;
        linear_coef(y, a.scale_y, a.src_h, false, sy, b0, b1);
        const uint8_t* r0 = src + (long long)max(0, min(a.src_h - 1, sy)) * a.src_pitch;
        const uint8_t* r1 = src + (long long)max(0, min(a.src_h - 1, sy + 1)) * a.src_pitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < n_px) {
                int sx, a0, a1;
                linear_coef(x_first + i, a.scale_x, a.src_w, true, sx, a0, a1);
                const int o0 = sx * C, o1 = min(sx + 1, a.src_w - 1) * C;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int h0 = __ldg(r0 + o0 + c) * a0 + __ldg(r0 + o1 + c) * a1;
                    const int h1 = __ldg(r1 + o0 + c) * a0 + __ldg(r1 + o1 + c) * a1;
                    px[i * C + c] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
                }
            }
        }
    }

    uint8_t* out = a.dst + (long long)blockIdx.z * a.dst_frame_stride + (long long)y * a.dst_pitch +
                   (long long)x_first * C;
    if (n_px == 4 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
        uint32_t* o32 = reinterpret_cast<uint32_t*>(out);
#pragma unroll
        for (int wd = 0; wd < C; ++wd)
            o32[wd] = (uint32_t)px[4 * wd] | ((uint32_t)px[4 * wd + 1] << 8) | ((uint32_t)px[4 * wd + 2] << 16) |
                      ((uint32_t)px[4 * wd + 3] << 24);
    } else {
        for (int b = 0; b < n_px * C; ++b) out[b] = px[b];
    }
}

// Separable form for the general (non-area) case.  A CTA of 256 threads produces a 128 x 16 output
// tile in two phases through shared memory:
//   1. horizontal pass: for every source row the tile needs and every output column, h >> 4 (at
//      most 255 * 2048 >> 4 = 32640, a uint16) - thread t owns column t & 127, so its column
//      coefficients are computed once and the byte taps of a warp fall into one or two cache lines;
//   2. vertical pass: a thread takes 4 adjacent pixels of two tile rows, 8-byte reads of the two
//      source rows' h values, OpenCV's (((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2.
// Every source byte is fetched once per tile row it serves instead of once per output pixel.
#define RSEP_TW 128
#define RSEP_TH 16
#define RSEP_SMEM_MAX (40 * 1024)
#define RSEP_TILES_Y 4   // tiles a CTA walks down its column strip

template <int C>
__global__ void __launch_bounds__(256)
mcs_resize_sep_kernel(const __grid_constant__ ResizeArgs a) {
    extern __shared__ __align__(16) uint16_t hbuf[];   // [rows][RSEP_TW * C]
    __shared__ int4 rowc[RSEP_TH];   // per tile row {first source row, second (relative to ry0), b0 << 16, b1 << 16}
    const int x0 = blockIdx.x * RSEP_TW;
    const int tid = threadIdx.x;
    const uint8_t* src = a.src + (long long)blockIdx.z * a.src_frame_stride;
    constexpr int ROW = RSEP_TW * C;   // uint16 per staged row
    // column coefficients: once per CTA, which walks RSEP_TILES_Y tiles down its 128-column strip (the float64
    // coefficient recipe and the address set-up cost ~300 instructions per thread, as much as 8 pixels of work)
    const int col = tid & (RSEP_TW - 1);
    int sx, a0, a1;
    linear_coef(min(x0 + col, a.dst_w - 1), a.scale_x, a.src_w, true, sx, a0, a1);
    const int o0 = sx * C, o1 = min(sx + 1, a.src_w - 1) * C;
    const int lane = tid & 31, warp = tid >> 5;
    const int x_first = x0 + 4 * lane;
    const int n_px = min(4, a.dst_w - x_first);   // <= 0: this thread has no output columns

    for (int ty = 0; ty < RSEP_TILES_Y; ++ty) {
        const int y0 = (blockIdx.y * RSEP_TILES_Y + ty) * RSEP_TH;
        if (y0 >= a.dst_h) break;                  // uniform over the CTA
        const int y_last = min(y0 + RSEP_TH, a.dst_h) - 1;
        int ry0, ry1;
        {
            int s0, s1, w0, w1;
            linear_coef(y0, a.scale_y, a.src_h, false, s0, w0, w1);
            linear_coef(y_last, a.scale_y, a.src_h, false, s1, w0, w1);
            ry0 = max(0, min(a.src_h - 1, s0));
            ry1 = max(0, min(a.src_h - 1, s1 + 1));
        }
        if (ty) __syncthreads();                   // the previous tile's vertical pass is done with hbuf / rowc
        // phase 1: horizontal pass, h >> 4 of every source row the tile needs
        for (int r = ry0 + (tid >> 7); r <= ry1; r += 2) {
            const uint8_t* row = src + (long long)r * a.src_pitch;
            uint16_t* h = hbuf + (r - ry0) * ROW + col * C;
#pragma unroll
            for (int c = 0; c < C; ++c)
                h[c] = (uint16_t)((__ldg(row + o0 + c) * a0 + __ldg(row + o1 + c) * a1) >> 4);
        }
        if (tid < RSEP_TH) {
            int sy, b0, b1;
            linear_coef(min(y0 + tid, a.dst_h - 1), a.scale_y, a.src_h, false, sy, b0, b1);
            rowc[tid] = make_int4(max(0, min(a.src_h - 1, sy)) - ry0, max(0, min(a.src_h - 1, sy + 1)) - ry0, b0 << 16, b1 << 16);
        }
        __syncthreads();
        // phase 2: (((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2 with (b * h) >> 16 = umulhi(b << 16, h): two
        // IMAD.HI.U32 chained through their addend and one shift per value (b <= 2048, h <= 32640: no overflow)
        if (n_px <= 0) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int y = y0 + warp + 8 * half;
            if (y >= a.dst_h) break;
            const int4 rc = rowc[warp + 8 * half];
            const uint32_t b0 = (uint32_t)rc.z, b1 = (uint32_t)rc.w;
            // 4 pixels x C values = 4 * C uint16 = C 8-byte words per source row
            const uint2* h0 = reinterpret_cast<const uint2*>(hbuf + rc.x * ROW + 4 * lane * C);
            const uint2* h1 = reinterpret_cast<const uint2*>(hbuf + rc.y * ROW + 4 * lane * C);
            uint32_t o32[C];   // the 4 * C output bytes in memory order
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const uint2 u = h0[k], v = h1[k];
                const uint32_t v0 = (__umulhi(b0, u.x & 0xffffu) + __umulhi(b1, v.x & 0xffffu) + 2u) >> 2;
                const uint32_t v1 = (__umulhi(b0, u.x >> 16) + __umulhi(b1, v.x >> 16) + 2u) >> 2;
                const uint32_t v2 = (__umulhi(b0, u.y & 0xffffu) + __umulhi(b1, v.y & 0xffffu) + 2u) >> 2;
                const uint32_t v3 = (__umulhi(b0, u.y >> 16) + __umulhi(b1, v.y >> 16) + 2u) >> 2;
                o32[k] = __byte_perm(__byte_perm(v0, v1, 0x0040), __byte_perm(v2, v3, 0x0040), 0x5410);
            }
            uint8_t* out = a.dst + (long long)blockIdx.z * a.dst_frame_stride + (long long)y * a.dst_pitch +
                           (long long)x_first * C;
            if (n_px == 4 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
#pragma unroll
                for (int wd = 0; wd < C; ++wd) reinterpret_cast<uint32_t*>(out)[wd] = o32[wd];
            } else {
                for (int b = 0; b < n_px * C; ++b) out[b] = (uint8_t)(o32[b >> 2] >> (8 * (b & 3)));
            }
        }
    }
}

extern "C" int mcs_resize_linear_u8(const uint8_t* src, int src_w, int src_h, int64_t src_pitch_bytes,
                                    int64_t src_frame_stride, uint8_t* dst, int dst_w, int dst_h,
                                    int64_t dst_pitch_bytes, int64_t dst_frame_stride, int channels,
                                    int n_frames, void* cuda_stream) {
    MCS_CHECK_ARG(channels == 1 || channels == 3 || channels == 4,
                  "mcs_resize_linear_u8: channels=%d (supported: 1, 3, 4)", channels);
    MCS_CHECK_ARG(n_frames >= 0 && n_frames <= 65535, "mcs_resize_linear_u8: n_frames=%d outside 0..65535", n_frames);
    MCS_CHECK_ARG(src_w > 0 && src_h > 0 && src_w < (1 << 24) && src_h < (1 << 24),
                  "mcs_resize_linear_u8: source size %dx%d out of range", src_w, src_h);
    MCS_CHECK_ARG(dst_w >= 0 && dst_h >= 0 && dst_w < (1 << 24) && dst_h < (1 << 24),
                  "mcs_resize_linear_u8: destination size %dx%d out of range", dst_w, dst_h);
    if (n_frames == 0 || dst_w == 0 || dst_h == 0) return MCS_OK;
    MCS_CHECK_ARG(src != nullptr && dst != nullptr, "mcs_resize_linear_u8: NULL image pointer");
    MCS_CHECK_ARG(src_pitch_bytes >= (int64_t)src_w * channels && dst_pitch_bytes >= (int64_t)dst_w * channels,
                  "mcs_resize_linear_u8: pitch smaller than a row");
    ResizeArgs a;
    a.src = src;
    a.dst = dst;
    a.src_pitch = src_pitch_bytes;
    a.src_frame_stride = n_frames > 1 ? src_frame_stride : 0;
    a.dst_pitch = dst_pitch_bytes;
    a.dst_frame_stride = n_frames > 1 ? dst_frame_stride : 0;
    // cv::resize: inv_scale = (double)dsize / ssize, then scale = 1. / inv_scale
    volatile double inv_x = (double)dst_w / (double)src_w, inv_y = (double)dst_h / (double)src_h;
    a.scale_x = 1.0 / inv_x;
    a.scale_y = 1.0 / inv_y;
    a.src_w = src_w; a.src_h = src_h; a.dst_w = dst_w; a.dst_h = dst_h;
    a.area2 = (a.scale_x == 2.0 && a.scale_y == 2.0) ? 1 : 0;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    // separable two-phase kernel unless the tile's source rows exceed the shared-memory budget
    // (decimation beyond ~3 x) or the area kernel applies
    const long long sep_rows = (long long)(RSEP_TH * a.scale_y) + 4;
    const long long sep_smem = sep_rows * RSEP_TW * channels * 2;
    if (!a.area2 && sep_smem <= RSEP_SMEM_MAX) {
        const dim3 sgrid((dst_w + RSEP_TW - 1) / RSEP_TW, (dst_h + RSEP_TH * RSEP_TILES_Y - 1) / (RSEP_TH * RSEP_TILES_Y), n_frames);
        switch (channels) {
            case 1: mcs_resize_sep_kernel<1><<<sgrid, 256, (size_t)sep_smem, stream>>>(a); break;
            case 3: mcs_resize_sep_kernel<3><<<sgrid, 256, (size_t)sep_smem, stream>>>(a); break;
            default: mcs_resize_sep_kernel<4><<<sgrid, 256, (size_t)sep_smem, stream>>>(a); break;
        }
        mcs_count_launch(1);
        MCS_CHECK_CUDA(cudaGetLastError());
        return MCS_OK;
    }
    const dim3 block(32, 8, 1);
    const dim3 grid((dst_w + 127) / 128, (dst_h + 7) / 8, n_frames);
    switch (channels) {
        case 1: mcs_resize_linear_kernel<1><<<grid, block, 0, stream>>>(a); break;
        case 3: mcs_resize_linear_kernel<3><<<grid, block, 0, stream>>>(a); break;
        default: mcs_resize_linear_kernel<4><<<grid, block, 0, stream>>>(a); break;
    }
    mcs_count_launch(1);
    MCS_CHECK_CUDA(cudaGetLastError());
    return MCS_OK;
}
