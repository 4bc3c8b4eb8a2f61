// Fused warp + paste compositing kernels (sm_100a).
//
// One launch produces whole panoramas: every output pixel finds the innermost layer whose
// rectangle contains it (the reference's repeated "warp A, then paste B over it",
// StitcherClass.py:239-241 applied N-1 times by :131-136) and is either copied (camera 0)
// or resampled with OpenCV's fixed-point bilinear recipe (cv2.warpPerspective, :239).
//
// Coordinate recipe (bit-exact with OpenCV's WarpPerspectiveInvoker):
//   (xl, yl) = pixel in the layer's own canvas frame;  xb = xl & ~63, x1 = xl & 63
//   X0 = (Mi0*xb + Mi1*yl) + Mi2  (likewise Y0, W0)          -- float64, round-to-nearest, NO fma
//   W  = W0 + Mi6*x1 ; W = W ? 32/W : 0
//   X  = rint(clamp((X0 + Mi0*x1) * W))   Y likewise          -- 1/32-px fixed point
//   out = (sum_taps w*p + 16384) >> 15, w = 32 * {(32-ay)(32-ax), (32-ay)ax, ay(32-ax), ay*ax}
// The float64 operations use the __d*_rn intrinsics, which the compiler never contracts.
#include "mcs_device.cuh"

#include <string.h>
#include <stdlib.h>

struct LayerArgs {
    McsLayer g;
    const uint8_t* src;
    long long pitch;
    long long frame_stride;
};

struct StitchArgs {
    LayerArgs L[MCS_MAX_LAYERS];
    uint8_t* dst;
    long long dst_pitch;
    long long dst_frame_stride;
    int n_layers;
    int out_w, out_h;
    int n_frames;
};

// --------------------------------------------------------------------------------------------
// per-pixel building blocks

template <int C>
__device__ __forceinline__ void load_px(const uint8_t* __restrict__ p, int (&v)[C]) {
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = __ldg(p + c);
}

// Fixed-point bilinear sample of `src` at (X, Y) in 1/32 px.  `touched` reports whether any tap
// was inside the source (used by the ownership statistics).
template <int C>
__device__ __forceinline__ bool sample_u8(const uint8_t* __restrict__ src, long long pitch, int src_w,
                                          int src_h, int X, int Y, int (&out)[C]) {
    int sx = X >> 5, sy = Y >> 5;
    // saturate_cast<short> of the integer part: anything clamped is outside any legal source anyway
    sx = max(-32768, min(32767, sx));
    sy = max(-32768, min(32767, sy));
    const int ax = X & 31, ay = Y & 31;
    const bool x0in = (unsigned)sx < (unsigned)src_w, x1in = (unsigned)(sx + 1) < (unsigned)src_w;
    const bool y0in = (unsigned)sy < (unsigned)src_h, y1in = (unsigned)(sy + 1) < (unsigned)src_h;
    int p00[C], p01[C], p10[C], p11[C];
#pragma unroll
    for (int c = 0; c < C; ++c) p00[c] = p01[c] = p10[c] = p11[c] = 0;
    const uint8_t* r0 = src + (long long)sy * pitch + (long long)sx * C;
    const uint8_t* r1 = r0 + pitch;
    if (y0in && x0in) load_px<C>(r0, p00);
    if (y0in && x1in) load_px<C>(r0 + C, p01);
    if (y1in && x0in) load_px<C>(r1, p10);
    if (y1in && x1in) load_px<C>(r1 + C, p11);
    const int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32;
    const int w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
#pragma unroll
    for (int c = 0; c < C; ++c)
        out[c] = (w00 * p00[c] + w01 * p01[c] + w10 * p10[c] + w11 * p11[c] + 16384) >> 15;
    return (x0in || x1in) && (y0in || y1in);
}

__device__ __forceinline__ int find_owner(const StitchArgs& a, int x, int y) {
    for (int k = 0; k < a.n_layers; ++k) {
        const McsLayer& g = a.L[k].g;
        if (x >= g.rx0 && x < g.rx1 && y >= g.ry0 && y < g.ry1) return k;
    }
    return -1;
}

// --------------------------------------------------------------------------------------------
// Variant 1 ("gather"): one thread per 4 consecutive output pixels, source taps fetched straight
// from global memory.  Handles every layout (any pitch / alignment / homography); it is the
// fallback of the tiled variant, never of a CPU path.
//
// STATS = true turns the kernel into the ownership counter behind mcs_plan_owned_pixels.
template <int C, bool STATS>
__global__ void __launch_bounds__(256)
mcs_stitch_gather_kernel(const __grid_constant__ StitchArgs a, unsigned long long* __restrict__ owned) {
    const int x_first = blockIdx.x * MCS_TILE_W + threadIdx.x * 4;
    const int y = blockIdx.y * MCS_TILE_H + threadIdx.y;
    if (y >= a.out_h || x_first >= a.out_w) return;
    const int frame = blockIdx.z;

    int prev_owner = -2, prev_xb = 0;
    RowBlock rb = {0.0, 0.0, 0.0};
    uint8_t px[4 * C];
    const int n_px = min(4, a.out_w - x_first);

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = 0;
        const int x = x_first + i;
        if (i < n_px) {
            const int k = find_owner(a, x, y);
            if (k >= 0) {
                const LayerArgs& L = a.L[k];
                const int xl = x - L.g.ox, yl = y - L.g.oy;
                const uint8_t* src = L.src + (long long)frame * L.frame_stride;
                bool touched = true;
                if (L.g.kind == MCS_LAYER_COPY) {
                    if (!STATS) load_px<C>(src + (long long)yl * L.pitch + (long long)xl * C, v);
                } else {
                    const int xb = xl & ~63;
                    if (k != prev_owner || xb != prev_xb) {
                        rb = row_block(L.g.mi, xb, yl);
                        prev_owner = k;
                        prev_xb = xb;
                    }
                    int X, Y;
                    fixed_coords(L.g.mi[0], L.g.mi[3], L.g.mi[6], rb, xl & 63, X, Y);
                    if (STATS) {
                        const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
                        touched = ((unsigned)sx < (unsigned)L.g.src_w || (unsigned)(sx + 1) < (unsigned)L.g.src_w) &&
                                  ((unsigned)sy < (unsigned)L.g.src_h || (unsigned)(sy + 1) < (unsigned)L.g.src_h);
                    } else {
                        sample_u8<C>(src, L.pitch, L.g.src_w, L.g.src_h, X, Y, v);
                    }
                }
                if (STATS && touched) atomicAdd(owned + k, 1ULL);
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) px[i * C + c] = (uint8_t)v[c];
    }
    if (STATS) return;

    uint8_t* out = a.dst + (long long)frame * a.dst_frame_stride + (long long)y * a.dst_pitch +
                   (long long)x_first * C;
    const bool aligned4 = ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
    if (n_px == 4 && aligned4) {
        uint32_t* o32 = reinterpret_cast<uint32_t*>(out);
#pragma unroll
        for (int wd = 0; wd < C; ++wd)
            o32[wd] = (uint32_t)px[4 * wd] | ((uint32_t)px[4 * wd + 1] << 8) |
                      ((uint32_t)px[4 * wd + 2] << 16) | ((uint32_t)px[4 * wd + 3] << 24);
    } else {
        for (int b = 0; b < n_px * C; ++b) out[b] = px[b];
    }
}

// --------------------------------------------------------------------------------------------
// Variant 2 ("tiled"): persistent CTAs walk the plan's tile table.  For every tile one elected
// thread has the TMA engine stage the bounding box of the source pixels the tile touches
// (cp.async.bulk.tensor, zero fill outside the image = BORDER_CONSTANT 0) into one of two
// shared-memory buffers while the CTA works on the previous tile; the 256 threads then resample
// 4 consecutive pixels each out of shared memory, assemble the 128 x 16 output cell in shared
// memory and stream it to the panorama with 16-byte stores realigned to the destination.
// Every source byte is fetched from L2/HBM once per tile that touches it and every output byte
// is written exactly once.

#include <cuda.h>   // CUtensorMap

#define TILED_THREADS 256
#define TILED_WARPS (TILED_THREADS / 32)

struct TiledArgs {
    CUtensorMap tmap[MCS_MAX_LAYERS];   // source of each layer as (row words, rows, frames) of uint32
    const McsTile* tiles;
    const McsLayer* layers;
    uint8_t* dst;
    long long dst_pitch;
    long long dst_frame_stride;
    int n_tiles;
    int n_frames;
    int box_bytes;                      // bytes of one staging buffer
    int debug;                          // MCS_DEBUG_TILED: 1 = no TMA (all zero), 2 = TMA + wait only
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// Stream `h` rows of `nbytes` bytes from shared memory (row r at s_base + r*s_pitch + s_off, any
// alignment) - or zeros - to global rows (row r at g + r*g_pitch, any alignment).  The body of
// each row goes out as 16-byte stores aligned to the destination; the source words are realigned
// with funnel shifts.  Warp w handles rows w, w + 8.
__device__ __forceinline__ void write_rows(const uint8_t* s_base, int s_pitch, int s_off, uint8_t* g,
                                           long long g_pitch, int nbytes, int h, bool zeros) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < h; r += TILED_WARPS) {
        uint8_t* gr = g + (long long)r * g_pitch;
        const uint8_t* sr = s_base + r * s_pitch + s_off;
        const int head = min(nbytes, (int)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(gr) & 15)) & 15));
        const int nchunks = (nbytes - head) >> 4;
        const int tail0 = head + (nchunks << 4);
        if (lane < head) gr[lane] = zeros ? (uint8_t)0 : sr[lane];
        if (lane >= 16 && tail0 + (lane - 16) < nbytes) gr[tail0 + lane - 16] = zeros ? (uint8_t)0 : sr[tail0 + lane - 16];
        for (int c = lane; c < nchunks; c += 32) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (!zeros) {
                const uint32_t sa = smem_u32(sr + head + (c << 4));
                const uint32_t* w = reinterpret_cast<const uint32_t*>(sr + head + (c << 4) - (sa & 3));
                const uint32_t sh = (sa & 3) * 8;
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
                v.x = __funnelshift_r(w0, w1, sh);
                v.y = __funnelshift_r(w1, w2, sh);
                v.z = __funnelshift_r(w2, w3, sh);
                v.w = __funnelshift_r(w3, w4, sh);
            }
            *reinterpret_cast<uint4*>(gr + head + (c << 4)) = v;
        }
    }
}

// Resample one pixel at 1/32-px (X, Y) from the staged box.  `box` points at box word (0,0);
// `sp` is the box pitch in bytes; (bx, by) the box origin (bx in 4-byte words).
template <int C>
__device__ __forceinline__ void sample_box(const uint8_t* box, int sp, int bx, int by, int src_w, int src_h,
                                           int X, int Y, uint32_t (&out)[C]) {
    const int sx = max(-2, min(src_w, X >> 5)), sy = max(-2, min(src_h, Y >> 5));
    const uint32_t ax = X & 31, ay = Y & 31;
    const int off = sx * C - 4 * bx;                       // byte offset of tap (sx, .) in a box row
    const uint8_t* p = box + (sy - by) * sp + (off & ~3);
    const uint32_t sh = (off & 3) * 8;
    uint32_t lo0, hi0 = 0, lo1, hi1 = 0;                   // bytes [0,4) and [4,8) of the 2-tap run, rows 0/1
    {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
        const uint32_t* v = reinterpret_cast<const uint32_t*>(p + sp);
        if (C == 4) {
            lo0 = w[0]; hi0 = w[1]; lo1 = v[0]; hi1 = v[1];
        } else if (C == 3) {
            const uint32_t a0 = w[0], a1 = w[1], a2 = w[2], b0 = v[0], b1 = v[1], b2 = v[2];
            lo0 = __funnelshift_r(a0, a1, sh); hi0 = __funnelshift_r(a1, a2, sh);
            lo1 = __funnelshift_r(b0, b1, sh); hi1 = __funnelshift_r(b1, b2, sh);
        } else {
            const uint32_t a0 = w[0], a1 = w[1], b0 = v[0], b1 = v[1];
            lo0 = __funnelshift_r(a0, a1, sh);
            lo1 = __funnelshift_r(b0, b1, sh);
        }
    }
    const uint32_t w00 = (32 - ay) * (32 - ax), w01 = (32 - ay) * ax, w10 = ay * (32 - ax), w11 = ay * ax;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        // tap 0 = bytes [0, C), tap 1 = bytes [C, 2C) of the 8-byte run (lo, hi)
        const int i0 = c, i1 = C + c;
        const uint32_t p00 = ((i0 < 4 ? lo0 : hi0) >> (8 * (i0 & 3))) & 0xff;
        const uint32_t p01 = ((i1 < 4 ? lo0 : hi0) >> (8 * (i1 & 3))) & 0xff;
        const uint32_t p10 = ((i0 < 4 ? lo1 : hi1) >> (8 * (i0 & 3))) & 0xff;
        const uint32_t p11 = ((i1 < 4 ? lo1 : hi1) >> (8 * (i1 & 3))) & 0xff;
        // (sum * 32 + 16384) >> 15 == (sum + 512) >> 10
        out[c] = (w00 * p00 + w01 * p01 + w10 * p10 + w11 * p11 + 512) >> 10;
    }
}

template <int C>
__global__ void __launch_bounds__(TILED_THREADS)
mcs_stitch_tiled_kernel(const __grid_constant__ TiledArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int OUT_PITCH = MCS_CELL_W * C + 16;
    uint8_t* buf[2] = {smem, smem + a.box_bytes};
    uint8_t* s_out = smem + 2 * a.box_bytes;                                   // [16][OUT_PITCH] + 16 spare
    RowBlock* s_rows = reinterpret_cast<RowBlock*>(s_out + MCS_CELL_H * OUT_PITCH + 16);   // [16][2]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_rows + MCS_CELL_H * 2);    // [2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n_items = (long long)a.n_tiles * a.n_frames;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // issue the TMA load of work item `item` into staging buffer `slot` (thread 0 only)
    auto issue = [&](long long item, int slot) {
        const int t = (int)(item % a.n_tiles), frame = (int)(item / a.n_tiles);
        const McsTile tile = a.tiles[t];
        const McsLayer& L = a.layers[tile.layer];
        mbar_expect_tx(&s_bar[slot], (uint32_t)(L.bw4 * 4 * L.bh));
        tma_load_3d(buf[slot], &a.tmap[tile.layer], tile.bx, tile.by, frame, &s_bar[slot]);
    };
    auto has_load = [&](long long item) -> bool {
        return a.debug != 1 && a.tiles[(int)(item % a.n_tiles)].cls != MCS_TILE_ZERO;
    };

    int loads_issued = 0;     // thread 0: loads issued so far
    int loads_used = 0;       // all threads: loads consumed so far
    long long item = blockIdx.x;
    if (tid == 0 && item < n_items && has_load(item)) {
        issue(item, 0);
        loads_issued = 1;
    }

    for (; item < n_items; item += gridDim.x) {
        const int t = (int)(item % a.n_tiles), frame = (int)(item / a.n_tiles);
        const McsTile tile = a.tiles[t];
        const bool loaded = a.debug != 1 && tile.cls != MCS_TILE_ZERO;
        const int slot = loads_used & 1;

        if (tile.cls == MCS_TILE_WARP && tid < 2 * MCS_CELL_H) {
            const McsLayer& L = a.layers[tile.layer];
            const int r = tid >> 1, b = tid & 1;
            s_rows[tid] = row_block(L.mi, tile.cx0 - L.ox + 64 * b, tile.y0 + r - L.oy);
        }
        __syncthreads();   // (A) row table ready; the previous tile's write-out has finished

        // prefetch the next tile's source box: its buffer was last read by the tile before this one
        if (tid == 0) {
            const long long nxt = item + gridDim.x;
            if (nxt < n_items && has_load(nxt)) {
                issue(nxt, loads_issued & 1);
                loads_issued += 1;
            }
        }

        uint8_t* g = a.dst + (long long)frame * a.dst_frame_stride + (long long)tile.y0 * a.dst_pitch +
                     (long long)(tile.cx0 + tile.c0) * C;
        const int nbytes = (tile.c1 - tile.c0) * C;

        if (!loaded) {
            write_rows(nullptr, 0, 0, g, a.dst_pitch, nbytes, tile.h, true);
            continue;
        }
        mbar_wait(&s_bar[slot], (uint32_t)((loads_used >> 1) & 1));
        loads_used += 1;
        if (a.debug == 2) { write_rows(nullptr, 0, 0, g, a.dst_pitch, nbytes, tile.h, true); continue; }
        const McsLayer& L = a.layers[tile.layer];
        const int sp = L.bw4 * 4;

        if (tile.cls == MCS_TILE_COPY) {
            const int s_off = (tile.cx0 + tile.c0 - L.ox) * C - 4 * tile.bx;
            write_rows(buf[slot], sp, s_off, g, a.dst_pitch, nbytes, tile.h, false);
            continue;
        }

        // ---- WARP: lane -> cell columns 4*lane .. 4*lane+3, warp -> rows warp, warp+8 ----------
        {
            const double m0 = L.mi[0], m3 = L.mi[3], m6 = L.mi[6];
            const int src_w = L.src_w, src_h = L.src_h;
            const int col0 = 4 * lane;
            const int blk = lane >> 4;
            const uint8_t* box = buf[slot];
            if (col0 + 4 > tile.c0 && col0 < tile.c1) {
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = warp + rr * TILED_WARPS;
                    if (r >= tile.h) break;
                    const RowBlock rb = s_rows[2 * r + blk];
                    uint32_t px[4][C];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int col = col0 + j;
#pragma unroll
                        for (int c = 0; c < C; ++c) px[j][c] = 0;
                        if (col >= tile.c0 && col < tile.c1) {
                            int X, Y;
                            fixed_coords(m0, m3, m6, rb, col & 63, X, Y);
                            sample_box<C>(box, sp, tile.bx, tile.by, src_w, src_h, X, Y, px[j]);
                        }
                    }
                    uint32_t* o = reinterpret_cast<uint32_t*>(s_out + r * OUT_PITCH + col0 * C);
                    if (C == 3) {
                        o[0] = px[0][0] | (px[0][1] << 8) | (px[0][2] << 16) | (px[1][0] << 24);
                        o[1] = px[1][1] | (px[1][2] << 8) | (px[2][0] << 16) | (px[2][1] << 24);
                        o[2] = px[2][2] | (px[3][0] << 8) | (px[3][1] << 16) | (px[3][2] << 24);
                    } else if (C == 4) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            o[j] = px[j][0] | (px[j][1 % C] << 8) | (px[j][2 % C] << 16) | (px[j][3 % C] << 24);
                    } else {
                        o[0] = px[0][0] | (px[1][0] << 8) | (px[2][0] << 16) | (px[3][0] << 24);
                    }
                }
            }
        }
        __syncthreads();   // (B) output cell complete
        write_rows(s_out, OUT_PITCH, tile.c0 * C, g, a.dst_pitch, nbytes, tile.h, false);
    }
}

// --------------------------------------------------------------------------------------------
static int fill_args(const mcs_plan* plan, StitchArgs& a, const uint8_t* const* src,
                     const int64_t* src_pitch, const int64_t* src_frame_stride, int n_frames,
                     uint8_t* dst, int64_t dst_pitch, int64_t dst_frame_stride) {
    a.n_layers = plan->n_layers;
    a.out_w = plan->out_w;
    a.out_h = plan->out_h;
    a.n_frames = n_frames;
    a.dst = dst;
    a.dst_pitch = dst_pitch;
    a.dst_frame_stride = dst_frame_stride;
    for (int k = 0; k < plan->n_layers; ++k) {
        a.L[k].g = plan->layers[k];
        a.L[k].src = src ? src[k] : nullptr;
        a.L[k].pitch = src_pitch ? src_pitch[k] : 0;
        a.L[k].frame_stride = (src_frame_stride && n_frames > 1) ? src_frame_stride[k] : 0;
    }
    return MCS_OK;
}

template <bool STATS>
static cudaError_t launch_gather(const mcs_plan* plan, const StitchArgs& a, unsigned long long* owned,
                                 cudaStream_t stream) {
    dim3 block(MCS_TILE_W / 4, MCS_TILE_H, 1);
    dim3 grid((plan->out_w + MCS_TILE_W - 1) / MCS_TILE_W, (plan->out_h + MCS_TILE_H - 1) / MCS_TILE_H,
              a.n_frames);
    switch (plan->channels) {
        case 1: mcs_stitch_gather_kernel<1, STATS><<<grid, block, 0, stream>>>(a, owned); break;
        case 3: mcs_stitch_gather_kernel<3, STATS><<<grid, block, 0, stream>>>(a, owned); break;
        default: mcs_stitch_gather_kernel<4, STATS><<<grid, block, 0, stream>>>(a, owned); break;
    }
    mcs_count_launch(1);
    return cudaGetLastError();
}

// ---- tiled variant, host side --------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static size_t tiled_smem_bytes(const mcs_plan* plan) {
    const int out_pitch = MCS_CELL_W * plan->channels + 16;
    return 2 * (size_t)plan->box_bytes + (size_t)MCS_CELL_H * out_pitch + 16 +
           sizeof(RowBlock) * MCS_CELL_H * 2 + 2 * sizeof(uint64_t);
}

// Why the tiled variant cannot serve this call (nullptr = it can).
static const char* tiled_blocker(const mcs_plan* plan, const uint8_t* const* src, const int64_t* pitch,
                                 const int64_t* fstride, int n_frames) {
    if (!plan->tiled_ok) return plan->tiled_why;
    if (!get_encode_fn()) return "cuTensorMapEncodeTiled unavailable";
    for (int k = 0; k < plan->n_layers; ++k) {
        if ((reinterpret_cast<uintptr_t>(src[k]) & 15) != 0) return "source base not 16-byte aligned";
        if ((pitch[k] & 15) != 0) return "source pitch not a multiple of 16 bytes";
        if (n_frames > 1 && (fstride[k] & 15) != 0) return "source frame stride not a multiple of 16 bytes";
        if (n_frames > 1 && fstride[k] <= 0) return "non-positive source frame stride";
    }
    return nullptr;
}

static int launch_tiled(mcs_plan* plan, const uint8_t* const* src, const int64_t* pitch,
                        const int64_t* fstride, int n_frames, uint8_t* dst, int64_t dst_pitch,
                        int64_t dst_frame_stride, cudaStream_t stream) {
    TiledArgs a;
    memset(&a, 0, sizeof(a));
    bool hit = plan->cache_valid && plan->cache_frames == n_frames;
    for (int k = 0; hit && k < plan->n_layers; ++k)
        hit = plan->cache_src[k] == src[k] && plan->cache_pitch[k] == pitch[k] &&
              (n_frames == 1 || plan->cache_fstride[k] == fstride[k]);
    CUtensorMap* cache = reinterpret_cast<CUtensorMap*>(
        (reinterpret_cast<uintptr_t>(plan->tmap_cache) + 63) & ~(uintptr_t)63);
    if (!hit) {
        EncodeTiledFn enc = get_encode_fn();
        for (int k = 0; k < plan->n_layers; ++k) {
            const McsLayer& L = plan->layers[k];
            const cuuint64_t dims[3] = {(cuuint64_t)(L.src_w * plan->channels / 4), (cuuint64_t)L.src_h,
                                        (cuuint64_t)n_frames};
            const cuuint64_t strides[2] = {(cuuint64_t)pitch[k],
                                           (cuuint64_t)(n_frames > 1 ? fstride[k] : pitch[k] * L.src_h)};
            const cuuint32_t box[3] = {(cuuint32_t)L.bw4, (cuuint32_t)L.bh, 1u};
            const cuuint32_t estr[3] = {1u, 1u, 1u};
            CUresult r = enc(&cache[k], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(src[k]), dims,
                             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                mcs_set_error("mcs_stitch_u8: cuTensorMapEncodeTiled failed for layer %d (CUresult %d)", k, (int)r);
                plan->cache_valid = 0;
                return MCS_ERR_CUDA;
            }
            plan->cache_src[k] = src[k];
            plan->cache_pitch[k] = pitch[k];
            plan->cache_fstride[k] = n_frames > 1 ? fstride[k] : 0;
        }
        plan->cache_frames = n_frames;
        plan->cache_valid = 1;
    }
    for (int k = 0; k < plan->n_layers; ++k) a.tmap[k] = cache[k];
    a.tiles = plan->d_tiles;
    a.layers = plan->d_layers;
    a.dst = dst;
    a.dst_pitch = dst_pitch;
    a.dst_frame_stride = dst_frame_stride;
    a.n_tiles = plan->n_tiles;
    a.n_frames = n_frames;
    a.box_bytes = plan->box_bytes;
    { const char* d = getenv("MCS_DEBUG_TILED"); a.debug = d ? atoi(d) : 0; }

    const size_t smem = tiled_smem_bytes(plan);
    void (*kern)(TiledArgs) = plan->channels == 1   ? mcs_stitch_tiled_kernel<1>
                              : plan->channels == 3 ? mcs_stitch_tiled_kernel<3>
                                                    : mcs_stitch_tiled_kernel<4>;
    MCS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, n_sm = 0, dev = 0;
    MCS_CHECK_CUDA(cudaGetDevice(&dev));
    MCS_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    MCS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TILED_THREADS, smem));
    if (per_sm < 1) {
        mcs_set_error("mcs_stitch_u8: tiled kernel does not fit (smem %zu bytes)", smem);
        return MCS_ERR_UNSUPPORTED;
    }
    const long long n_items = (long long)plan->n_tiles * n_frames;
    long long grid = (long long)n_sm * per_sm;
    if (grid > n_items) grid = n_items;
    kern<<<(unsigned)grid, TILED_THREADS, smem, stream>>>(a);
    mcs_count_launch(1);
    MCS_CHECK_CUDA(cudaGetLastError());
    return MCS_OK;
}

extern "C" int mcs_stitch_u8(const mcs_plan* plan_c, const uint8_t* const* src,
                             const int64_t* src_pitch_bytes, const int64_t* src_frame_stride,
                             int n_frames, uint8_t* dst, int64_t dst_pitch_bytes,
                             int64_t dst_frame_stride, void* cuda_stream) {
    mcs_plan* plan = const_cast<mcs_plan*>(plan_c);
    MCS_CHECK_ARG(plan != nullptr, "mcs_stitch_u8: plan is NULL");
    MCS_CHECK_ARG(src && src_pitch_bytes, "mcs_stitch_u8: NULL source table");
    MCS_CHECK_ARG(n_frames >= 0 && n_frames <= 65535, "mcs_stitch_u8: n_frames=%d outside 0..65535", n_frames);
    if (n_frames == 0 || plan->out_w == 0 || plan->out_h == 0) return MCS_OK;
    MCS_CHECK_ARG(dst != nullptr, "mcs_stitch_u8: dst is NULL");
    MCS_CHECK_ARG(dst_pitch_bytes >= (int64_t)plan->out_w * plan->channels,
                  "mcs_stitch_u8: dst pitch %lld < row bytes %lld", (long long)dst_pitch_bytes,
                  (long long)plan->out_w * plan->channels);
    MCS_CHECK_ARG(n_frames == 1 || src_frame_stride != nullptr, "mcs_stitch_u8: NULL frame strides");
    for (int k = 0; k < plan->n_layers; ++k) {
        MCS_CHECK_ARG(src[k] != nullptr, "mcs_stitch_u8: source %d is NULL", k);
        MCS_CHECK_ARG(src_pitch_bytes[k] >= (int64_t)plan->layers[k].src_w * plan->channels,
                      "mcs_stitch_u8: source %d pitch %lld < row bytes", k, (long long)src_pitch_bytes[k]);
    }
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const int force = plan->force_variant;
    const char* blocker = tiled_blocker(plan, src, src_pitch_bytes, src_frame_stride, n_frames);
    if (force == 2 && blocker) {
        mcs_set_error("mcs_stitch_u8: tiled variant forced but unavailable: %s", blocker);
        return MCS_ERR_UNSUPPORTED;
    }
    if (!blocker && force != 1) {
        int rc = launch_tiled(plan, src, src_pitch_bytes, src_frame_stride, n_frames, dst, dst_pitch_bytes,
                              dst_frame_stride, stream);
        if (rc != MCS_OK) return rc;
        plan->last_variant = 2;
        return MCS_OK;
    }
    StitchArgs a;
    fill_args(plan, a, src, src_pitch_bytes, src_frame_stride, n_frames, dst, dst_pitch_bytes,
              dst_frame_stride);
    MCS_CHECK_CUDA(launch_gather<false>(plan, a, nullptr, stream));
    plan->last_variant = 1;
    return MCS_OK;
}

extern "C" int mcs_plan_owned_pixels(const mcs_plan* plan, int64_t* owned_host, void* cuda_stream) {
    MCS_CHECK_ARG(plan != nullptr && owned_host != nullptr, "mcs_plan_owned_pixels: NULL argument");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    unsigned long long* d = nullptr;
    MCS_CHECK_CUDA(cudaMalloc(&d, sizeof(unsigned long long) * MCS_MAX_LAYERS));
    int rc = MCS_OK;
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(unsigned long long) * MCS_MAX_LAYERS, stream);
    if (e == cudaSuccess && plan->out_w > 0 && plan->out_h > 0) {
        StitchArgs a;
        fill_args(plan, a, nullptr, nullptr, nullptr, 1, nullptr, 0, 0);
        e = launch_gather<true>(plan, a, d, stream);
    }
    unsigned long long h[MCS_MAX_LAYERS];
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        mcs_set_error("mcs_plan_owned_pixels: %s", cudaGetErrorString(e));
        rc = MCS_ERR_CUDA;
    } else {
        for (int k = 0; k < plan->n_layers; ++k) owned_host[k] = (int64_t)h[k];
    }
    cudaFree(d);
    return rc;
}
