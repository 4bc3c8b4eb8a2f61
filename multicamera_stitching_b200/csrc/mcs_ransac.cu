// Batched RANSAC homography scoring (sm_100a): one warp per hypothesis.
//
// Replaces the hypothesis loop inside cv2.findHomography(ptsA, ptsB, cv2.RANSAC, thresh)
// (StitcherClass.py:443-444): every warp solves the 4-point homography of its minimal sample
// in closed form (unit square -> quad factorisation, float64, evaluated redundantly by all
// lanes so no broadcast is needed), then the 32 lanes stride over the matched points counting
// reprojection inliers in float32 like OpenCV's computeError.  A second tiny kernel picks the
// winner per pair (lowest index among equal counts) and writes its inlier mask.
#include "mcs_common.h"

#include <math.h>

#define RANSAC_THREADS 256
#define RANSAC_WARPS (RANSAC_THREADS / 32)

// Projective map taking the unit square (0,0),(1,0),(1,1),(0,1) onto the quad p0..p3.
// Returns false for a degenerate quad (three collinear corners).
__device__ __forceinline__ bool square_to_quad(const double (&x)[4], const double (&y)[4], double (&m)[9]) {
    const double dx1 = x[1] - x[2], dx2 = x[3] - x[2], sx = x[0] - x[1] + x[2] - x[3];
    const double dy1 = y[1] - y[2], dy2 = y[3] - y[2], sy = y[0] - y[1] + y[2] - y[3];
    const double den = dx1 * dy2 - dy1 * dx2;
    const double scale = fabs(dx1) + fabs(dx2) + fabs(dy1) + fabs(dy2);
    if (!(fabs(den) > 1e-12 * scale * scale)) return false;
    const double g = (sx * dy2 - sy * dx2) / den;
    const double h = (dx1 * sy - dy1 * sx) / den;
    m[0] = x[1] - x[0] + g * x[1];
    m[1] = x[3] - x[0] + h * x[3];
    m[2] = x[0];
    m[3] = y[1] - y[0] + g * y[1];
    m[4] = y[3] - y[0] + h * y[3];
    m[5] = y[0];
    m[6] = g;
    m[7] = h;
    m[8] = 1.0;
    return true;
}

__device__ __forceinline__ double cross2(double ax, double ay, double bx, double by, double cx, double cy) {
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}

// H (A -> B) from four correspondences; false if the sample is degenerate.
__device__ bool homography_4pt(const double (&ax)[4], const double (&ay)[4], const double (&bx)[4],
                               const double (&by)[4], double (&H)[9]) {
    // OpenCV's HomographyEstimatorCallback::checkSubset rejects samples whose triples change
    // orientation between the two images (the map would fold the plane); collinear triples
    // are caught by the zero cross products as well.
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3, k = (i + 2) & 3;
        const double ca = cross2(ax[i], ay[i], ax[j], ay[j], ax[k], ay[k]);
        const double cb = cross2(bx[i], by[i], bx[j], by[j], bx[k], by[k]);
        if (!(ca * cb > 0.0)) return false;
    }
    double SA[9], SB[9];
    if (!square_to_quad(ax, ay, SA) || !square_to_quad(bx, by, SB)) return false;
    // H = SB * adj(SA)
    double adj[9];
    adj[0] = SA[4] * SA[8] - SA[5] * SA[7];
    adj[1] = SA[2] * SA[7] - SA[1] * SA[8];
    adj[2] = SA[1] * SA[5] - SA[2] * SA[4];
    adj[3] = SA[5] * SA[6] - SA[3] * SA[8];
    adj[4] = SA[0] * SA[8] - SA[2] * SA[6];
    adj[5] = SA[2] * SA[3] - SA[0] * SA[5];
    adj[6] = SA[3] * SA[7] - SA[4] * SA[6];
    adj[7] = SA[1] * SA[6] - SA[0] * SA[7];
    adj[8] = SA[0] * SA[4] - SA[1] * SA[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            H[3 * r + c] = SB[3 * r] * adj[c] + SB[3 * r + 1] * adj[3 + c] + SB[3 * r + 2] * adj[6 + c];
    const double h8 = H[8];
    if (!(fabs(h8) > 1e-300) || !isfinite(h8)) return false;
    const double inv = 1.0 / h8;
    bool finite = true;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        H[i] *= inv;
        finite = finite && isfinite(H[i]);
    }
    H[8] = 1.0;
    return finite;
}

__device__ __forceinline__ bool is_inlier(const float (&Hf)[9], float ax, float ay, float bx, float by,
                                          float thresh2) {
    const float ww = 1.0f / (Hf[6] * ax + Hf[7] * ay + 1.0f);
    const float dx = (Hf[0] * ax + Hf[1] * ay + Hf[2]) * ww - bx;
    const float dy = (Hf[3] * ax + Hf[4] * ay + Hf[5]) * ww - by;
    return dx * dx + dy * dy <= thresh2;
}

__global__ void __launch_bounds__(RANSAC_THREADS)
mcs_ransac_score_kernel(const float* __restrict__ ptsA, const float* __restrict__ ptsB,
                        const int32_t* __restrict__ n_arr, int n_max,
                        const int32_t* __restrict__ samples, int k, float thresh2,
                        int32_t* __restrict__ inlier_counts, double* __restrict__ H_k) {
    const int pair = blockIdx.y;
    const int hyp = blockIdx.x * RANSAC_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (hyp >= k) return;
    const int n = n_arr ? min(n_arr[pair], n_max) : n_max;
    const float2* A = reinterpret_cast<const float2*>(ptsA) + (size_t)pair * n_max;
    const float2* B = reinterpret_cast<const float2*>(ptsB) + (size_t)pair * n_max;
    const int32_t* smp = samples + ((size_t)pair * k + hyp) * 4;

    double ax[4], ay[4], bx[4], by[4], H[9];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int s = __ldg(smp + i);
        ok = ok && s >= 0 && s < n;
        const int sc = ok ? s : 0;
        const float2 a = (n > 0) ? __ldg(A + sc) : make_float2(0.f, 0.f);
        const float2 b = (n > 0) ? __ldg(B + sc) : make_float2(0.f, 0.f);
        ax[i] = a.x; ay[i] = a.y; bx[i] = b.x; by[i] = b.y;
    }
    ok = ok && homography_4pt(ax, ay, bx, by, H);

    int count = -1;
    if (ok) {  // warp-uniform: every lane evaluated the same sample
        float Hf[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Hf[i] = (float)H[i];
        int c = 0;
        for (int j = lane; j < n; j += 32) {
            const float2 a = __ldg(A + j), b = __ldg(B + j);
            c += is_inlier(Hf, a.x, a.y, b.x, b.y, thresh2) ? 1 : 0;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        count = c;
    }
    const size_t o = (size_t)pair * k + hyp;
    if (lane == 0) inlier_counts[o] = count;
    if (lane < 9) H_k[o * 9 + lane] = ok ? H[lane] : 0.0;
}

__global__ void __launch_bounds__(RANSAC_THREADS)
mcs_ransac_select_kernel(const float* __restrict__ ptsA, const float* __restrict__ ptsB,
                         const int32_t* __restrict__ n_arr, int n_max, int k, float thresh2,
                         const int32_t* __restrict__ inlier_counts, const double* __restrict__ H_k,
                         int32_t* __restrict__ best_idx, uint8_t* __restrict__ best_mask) {
    __shared__ long long s_key[RANSAC_WARPS];
    __shared__ int s_best;
    const int pair = blockIdx.x;
    const int n = n_arr ? min(n_arr[pair], n_max) : n_max;
    // argmax(count) with lowest index on ties == max of (count << 32 | ~index)
    long long key = -1;
    for (int h = threadIdx.x; h < k; h += RANSAC_THREADS) {
        const int c = inlier_counts[(size_t)pair * k + h];
        if (c >= 0) {
            const long long cand = ((long long)c << 32) | (long long)(0x7fffffff - h);
            key = cand > key ? cand : key;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const long long o = __shfl_xor_sync(0xffffffffu, key, off);
        key = o > key ? o : key;
    }
    if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long b = -1;
        for (int w = 0; w < RANSAC_WARPS; ++w) b = s_key[w] > b ? s_key[w] : b;
        s_best = b < 0 ? -1 : (int)(0x7fffffff - (int)(b & 0x7fffffff));
        best_idx[pair] = s_best;
    }
    __syncthreads();
    const int best = s_best;
    float Hf[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Hf[i] = best >= 0 ? (float)H_k[((size_t)pair * k + best) * 9 + i] : 0.f;
    const float2* A = reinterpret_cast<const float2*>(ptsA) + (size_t)pair * n_max;
    const float2* B = reinterpret_cast<const float2*>(ptsB) + (size_t)pair * n_max;
    for (int j = threadIdx.x; j < n_max; j += RANSAC_THREADS) {
        uint8_t m = 0;
        if (best >= 0 && j < n) {
            const float2 a = A[j], b = B[j];
            m = is_inlier(Hf, a.x, a.y, b.x, b.y, thresh2) ? 1 : 0;
        }
        best_mask[(size_t)pair * n_max + j] = m;
    }
}

extern "C" int mcs_ransac_homography(const float* ptsA, const float* ptsB, const int32_t* n,
                                     int n_max, const int32_t* samples, int k, float reproj_thresh,
                                     int32_t* inlier_counts, double* H_k, int32_t* best_idx,
                                     uint8_t* best_mask, int batch, void* cuda_stream) {
    MCS_CHECK_ARG(batch >= 0 && batch <= 65535, "mcs_ransac_homography: batch=%d outside 0..65535", batch);
    MCS_CHECK_ARG(n_max >= 0 && k >= 0, "mcs_ransac_homography: negative size");
    MCS_CHECK_ARG(reproj_thresh >= 0.f, "mcs_ransac_homography: negative threshold");
    if (batch == 0) return MCS_OK;
    MCS_CHECK_ARG(best_idx && (best_mask || n_max == 0), "mcs_ransac_homography: NULL output");
    MCS_CHECK_ARG(k == 0 || (samples && inlier_counts && H_k), "mcs_ransac_homography: NULL hypothesis buffer");
    MCS_CHECK_ARG(n_max == 0 || (ptsA && ptsB), "mcs_ransac_homography: NULL points");
    MCS_CHECK_ARG(((uintptr_t)ptsA & 7) == 0 && ((uintptr_t)ptsB & 7) == 0,
                  "mcs_ransac_homography: point buffers must be 8-byte aligned");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const float t2 = reproj_thresh * reproj_thresh;
    if (k > 0) {
        dim3 grid((k + RANSAC_WARPS - 1) / RANSAC_WARPS, batch, 1);
        mcs_ransac_score_kernel<<<grid, RANSAC_THREADS, 0, stream>>>(ptsA, ptsB, n, n_max, samples, k, t2,
                                                                     inlier_counts, H_k);
        mcs_count_launch(1);
        MCS_CHECK_CUDA(cudaGetLastError());
    }
    mcs_ransac_select_kernel<<<batch, RANSAC_THREADS, 0, stream>>>(ptsA, ptsB, n, n_max, k, t2, inlier_counts,
                                                                   H_k, best_idx, best_mask);
    mcs_count_launch(1);
    MCS_CHECK_CUDA(cudaGetLastError());
    return MCS_OK;
}
