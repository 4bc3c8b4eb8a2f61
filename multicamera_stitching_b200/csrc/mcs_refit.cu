// Host end of cv2.findHomography(..., RANSAC) as StitcherBase.matchKeypoints calls it (StitcherClass.py:443-444):
// after the RANSAC loop OpenCV refits the winning model on its inliers (normalised DLT, the structure of
// HomographyEstimatorCallback::runKernel) and polishes it with Levenberg-Marquardt on the reprojection error
// (HomographyRefineCallback).  The GPU scores the hypotheses (mcs_ransac.cu); this is the refit, in plain host
// C++ inside the library so that the recalibration call carries no numpy loop: a few tens of microseconds per
// pair instead of ~0.6 ms of small-array numpy calls.  float64 throughout, single pass over the points per
// accumulation, no allocation proportional to anything but the inlier count.
#include "mcs_common.h"

#include <exception>
#include <math.h>
#include <string.h>
#include <vector>

namespace {

struct Norm {   // p -> (p - c) * s per axis
    double cx, cy, sx, sy;
};

Norm normalise(const std::vector<double>& p) {   // p = x0, y0, x1, y1, ...
    const size_t n = p.size() / 2;
    Norm t = {0, 0, 1, 1};
    for (size_t i = 0; i < n; ++i) { t.cx += p[2 * i]; t.cy += p[2 * i + 1]; }
    t.cx /= (double)n;
    t.cy /= (double)n;
    double dx = 0, dy = 0;
    for (size_t i = 0; i < n; ++i) { dx += fabs(p[2 * i] - t.cx); dy += fabs(p[2 * i + 1] - t.cy); }
    dx /= (double)n;
    dy /= (double)n;
    t.sx = dx > 1e-12 ? 1.0 / dx : 1.0;
    t.sy = dy > 1e-12 ? 1.0 / dy : 1.0;
    return t;
}

// Eigenvector of the smallest eigenvalue of the symmetric 9 x 9 matrix m (cyclic Jacobi rotations).
void smallest_eigenvector9(double m[9][9], double v_out[9]) {
    double v[9][9];
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, diag = 0;
        for (int i = 0; i < 9; ++i) {
            diag += m[i][i] * m[i][i];
            for (int j = i + 1; j < 9; ++j) off += m[i][j] * m[i][j];
        }
        if (off <= 1e-30 * diag || off == 0.0) break;
        for (int p = 0; p < 8; ++p)
            for (int q = p + 1; q < 9; ++q) {
                if (m[p][q] == 0.0) continue;
                const double theta = (m[q][q] - m[p][p]) / (2.0 * m[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 9; ++k) {   // columns p, q
                    const double a = m[k][p], b = m[k][q];
                    m[k][p] = c * a - s * b;
                    m[k][q] = s * a + c * b;
                }
                for (int k = 0; k < 9; ++k) {   // rows p, q
                    const double a = m[p][k], b = m[q][k];
                    m[p][k] = c * a - s * b;
                    m[q][k] = s * a + c * b;
                }
                for (int k = 0; k < 9; ++k) {
                    const double a = v[k][p], b = v[k][q];
                    v[k][p] = c * a - s * b;
                    v[k][q] = s * a + c * b;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < 9; ++i)
        if (m[i][i] < m[best][best]) best = i;
    for (int k = 0; k < 9; ++k) v_out[k] = v[k][best];
}

// Least-squares homography a -> b, normalised DLT.  Returns false when the result is not finite.
bool fit_dlt(const std::vector<double>& a, const std::vector<double>& b, double H[9]) {
    const size_t n = a.size() / 2;
    const Norm ta = normalise(a), tb = normalise(b);
    double M[9][9];
    memset(M, 0, sizeof(M));
    for (size_t i = 0; i < n; ++i) {
        const double x = (a[2 * i] - ta.cx) * ta.sx, y = (a[2 * i + 1] - ta.cy) * ta.sy;
        const double u = (b[2 * i] - tb.cx) * tb.sx, v = (b[2 * i + 1] - tb.cy) * tb.sy;
        // rows (x, y, 1, 0, 0, 0, -ux, -uy, -u) and (0, 0, 0, x, y, 1, -vx, -vy, -v): only their non-zero products
        const double e[3] = {x, y, 1.0};
        const double r1[3] = {-u * x, -u * y, -u}, r2[3] = {-v * x, -v * y, -v};
        for (int p = 0; p < 3; ++p) {
            for (int q = p; q < 3; ++q) {
                const double ee = e[p] * e[q];
                M[p][q] += ee;
                M[3 + p][3 + q] += ee;
                M[6 + p][6 + q] += r1[p] * r1[q] + r2[p] * r2[q];
            }
            for (int q = 0; q < 3; ++q) {
                M[p][6 + q] += e[p] * r1[q];
                M[3 + p][6 + q] += e[p] * r2[q];
            }
        }
    }
    for (int p = 0; p < 9; ++p)
        for (int q = 0; q < p; ++q) M[p][q] = M[q][p];
    double h[9];
    smallest_eigenvector9(M, h);
    // H = inv(Tb) * Hn * Ta,  Ta = [sx 0 -cx sx; 0 sy -cy sy; 0 0 1],  inv(Tb) = [1/sx 0 cx; 0 1/sy cy; 0 0 1]
    double G[9];   // Hn * Ta
    for (int r = 0; r < 3; ++r) {
        G[3 * r + 0] = h[3 * r + 0] * ta.sx;
        G[3 * r + 1] = h[3 * r + 1] * ta.sy;
        G[3 * r + 2] = -h[3 * r + 0] * ta.cx * ta.sx - h[3 * r + 1] * ta.cy * ta.sy + h[3 * r + 2];
    }
    for (int c = 0; c < 3; ++c) {
        H[0 + c] = G[0 + c] / tb.sx + tb.cx * G[6 + c];
        H[3 + c] = G[3 + c] / tb.sy + tb.cy * G[6 + c];
        H[6 + c] = G[6 + c];
    }
    const double w = H[8];
    for (int k = 0; k < 9; ++k) {
        H[k] /= w;
        if (!isfinite(H[k])) return false;
    }
    return true;
}

double residual_cost(const double h[8], const std::vector<double>& a, const std::vector<double>& b) {
    const size_t n = a.size() / 2;
    double cost = 0;
    for (size_t i = 0; i < n; ++i) {
        const double x = a[2 * i], y = a[2 * i + 1];
        const double w = h[6] * x + h[7] * y + 1.0;
        const double ex = (h[0] * x + h[1] * y + h[2]) / w - b[2 * i], ey = (h[3] * x + h[4] * y + h[5]) / w - b[2 * i + 1];
        cost += ex * ex + ey * ey;
    }
    return cost;
}

// 8 x 8 linear solve, Gaussian elimination with partial pivoting; false when singular.
bool solve8(double A[8][8], double rhs[8], double x[8]) {
    for (int c = 0; c < 8; ++c) {
        int piv = c;
        for (int r = c + 1; r < 8; ++r)
            if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (!(fabs(A[piv][c]) > 0.0) || !isfinite(A[piv][c])) return false;
        if (piv != c) {
            for (int k = 0; k < 8; ++k) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
            const double t = rhs[c]; rhs[c] = rhs[piv]; rhs[piv] = t;
        }
        for (int r = c + 1; r < 8; ++r) {
            const double f = A[r][c] / A[c][c];
            if (f == 0.0) continue;
            for (int k = c; k < 8; ++k) A[r][k] -= f * A[c][k];
            rhs[r] -= f * rhs[c];
        }
    }
    for (int r = 7; r >= 0; --r) {
        double s = rhs[r];
        for (int k = r + 1; k < 8; ++k) s -= A[r][k] * x[k];
        x[r] = s / A[r][r];
        if (!isfinite(x[r])) return false;
    }
    return true;
}

// Levenberg-Marquardt over the 8 free parameters (h22 = 1) on the reprojection error.
void refine_lm(double H[9], const std::vector<double>& a, const std::vector<double>& b, int iters) {
    const size_t n = a.size() / 2;
    double h[8];
    for (int k = 0; k < 8; ++k) h[k] = H[k] / H[8];
    double lam = 1e-3;
    double cost = residual_cost(h, a, b);
    for (int it = 0; it < iters; ++it) {
        double JtJ[8][8], g[8];
        memset(JtJ, 0, sizeof(JtJ));
        memset(g, 0, sizeof(g));
        for (size_t i = 0; i < n; ++i) {
            const double x = a[2 * i], y = a[2 * i + 1];
            const double w = h[6] * x + h[7] * y + 1.0, iw = 1.0 / w;
            const double px = (h[0] * x + h[1] * y + h[2]) * iw, py = (h[3] * x + h[4] * y + h[5]) * iw;
            const double ex = px - b[2 * i], ey = py - b[2 * i + 1];
            const double xw = x * iw, yw = y * iw;
            // Jacobian rows (xw, yw, iw, 0, 0, 0, -xw px, -yw px) and (0, 0, 0, xw, yw, iw, -xw py, -yw py)
            const double e[3] = {xw, yw, iw};
            const double a1[2] = {-xw * px, -yw * px}, a2[2] = {-xw * py, -yw * py};
            for (int p = 0; p < 3; ++p) {
                g[p] += e[p] * ex;
                g[3 + p] += e[p] * ey;
                for (int q = p; q < 3; ++q) {
                    const double ee = e[p] * e[q];
                    JtJ[p][q] += ee;
                    JtJ[3 + p][3 + q] += ee;
                }
                for (int q = 0; q < 2; ++q) {
                    JtJ[p][6 + q] += e[p] * a1[q];
                    JtJ[3 + p][6 + q] += e[p] * a2[q];
                }
            }
            for (int p = 0; p < 2; ++p) {
                g[6 + p] += a1[p] * ex + a2[p] * ey;
                for (int q = p; q < 2; ++q) JtJ[6 + p][6 + q] += a1[p] * a1[q] + a2[p] * a2[q];
            }
        }
        for (int p = 0; p < 8; ++p)
            for (int q = 0; q < p; ++q) JtJ[p][q] = JtJ[q][p];
        bool improved = false;
        for (int attempt = 0; attempt < 6; ++attempt) {
            double A[8][8], rhs[8], step[8], h2[8];
            memcpy(A, JtJ, sizeof(A));
            for (int p = 0; p < 8; ++p) {
                A[p][p] += lam * JtJ[p][p];
                rhs[p] = -g[p];
            }
            if (!solve8(A, rhs, step)) {
                lam *= 10;
                continue;
            }
            for (int p = 0; p < 8; ++p) h2[p] = h[p] + step[p];
            const double cost2 = residual_cost(h2, a, b);
            if (cost2 < cost) {
                const bool converged = cost - cost2 <= 1e-9 * cost;   // the fit has stopped moving
                memcpy(h, h2, sizeof(h));
                cost = cost2;
                lam = lam * 0.1 > 1e-12 ? lam * 0.1 : 1e-12;
                improved = !converged;
                break;
            }
            lam *= 10;
        }
        if (!improved || cost < 1e-18) break;
    }
    for (int k = 0; k < 8; ++k) H[k] = h[k];
    H[8] = 1.0;
}

}  // namespace

extern "C" int mcs_refit_homography(const float* pts_a, const float* pts_b, const uint8_t* inlier_mask, int n,
                                    const double* h0, int lm_iters, double* h_out) {
    MCS_CHECK_ARG(n >= 0 && lm_iters >= 0, "mcs_refit_homography: negative count");
    MCS_CHECK_ARG(h0 != nullptr && h_out != nullptr, "mcs_refit_homography: NULL homography");
    MCS_CHECK_ARG(n == 0 || (pts_a != nullptr && pts_b != nullptr), "mcs_refit_homography: NULL points");
    try {   // std::vector may throw; nothing crosses the C boundary
    std::vector<double> a, b;
    a.reserve(2 * (size_t)n);
    b.reserve(2 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        if (inlier_mask && !inlier_mask[i]) continue;
        a.push_back((double)pts_a[2 * i]);
        a.push_back((double)pts_a[2 * i + 1]);
        b.push_back((double)pts_b[2 * i]);
        b.push_back((double)pts_b[2 * i + 1]);
    }
    double H[9];
    memcpy(H, h0, sizeof(H));
    const size_t m = a.size() / 2;
    if (m >= 4) {
        double F[9];
        if (m > 4 && fit_dlt(a, b, F)) memcpy(H, F, sizeof(H));
        if (H[8] != 0.0 && isfinite(H[8])) refine_lm(H, a, b, lm_iters);
    }
    memcpy(h_out, H, sizeof(H));
    return MCS_OK;
    } catch (const std::exception& e) {
        mcs_set_error("mcs_refit_homography: %s", e.what());
        return MCS_ERR_NOMEM;
    }
}
