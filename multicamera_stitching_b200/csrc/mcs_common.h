// Internal definitions shared by the translation units of libmcs_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>

#include "mcs.h"

#define MCS_ABI_VERSION 1

// ---- error plumbing -------------------------------------------------------
void mcs_set_error(const char* fmt, ...);
void mcs_count_launch(int n);

#define MCS_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            mcs_set_error(__VA_ARGS__);          \
            return MCS_ERR_INVALID;              \
        }                                        \
    } while (0)

#define MCS_CHECK_CUDA(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            mcs_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                          __FILE__, __LINE__);                                      \
            return MCS_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

// ---- compositing plan -----------------------------------------------------
// One camera of the panorama, in the form the kernels consume.
struct McsLayer {
    double mi[9];        // inverse homography (output-canvas frame -> source), cv::invert closed form
    int rx0, ry0, rx1, ry1;  // visible rectangle, output coordinates, half open
    int ox, oy;          // origin of the layer's canvas frame in output coordinates
    int src_w, src_h;
    int kind;            // MCS_LAYER_*
    int affine;          // mi[6] == mi[7] == 0: the perspective divide is a per-plan constant
};

// Work item of the tiled kernel: one 64 x 16 output tile (host-built at plan creation).
struct McsTile {
    int x0, y0;          // output coordinates of the tile origin
    short w, h;          // valid extent (clipped at the panorama border)
    short layer;         // owner layer, or -1 = background, -2 = mixed ownership
    short cls;           // MCS_TILE_*
    int sx0, sy0;        // source-pixel origin of the staged box (WARP_STAGED)
    short sw, sh;        // staged box extent in source pixels
    int pad;
};

#define MCS_TILE_W 64
#define MCS_TILE_H 16

#define MCS_TILE_BACKGROUND 0
#define MCS_TILE_COPY 1
#define MCS_TILE_WARP 2
#define MCS_TILE_MIXED 3

struct mcs_plan {
    int n_layers;
    int channels;
    int out_w, out_h;
    int device;
    McsLayer layers[MCS_MAX_LAYERS];
    int last_variant;
};

// 3x3 float64 inverse with cv::invert's association (host, no FMA contraction).
bool mcs_invert3x3(const double* m, double* out);
