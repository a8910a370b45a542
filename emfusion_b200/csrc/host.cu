// host.cu -- host-only helpers of the C ABI.
#include "common.cuh"
#include <math.h>

extern "C" EMF_API const char* emf_version(void) { return "emf_b200 0.1.0 sm_100a"; }

// Screen rectangle of the raycast box (+-((R-1) intdiv 2) * voxel, reference
// src/core/cuda/TSDF.cu:490) seen from the camera.  T_co maps camera -> volume, so a box
// corner c sits at R^T (c - t) in the camera frame.  The rectangle is padded by 2 px; rays
// outside it cannot pass the slab test, so skipping them cannot change any result.
namespace emfb {
// screen rectangle (padded by 2 px) of the box [-b, b] of a volume seen from the camera; full frame if a corner is behind it
void box_screen_rect(const double b[3], const emf_pose* T_co, const float K[9], int width, int height, int rect_out[4]);
}
extern "C" EMF_API int emf_volume_screen_rect(const int res[3], float voxel_size, const emf_pose* T_co, const float K[9],
                                      int width, int height, int rect_out[4]) {
    if (!res || !T_co || !K || !rect_out || width <= 0 || height <= 0) return EMF_ERR_INVALID;
    const double b[3] = {(double)((res[0] - 1) / 2) * voxel_size, (double)((res[1] - 1) / 2) * voxel_size,
                         (double)((res[2] - 1) / 2) * voxel_size};
    emfb::box_screen_rect(b, T_co, K, width, height, rect_out);
    return EMF_OK;
}

void emfb::box_screen_rect(const double b[3], const emf_pose* T_co, const float K[9], int width, int height, int rect_out[4]) {
    double x0 = 1e30, y0 = 1e30, x1 = -1e30, y1 = -1e30;
    bool full = false;
    for (int c = 0; c < 8 && !full; ++c) {
        const double p[3] = {(c & 1 ? b[0] : -b[0]) - T_co->t[0], (c & 2 ? b[1] : -b[1]) - T_co->t[1],
                             (c & 4 ? b[2] : -b[2]) - T_co->t[2]};
        double q[3];
        for (int k = 0; k < 3; ++k) q[k] = T_co->R[k] * p[0] + T_co->R[3 + k] * p[1] + T_co->R[6 + k] * p[2];
        if (q[2] < 1e-3) { full = true; break; }
        const double u = K[0] * q[0] + K[1] * q[1] + K[2] * q[2];
        const double v = K[3] * q[0] + K[4] * q[1] + K[5] * q[2];
        const double wq = K[6] * q[0] + K[7] * q[1] + K[8] * q[2];
        if (wq < 1e-6) { full = true; break; }
        const double px = u / wq, py = v / wq;
        x0 = fmin(x0, px); x1 = fmax(x1, px); y0 = fmin(y0, py); y1 = fmax(y1, py);
    }
    if (full) { rect_out[0] = 0; rect_out[1] = 0; rect_out[2] = width; rect_out[3] = height; return; }
    int ix0 = (int)floor(x0) - 2, iy0 = (int)floor(y0) - 2, ix1 = (int)ceil(x1) + 3, iy1 = (int)ceil(y1) + 3;
    if (ix0 < 0) ix0 = 0;
    if (iy0 < 0) iy0 = 0;
    if (ix1 > width) ix1 = width;
    if (iy1 > height) iy1 = height;
    if (ix1 < ix0) ix1 = ix0;
    if (iy1 < iy0) iy1 = iy0;
    rect_out[0] = ix0; rect_out[1] = iy0; rect_out[2] = ix1; rect_out[3] = iy1;
}
