/*
 * emf_oracle.c -- CPU restatement of EM-Fusion's dense hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under emfusion_b200/ (the product) may
 * include, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker or
 * as the timed CPU baseline.
 *
 * PARITY STATUS: the reference (EmbodiedVision/emfusion @ ab56b83) ships no
 * tests, golden vectors or fixtures for this path, so
 * this oracle is pinned the only way available: against the reference's own
 * CUDA kernels, compiled unchanged from /root/reference against a type shim
 * (oracle/Makefile -> oracle/_ref/libemf_ref.so) and run on a B200
 * (tests/test_gpu_parity.py::test_oracle_vs_reference_kernels).  Until that
 * test has run green on a GPU box the oracle is "parity unpinned".
 *
 * Arithmetic convention.  The reference kernels are compiled by nvcc with
 * default flags (FMA contraction on, IEEE div/sqrt, no FTZ).  Where nvcc
 * 12.9 / ptxas for sm_100a contracts a multiply-add in the reference build,
 * this file calls fmaf() at exactly that place (patterns read off the PTX
 * and SASS of the reference build; see DESIGN.md "Canonical arithmetic"), and
 * it is compiled with -ffp-contract=off so the host compiler adds none of its
 * own.  Building with -DEMFO_NOFMA evaluates every multiply-add unfused,
 * which bounds the sensitivity of the results to contraction.
 *
 * Layout (all volumes, reference src/core/TSDF.cpp:35-42): continuous float
 * array, element (z*Ry + y, x), x fastest.  Images are continuous row-major.
 *
 * Each function cites the reference file:line it follows
 * (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <limits.h>

#ifdef EMFO_NOFMA
#define FMA(a, b, c) ((a) * (b) + (c))
#else
#define FMA(a, b, c) fmaf((a), (b), (c))
#endif

#define EMFO_API __attribute__((visibility("default")))

/* cvt.rni.s32.f32: round to nearest even, saturating, NaN -> 0
 * (__float2int_rn, reference src/core/cuda/TSDF.cu:362-363). */
static inline int f2i_rn(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT_MAX;
    if (v <= -2147483648.0f) return INT_MIN;
    return (int)lrintf(v); /* default rounding mode = nearest even */
}

/* M*v as the reference build evaluates it at most sites ("yxz" pattern):
 * fma(m2, z, fma(m0, x, m1*y))   (include/EMFusion/core/cuda/common.cuh:94-107) */
static inline float dot_yxz(const float* m, float x, float y, float z) {
    return FMA(m[2], z, FMA(m[0], x, m[1] * y));
}
/* ... and as it evaluates it in kernel_getVolumeVals ("xyz" pattern):
 * fma(m2, z, fma(m1, y, m0*x)) */
static inline float dot_xyz(const float* m, float x, float y, float z) {
    return FMA(m[2], z, FMA(m[1], y, m[0] * x));
}
static inline float norm3(float x, float y, float z) {
    return sqrtf(FMA(z, z, FMA(x, x, y * y)));
}

/* interpolateTrilinear, include/EMFusion/core/cuda/TSDF.cuh:65-97.
 * lerp order x, then y, then z; each lerp is fma(1-a, lo, a*hi). */
static inline float trilinear(const float* vol, int rx, int ry, float vx, float vy, float vz) {
    const int lx = (int)vx, ly = (int)vy, lz = (int)vz;
    const float ax = vx - (float)lx, ay = vy - (float)ly, az = vz - (float)lz;
    const float bx = 1.0f - ax, by = 1.0f - ay, bz = 1.0f - az;
    const float* r00 = vol + ((int64_t)lz * ry + ly) * rx + lx;
    const float* r01 = r00 + rx;
    const float* r10 = r00 + (int64_t)ry * rx;
    const float* r11 = r10 + rx;
    const float c00 = FMA(bx, r00[0], ax * r00[1]);
    const float c01 = FMA(bx, r01[0], ax * r01[1]);
    const float c10 = FMA(bx, r10[0], ax * r10[1]);
    const float c11 = FMA(bx, r11[0], ax * r11[1]);
    const float d0 = FMA(by, c00, ay * c01);
    const float d1 = FMA(by, c10, ay * c11);
    return FMA(bz, d0, az * d1);
}
static inline void trilinear3(const float* vol3, int rx, int ry, float vx, float vy, float vz, float out[3]) {
    const int lx = (int)vx, ly = (int)vy, lz = (int)vz;
    const float ax = vx - (float)lx, ay = vy - (float)ly, az = vz - (float)lz;
    const float bx = 1.0f - ax, by = 1.0f - ay, bz = 1.0f - az;
    const float* r00 = vol3 + 3 * (((int64_t)lz * ry + ly) * rx + lx);
    const float* r01 = r00 + 3 * (int64_t)rx;
    const float* r10 = r00 + 3 * (int64_t)ry * rx;
    const float* r11 = r10 + 3 * (int64_t)rx;
    for (int c = 0; c < 3; ++c) {
        const float c00 = FMA(bx, r00[c], ax * r00[3 + c]);
        const float c01 = FMA(bx, r01[c], ax * r01[3 + c]);
        const float c10 = FMA(bx, r10[c], ax * r10[3 + c]);
        const float c11 = FMA(bx, r11[c], ax * r11[3 + c]);
        const float d0 = FMA(by, c00, ay * c01);
        const float d1 = FMA(by, c10, ay * c11);
        out[c] = FMA(bz, d0, az * d1);
    }
}

/* ------------------------------------------------------------------------
 * A.1 points: src/core/cuda/EMFusion.cu:29-61 (kernel_computePoints).
 * p = ((x-cx)*d/fx, (y-cy)*d/fy, d); mul then IEEE div, no contraction.
 * ---------------------------------------------------------------------- */
EMFO_API void emfo_compute_points(const float* depth, int w, int h, const float* K, float* points) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float d = depth[(size_t)y * w + x];
            float* p = points + 3 * ((size_t)y * w + x);
            p[0] = ((float)x - K[2]) * d / K[0];
            p[1] = ((float)y - K[5]) * d / K[4];
            p[2] = d;
        }
}

/* ------------------------------------------------------------------------
 * A.2 integrate: src/core/cuda/TSDF.cu:327-401 (kernel_updateTSDF); the
 * caller passes T_OC = cam_pose^-1 * pose (src/core/TSDF.cpp:112).
 * counts (optional, 6 x int64): [0] updated, [1] marked occluded-unseen (-1),
 * [2] occluded but already seen (no write), [3] invalid-depth / behind-camera
 * voxels that read the weight, [4] outside image, [5] zero-weight-sum skips.
 * ---------------------------------------------------------------------- */
EMFO_API void emfo_update_tsdf(const float* depth, const float* assoc, int w, int h,
                               float* tsdf, float* weights,
                               const float* R, const float* t, const float* K,
                               const int* res, float voxel, float trunc, float maxw,
                               int64_t* counts) {
    const int rx = res[0], ry = res[1], rz = res[2];
    int64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
#pragma omp parallel for schedule(static) reduction(+ : c0, c1, c2, c3, c4, c5)
    for (int yz = 0; yz < ry * rz; ++yz) {
        const int y = yz % ry, z = yz / ry;
        /* (i - (R-1)/2.f) * voxel; (R-1)*0.5 is exact in fp32 */
        const float oy = ((float)y - (float)(ry - 1) * 0.5f) * voxel;
        const float oz = ((float)z - (float)(rz - 1) * 0.5f) * voxel;
        float* trow = tsdf + (size_t)yz * rx;
        float* wrow = weights + (size_t)yz * rx;
        for (int x = 0; x < rx; ++x) {
            const float ox = ((float)x - (float)(rx - 1) * 0.5f) * voxel;
            const float pcx = t[0] + dot_yxz(R + 0, ox, oy, oz);
            const float pcy = t[1] + dot_yxz(R + 3, ox, oy, oz);
            const float pcz = t[2] + dot_yxz(R + 6, ox, oy, oz);
            if (!(pcz > 0.0f)) { /* :349 */
                if (wrow[x] == 0.0f) trow[x] = 0.0f;
                ++c3;
                continue;
            }
            const float qx = dot_yxz(K + 0, pcx, pcy, pcz);
            const float qy = dot_yxz(K + 3, pcx, pcy, pcz);
            const float qz = dot_yxz(K + 6, pcx, pcy, pcz);
            const int px = f2i_rn(qx / qz), py = f2i_rn(qy / qz); /* :362 */
            if (px < 0 || px >= w || py < 0 || py >= h) { ++c4; continue; }
            const float d = depth[(size_t)py * w + px];
            if (!(d > 0.0f)) { /* :369 */
                if (wrow[x] == 0.0f) trow[x] = 0.0f;
                ++c3;
                continue;
            }
            const float lx = ((float)px - K[2]) / K[0];
            const float ly = ((float)py - K[5]) / K[4];
            const float lambda = sqrtf(FMA(lx, lx, ly * ly) + 1.0f);
            const float inv_lambda = 1.0f / lambda;
            const float nrm = norm3(pcx, pcy, pcz);
            /* depthVal - (1/lambda)*norm: ptxas fuses this into one FFMA */
            const float sdf = FMA(-nrm, inv_lambda, d);
            const float wp = wrow[x];
            if (sdf >= -trunc) { /* :384 */
                const float v = copysignf(fminf(1.0f, fabsf(sdf / trunc)), sdf);
                const float tp = trow[x];
                const float a = (sdf < trunc) ? assoc[(size_t)py * w + px] : 1.0f;
                const float ws = wp + a;
                if (ws > 0.0f) {
                    trow[x] = FMA(wp, tp, v * a) / ws;
                    wrow[x] = fminf(ws, maxw);
                    ++c0;
                } else {
                    ++c5;
                }
            } else if (wp == 0.0f) {
                trow[x] = -1.0f;
                ++c1;
            } else {
                ++c2;
            }
        }
    }
    if (counts) { counts[0] = c0; counts[1] = c1; counts[2] = c2; counts[3] = c3; counts[4] = c4; counts[5] = c5; }
}

/* ------------------------------------------------------------------------
 * A.3 gradient: src/core/TSDF.cpp:120-123 (setTo 0) +
 * src/core/cuda/TSDF.cu:429-448 (kernel_computeTSDFGrads).
 * Forward differences, interior only, last plane of each axis stays 0.
 * ---------------------------------------------------------------------- */
EMFO_API void emfo_compute_grads(const float* tsdf, float* grads, const int* res) {
    const int rx = res[0], ry = res[1], rz = res[2];
#pragma omp parallel for schedule(static)
    for (int yz = 0; yz < ry * rz; ++yz) {
        const int y = yz % ry, z = yz / ry;
        const float* row = tsdf + (size_t)yz * rx;
        float* g = grads + 3 * (size_t)yz * rx;
        for (int x = 0; x < rx; ++x) {
            if (x >= rx - 1 || y >= ry - 1 || z >= rz - 1) {
                g[3 * x + 0] = g[3 * x + 1] = g[3 * x + 2] = 0.0f;
            } else {
                const float f = row[x];
                g[3 * x + 0] = row[x + 1] - f;
                g[3 * x + 1] = row[x + rx] - f;
                g[3 * x + 2] = row[x + (size_t)ry * rx] - f;
            }
        }
    }
}

/* ------------------------------------------------------------------------
 * A.4 raycast: src/core/cuda/TSDF.cu:466-573 (kernel_raycastTSDF) with
 * enterVolStep/exitVolStep include/EMFusion/core/cuda/TSDF.cuh:31-63; the
 * caller passes T_CO = pose^-1 * cam_pose (src/core/TSDF.cpp:162).
 * raylengths is in/out (non-zero input = far clip, :496-500); vertices,
 * normals, mask are written at hit pixels only (callers pre-clear).
 * hit_voxel (optional, h*w*3 int32): lo(v*) at hit pixels, untouched elsewhere.
 * steps (optional, int64[2]): [0] in-bounds march samples taken (each = one
 * tsdf + one weight trilinear in the reference), [1] rays that entered a box.
 * ---------------------------------------------------------------------- */
static inline int out_of(float vx, float vy, float vz, float pad, float frx, float fry, float frz) {
    return vx < 0.0f || vx + pad >= frx || vy < 0.0f || vy + pad >= fry || vz < 0.0f || vz + pad >= frz;
}

EMFO_API void emfo_raycast(const float* tsdf, const float* grads, const float* weights,
                           float* raylengths, float* vertices, float* normals, uint8_t* mask,
                           int w, int h, const float* R, const float* t, const float* K,
                           const int* res, float voxel, float trunc,
                           int32_t* hit_voxel, int64_t* steps, int32_t* step_img) {
    /* step_img (optional, w*h*3 int32): per pixel, in-bounds march samples taken with step == truncdist,
     * == voxel, == voxel/2 (measurement only: the roofline numerator and the divergence statistics) */
    const int rx = res[0], ry = res[1], rz = res[2];
    const float frx = (float)rx, fry = (float)ry, frz = (float)rz;
    /* boxBounds = (volSize - 1) / 2 * voxelSize with INTEGER division
     * (common.cuh:181-183; TSDF.cu:490) */
    const float bx = (float)((rx - 1) / 2) * voxel;
    const float by = (float)((ry - 1) / 2) * voxel;
    const float bz = (float)((rz - 1) / 2) * voxel;
    const float hx = (float)(rx - 1) * 0.5f, hy = (float)(ry - 1) * 0.5f, hz = (float)(rz - 1) * 0.5f;
    const float ox = t[0], oy = t[1], oz = t[2];
    const float half_voxel = voxel * 0.5f;
    int64_t n_steps = 0, n_enter = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : n_steps, n_enter)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const size_t pix = (size_t)y * w + x;
            const float ux = ((float)x - K[2]) / K[0];
            const float uy = ((float)y - K[5]) / K[4];
            /* rot_CO * (ux, uy, 1): mul(m1,uy); fma(m0,ux,.); add m2 */
            const float rayx = R[2] + FMA(R[0], ux, R[1] * uy);
            const float rayy = R[5] + FMA(R[3], ux, R[4] * uy);
            const float rayz = R[8] + FMA(R[6], ux, R[7] * uy);
            const float rn = norm3(rayx, rayy, rayz);
            const float dx = rayx / rn, dy = rayy / rn, dz = rayz / rn;
            const float tin = fmaxf(fmaxf(((dx > 0.0f ? -bx : bx) - ox) / dx,
                                          ((dy > 0.0f ? -by : by) - oy) / dy),
                                    ((dz > 0.0f ? -bz : bz) - oz) / dz);
            const float tout = fminf(fminf(((dx > 0.0f ? bx : -bx) - ox) / dx,
                                           ((dy > 0.0f ? by : -by) - oy) / dy),
                                     ((dz > 0.0f ? bz : -bz) - oz) / dz);
            float tcur = voxel + tin;
            float tmax = tout - voxel;
            const float old = raylengths[pix];
            if (old != 0.0f) tmax = fminf(old, tmax);
            if (tcur >= tmax) continue; /* also drops NaN-free misses; NaN compares false -> falls through like the GPU */
            ++n_enter;
            float step = trunc;
            float vx, vy, vz;
            for (;;) { /* coarse skip :509-515 */
                vx = hx + FMA(dx, tcur, ox) / voxel;
                vy = hy + FMA(dy, tcur, oy) / voxel;
                vz = hz + FMA(dz, tcur, oz) / voxel;
                if (out_of(vx, vy, vz, 1.0f, frx, fry, frz) && tcur < tmax) tcur = step + tcur;
                else break;
            }
            /* The reference reads the volume here even if the skip loop ran out
             * of ray (possible OOB read, :516); the march loop below then cannot
             * execute (tcur >= tmax, step > 0), so "no hit" is the defined result. */
            if (out_of(vx, vy, vz, 1.0f, frx, fry, frz)) continue;
            if (vx != vx || vy != vy || vz != vz) continue; /* 0/0 ray component: GPU marches NaNs to no hit */
            float f = trilinear(tsdf, rx, ry, vx, vy, vz);
            if (fabsf(f) < 1.0f) step = voxel;
            if (fabsf(f) < 0.8f) step = half_voxel;
            for (;;) {
                tcur = tcur + step;
                if (!(tcur <= tmax)) break;
                vx = hx + FMA(dx, tcur, ox) / voxel;
                vy = hy + FMA(dy, tcur, oy) / voxel;
                vz = hz + FMA(dz, tcur, oz) / voxel;
                if (out_of(vx, vy, vz, 2.0f, frx, fry, frz)) continue;
                ++n_steps;
                if (step_img) ++step_img[3 * pix + (step == trunc ? 0 : (step == voxel ? 1 : 2))];
                const float fn = trilinear(tsdf, rx, ry, vx, vy, vz);
                const float wn = trilinear(weights, rx, ry, vx, vy, vz);
                if (f < 0.0f && fn > 0.0f && wn > 0.0f) break; /* back face :532 */
                if (fabsf(fn) < 1.0f) step = voxel;
                if (fabsf(fn) < 0.8f) step = half_voxel;
                if (f > 0.0f && fn < 0.0f) { /* :540 */
                    /* t* uses the already-advanced tcur and already-adapted step (:543) */
                    const float ts = tcur - (f * step) / (fn - f);
                    const float mx = dx * ts, my = dy * ts, mz = dz * ts;
                    const float sx = hx + (ox + mx) / voxel;
                    const float sy = hy + (oy + my) / voxel;
                    const float sz = hz + (oz + mz) / voxel;
                    if (out_of(sx, sy, sz, 2.0f, frx, fry, frz)) continue; /* f NOT updated */
                    const float ws = trilinear(weights, rx, ry, sx, sy, sz);
                    if (ws > 0.0f) {
                        float g[3];
                        trilinear3(grads, rx, ry, sx, sy, sz, g);
                        raylengths[pix] = ts;
                        /* transpose(rot_CO) * (t* * dir) */
                        vertices[3 * pix + 0] = FMA(R[6], mz, FMA(R[0], mx, R[3] * my));
                        vertices[3 * pix + 1] = FMA(R[7], mz, FMA(R[1], mx, R[4] * my));
                        vertices[3 * pix + 2] = FMA(R[8], mz, FMA(R[2], mx, R[5] * my));
                        const float gn = norm3(g[0], g[1], g[2]);
                        const float nx = g[0] / gn, ny = g[1] / gn, nz = g[2] / gn;
                        normals[3 * pix + 0] = FMA(R[6], nz, FMA(R[0], nx, R[3] * ny));
                        normals[3 * pix + 1] = FMA(R[7], nz, FMA(R[1], nx, R[4] * ny));
                        normals[3 * pix + 2] = FMA(R[8], nz, FMA(R[2], nx, R[5] * ny));
                        mask[pix] = 1;
                        if (hit_voxel) {
                            hit_voxel[3 * pix + 0] = (int)sx;
                            hit_voxel[3 * pix + 1] = (int)sy;
                            hit_voxel[3 * pix + 2] = (int)sz;
                        }
                        break;
                    }
                }
                f = fn;
            }
        }
    if (steps) { steps[0] = n_steps; steps[1] = n_enter; }
}

/* ------------------------------------------------------------------------
 * gather: src/core/cuda/TSDF.cu:662-726 (kernel_getVolumeVals<float> +
 * launcher; output zero-filled first :705).  T_CO = pose^-1 * cam_pose.
 * in_bounds (optional int64*): number of pixels that gathered.
 * ---------------------------------------------------------------------- */
EMFO_API void emfo_get_volume_vals(const float* vol, const float* points, int w, int h,
                                   const float* R, const float* t, const int* res, float voxel,
                                   float* vals, int64_t* in_bounds) {
    const int rx = res[0], ry = res[1], rz = res[2];
    const float frx = (float)rx, fry = (float)ry, frz = (float)rz;
    int64_t nin = 0;
#pragma omp parallel for schedule(static) reduction(+ : nin)
    for (int i = 0; i < w * h; ++i) {
        vals[i] = 0.0f;
        const float* p = points + 3 * (size_t)i;
        if (p[2] <= 0.0f) continue;
        const float qx = t[0] + dot_xyz(R + 0, p[0], p[1], p[2]);
        const float qy = t[1] + dot_xyz(R + 3, p[0], p[1], p[2]);
        const float qz = t[2] + dot_xyz(R + 6, p[0], p[1], p[2]);
        /* p / voxelSize + (volSize-1)/2.f; the add is fma(R-1, 0.5, q) == q + exact half */
        const float vx = (float)(rx - 1) * 0.5f + qx / voxel;
        const float vy = (float)(ry - 1) * 0.5f + qy / voxel;
        const float vz = (float)(rz - 1) * 0.5f + qz / voxel;
        if (out_of(vx, vy, vz, 1.0f, frx, fry, frz)) continue;
        vals[i] = trilinear(vol, rx, ry, vx, vy, vz);
        ++nin;
    }
    if (in_bounds) *in_bounds = nin;
}

/* ------------------------------------------------------------------------
 * A.7 per-volume association.
 * Background: src/core/TSDF.cpp:125-156 (computeAssociation + computeLaplace).
 * Object (fg_probs != NULL): src/core/ObjTSDF.cpp:181-201.
 * Every step is a separate fp32 element-wise OpenCV-CUDA launch in the
 * reference (no contraction across them): abs, *(-trunc/sigma), expf,
 * *(1/(2 sigma)), [* fgProbVals], *alpha, +(1-alpha)*uniPrior, masked 0.
 * assoc_mask_out (optional) receives associationMask (tsdf gather == 0).
 * ---------------------------------------------------------------------- */
EMFO_API void emfo_assoc_volume(const float* tsdf, const float* fg_probs, const float* points, int w, int h,
                                const float* R, const float* t, const int* res, float voxel, float trunc,
                                float sigma, float alpha, float uni_prior,
                                float* assoc_out, uint8_t* assoc_mask_out, float* scratch) {
    const float k1 = -trunc / sigma;
    const float k2 = 1.0f / (2.0f * sigma);
    const float k3 = (1 - alpha) * uni_prior;
    float* fgv = scratch; /* w*h floats, only used for objects */
    emfo_get_volume_vals(tsdf, points, w, h, R, t, res, voxel, assoc_out, NULL);
    if (fg_probs) emfo_get_volume_vals(fg_probs, points, w, h, R, t, res, voxel, fgv, NULL);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < w * h; ++i) {
        const float f = assoc_out[i];
        const int invalid = (f == 0.0f);
        float v = fabsf(f);
        v = v * k1;
        v = expf(v);
        v = v * k2;
        if (fg_probs) v = v * fgv[i];
        v = v * alpha;
        v = v + k3;
        if (invalid) v = 0.0f;
        assoc_out[i] = v;
        if (assoc_mask_out) assoc_mask_out[i] = invalid ? 255 : 0;
    }
}

/* ------------------------------------------------------------------------
 * A.8 normaliser: src/core/EMFusion.cpp:653-665.  N = bg, then += each
 * object in ascending-id order (caller passes them in that order); every
 * image is divided by N with x/0 -> 0 (cv::cuda::divide semantics,
 * SURVEY.md section 8c (i); div0_is_zero = 0 switches to IEEE x/0).
 * imgs: n_vol pointers, imgs[0] = background.
 * ---------------------------------------------------------------------- */
EMFO_API void emfo_normalise(float** imgs, int n_vol, int n_pix, int div0_is_zero, float* norm_out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_pix; ++i) {
        float n = imgs[0][i];
        for (int v = 1; v < n_vol; ++v) n = n + imgs[v][i];
        if (norm_out) norm_out[i] = n;
        for (int v = 0; v < n_vol; ++v)
            imgs[v][i] = (n != 0.0f || !div0_is_zero) ? imgs[v][i] / n : 0.0f;
    }
}

/* ------------------------------------------------------------------------
 * A.6 composite: src/core/EMFusion.cpp:760-794.  Objects in list order;
 * an object takes a pixel iff objMask && (ray <= 0 || objRay < ray); then the
 * background wins the *label* where bgMask && ray - bgRay > 0.05 (the
 * composited raylength keeps the object's value -- reference behaviour);
 * vertices/normals come from the background where seg == 0.
 * The reference leaves diffRaylengths stale where bgMask == 0 (:773); this
 * restatement defines those pixels as "background does not win".
 * vis_count[k] = #pixels with seg == ids[k] inside the frame shrunk by
 * `boundary` on every side (:778-791).
 * ---------------------------------------------------------------------- */
EMFO_API void emfo_composite(int n_obj, const int* ids,
                             float** obj_ray, float** obj_vert, float** obj_norm, uint8_t** obj_mask,
                             const float* bg_ray, const float* bg_vert, const float* bg_norm, const uint8_t* bg_mask,
                             int w, int h, int boundary,
                             float* ray, float* vert, float* norm, uint8_t* seg, int64_t* vis_count) {
    for (int k = 0; k < n_obj; ++k) vis_count[k] = 0;
    for (int i = 0; i < w * h; ++i) {
        float r = 0.0f;
        int s = 0, win = -1;
        for (int k = 0; k < n_obj; ++k) {
            if (obj_mask[k][i] && (r <= 0.0f || obj_ray[k][i] < r)) {
                r = obj_ray[k][i];
                s = ids[k] > 255 ? 255 : ids[k]; /* modelSegmentation is CV_8U: setTo saturates */
                win = k;
            }
        }
        if (bg_mask[i] && (r - bg_ray[i] > 0.05f)) s = 0;
        ray[i] = r;
        seg[i] = (uint8_t)s;
        const float* vs = (s == 0) ? bg_vert : (obj_vert[win]);
        const float* ns = (s == 0) ? bg_norm : (obj_norm[win]);
        for (int c = 0; c < 3; ++c) {
            vert[3 * i + c] = vs[3 * i + c];
            norm[3 * i + c] = ns[3 * i + c];
        }
        const int x = i % w, y = i / w;
        if (s != 0 && x >= boundary && x < w - boundary && y >= boundary && y < h - boundary)
            for (int k = 0; k < n_obj; ++k)
                if (ids[k] == s) ++vis_count[k];
    }
}

/* ------------------------------------------------------------------------
 * fg/bg counts: src/core/cuda/ObjTSDF.cu:29-80 (kernel_updateFgBgProbs);
 * T_OC = cam_pose^-1 * pose (src/core/ObjTSDF.cpp:172).  fgbg is float2/voxel.
 * ---------------------------------------------------------------------- */
EMFO_API void emfo_update_fgbg(const uint8_t* mask, const uint8_t* occluded, int w, int h,
                               const float* tsdf, const float* weights, float* fgbg,
                               const float* R, const float* t, const float* K,
                               const int* res, float voxel) {
    const int rx = res[0], ry = res[1], rz = res[2];
#pragma omp parallel for schedule(static)
    for (int yz = 0; yz < ry * rz; ++yz) {
        const int y = yz % ry, z = yz / ry;
        const float oy = ((float)y - (float)(ry - 1) * 0.5f) * voxel;
        const float oz = ((float)z - (float)(rz - 1) * 0.5f) * voxel;
        for (int x = 0; x < rx; ++x) {
            const size_t i = (size_t)yz * rx + x;
            if (fabsf(tsdf[i]) >= 1.0f || weights[i] == 0.0f) continue;
            const float ox = ((float)x - (float)(rx - 1) * 0.5f) * voxel;
            const float pcx = t[0] + dot_yxz(R + 0, ox, oy, oz);
            const float pcy = t[1] + dot_yxz(R + 3, ox, oy, oz);
            const float pcz = t[2] + dot_yxz(R + 6, ox, oy, oz);
            if (pcz <= 0.0f) continue;
            const float qx = dot_yxz(K + 0, pcx, pcy, pcz);
            const float qy = dot_yxz(K + 3, pcx, pcy, pcz);
            const float qz = dot_yxz(K + 6, pcx, pcy, pcz);
            const int px = f2i_rn(qx / qz), py = f2i_rn(qy / qz);
            if (px < 0 || px >= w || py < 0 || py >= h) continue;
            const size_t p = (size_t)py * w + px;
            if (!occluded[p]) {
                const int m = mask[p] ? 1 : 0;
                fgbg[2 * i + 0] = fgbg[2 * i + 0] + (float)m;
                fgbg[2 * i + 1] = fgbg[2 * i + 1] + (float)(1 - m);
            }
        }
    }
}

/* fg probability: src/core/ObjTSDF.cpp:218-226 (computeFgProbs): fg/(fg+bg),
 * 0 where the sum is 0 (guarded divide) or NaN; fgVolMask = fgProb > 0.5. */
EMFO_API void emfo_compute_fg_probs(const float* fgbg, int64_t n_vox, float* fg_probs, uint8_t* fg_vol_mask) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_vox; ++i) {
        const float s = fgbg[2 * i] + fgbg[2 * i + 1];
        float p = (s != 0.0f) ? fgbg[2 * i] / s : 0.0f;
        if (p != p) p = 0.0f;
        fg_probs[i] = p;
        fg_vol_mask[i] = (p > 0.5f) ? 255 : 0;
    }
}

/* raycastWeights = tsdfWeights where fgVolMask else 0: src/core/ObjTSDF.cpp:209-210 */
EMFO_API void emfo_raycast_weights(const float* weights, const uint8_t* fg_vol_mask, int64_t n_vox, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_vox; ++i) out[i] = fg_vol_mask[i] ? weights[i] : 0.0f;
}

/* ------------------------------------------------------------------------
 * Tracker: the device part of one Levenberg-Marquardt iteration of emf::TSDF,
 * op by op as the reference chains it (src/core/TSDF.cpp:194-265, 375-394):
 *   grads        kernel_computePoseGradients  src/core/cuda/TSDF.cu:603-638 (after setTo(0); "xyz" contraction,
 *                gradient = trilinear of the float3 tsdfGrads volume / voxelSize, [p]x g with the build's contractions)
 *   tsdfVals     getVolumeVals(tsdfVol)       src/core/cuda/TSDF.cu:662-688
 *   intWeights   getVolumeVals(tsdfWeights); min(., maxTSDFWeight); normalize(NORM_INF): x * (float)(1/max|x|)
 *   trackWeights min(huberThresh / |tsdfVals|, 1) with cv::cuda::divide's x/0 = 0
 *   intWeights   = trackWeights * intWeights * associationWeights   (two multiplies)
 *   As, bs       kernel_computeAb :729-751, multSingletonCol by intWeights :821-853
 *   A, b         column sums (cv::cuda::reduce; here in double -- the reference's float order is unspecified)
 *   error        sum(tsdfVals^2 * intWeights) (cv::cuda::sum accumulates in double)
 * Outputs: grads6 (n x 6), tsdf_vals, int_weights, track_weights (n each), A[36], b[6], err, wmax -- doubles for the sums.
 * ---------------------------------------------------------------------- */
EMFO_API void emfo_track_linearise(const float* tsdf, const float* grads_vol, const float* weights, const float* points,
                                   const float* assoc, int w, int h, const float* R, const float* t, const int* res,
                                   float voxel, float huber, float maxw, float* grads6, float* tsdf_vals,
                                   float* int_weights, float* track_weights, double* A, double* b, double* err,
                                   double* wmax_out) {
    const int rx = res[0], ry = res[1], rz = res[2];
    const float frx = (float)rx, fry = (float)ry, frz = (float)rz;
    const int n = w * h;
    float wmax = 0.0f;
    for (int i = 0; i < n; ++i) {
        float* g = grads6 + 6 * (size_t)i;
        for (int k = 0; k < 6; ++k) g[k] = 0.0f;
        tsdf_vals[i] = 0.0f; int_weights[i] = 0.0f;
        const float* p = points + 3 * (size_t)i;
        if (!(p[2] <= 0.0f)) {
            const float qx = t[0] + dot_xyz(R + 0, p[0], p[1], p[2]);
            const float qy = t[1] + dot_xyz(R + 3, p[0], p[1], p[2]);
            const float qz = t[2] + dot_xyz(R + 6, p[0], p[1], p[2]);
            const float vx = (float)(rx - 1) * 0.5f + qx / voxel;
            const float vy = (float)(ry - 1) * 0.5f + qy / voxel;
            const float vz = (float)(rz - 1) * 0.5f + qz / voxel;
            if (!out_of(vx, vy, vz, 1.0f, frx, fry, frz)) {
                tsdf_vals[i] = trilinear(tsdf, rx, ry, vx, vy, vz);
                int_weights[i] = trilinear(weights, rx, ry, vx, vy, vz);
            }
            if (!out_of(vx, vy, vz, 2.0f, frx, fry, frz)) {
                float gi[3];
                trilinear3(grads_vol, rx, ry, vx, vy, vz, gi);
                g[0] = gi[0] / voxel; g[1] = gi[1] / voxel; g[2] = gi[2] / voxel;
                g[3] = FMA(qy, g[2], -(qz * g[1]));
                g[4] = FMA(-qx, g[2], qz * g[0]);      /* contraction as in the SASS of the reference build */
                g[5] = FMA(qx, g[1], -(qy * g[0]));
            }
        }
        const float af = fabsf(tsdf_vals[i]);
        float tw = af != 0.0f ? huber / af : 0.0f;
        track_weights[i] = tw < 1.0f ? tw : 1.0f;
        int_weights[i] = int_weights[i] < maxw ? int_weights[i] : maxw;
        if (fabsf(int_weights[i]) > wmax) wmax = fabsf(int_weights[i]);
    }
    const double nrm = (double)wmax;
    const float scale = nrm > 2.220446049250313e-16 ? (float)(1.0 / nrm) : 0.0f;
    for (int k = 0; k < 36; ++k) A[k] = 0.0;
    for (int k = 0; k < 6; ++k) b[k] = 0.0;
    double e = 0.0;
    for (int i = 0; i < n; ++i) {
        float iw = int_weights[i] * scale;
        iw = track_weights[i] * iw;
        iw = iw * assoc[i];
        int_weights[i] = iw;
        const float* g = grads6 + 6 * (size_t)i;
        for (int a = 0; a < 6; ++a) {
            for (int c = 0; c < 6; ++c) {
                const float as = g[a] * g[c];
                A[a * 6 + c] += (double)(as * iw);
            }
            const float bs = tsdf_vals[i] * g[a];
            b[a] += (double)(bs * iw);
        }
        const float sq = tsdf_vals[i] * tsdf_vals[i];
        e += (double)(sq * iw);
    }
    *err = e;
    if (wmax_out) *wmax_out = nrm;
}

/* ------------------------------------------------------------------------
 * Depth pre-filter: emf::EMFusion::preprocessDepth, src/core/EMFusion.cpp:294-305.
 * cv::cuda::bilateralFilter is OpenCV-CUDA code (un-vendored, version unpinned): PARITY UNPINNED -- its published kernel
 * (opencv_contrib cudaimgproc bilateral_filter.cu) is restated: disc of radius ksize/2, float accumulation in row-major
 * window order, exp(space2 * (-0.5/ss^2) + diff^2 * (-0.5/sc^2)), BORDER_REFLECT_101, sum1 / sum2; then NaN -> 0 and
 * raw == 0 -> 0.
 * ---------------------------------------------------------------------- */
static int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}
EMFO_API void emfo_preprocess_depth(const float* raw, int w, int h, int kernel_size, float sigma_depth, float sigma_spatial,
                                    float* out) {
    if (sigma_depth <= 0.0f) sigma_depth = 1.0f;
    if (sigma_spatial <= 0.0f) sigma_spatial = 1.0f;
    int r = kernel_size <= 0 ? (int)lrintf(sigma_spatial * 1.5f) : kernel_size / 2;
    if (r < 1) r = 1;
    const float ss = -0.5f / (sigma_spatial * sigma_spatial), sc = -0.5f / (sigma_depth * sigma_depth);
    const float r2 = (float)(r * r);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float centre = raw[(size_t)y * w + x];
            float sum1 = 0.0f, sum2 = 0.0f;
            for (int dy = -r; dy <= r; ++dy)
                for (int dx = -r; dx <= r; ++dx) {
                    const float space2 = (float)(dx * dx + dy * dy);
                    if (space2 > r2) continue;
                    const float v = raw[(size_t)reflect101(y + dy, h) * w + reflect101(x + dx, w)];
                    const float d = v - centre;
                    const float wgt = expf(FMA(space2, ss, (d * d) * sc));
                    sum1 = FMA(wgt, v, sum1);
                    sum2 = sum2 + wgt;
                }
            float res = sum1 / sum2;
            if (res != res) res = 0.0f;
            if (centre == 0.0f) res = 0.0f;
            out[(size_t)y * w + x] = res;
        }
}

EMFO_API int emfo_uses_fma(void) {
#ifdef EMFO_NOFMA
    return 0;
#else
    return 1;
#endif
}
