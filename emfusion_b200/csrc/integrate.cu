// integrate.cu -- depth -> TSDF integration for one or many volumes in ONE launch.
//
// Replaces emf::cuda::TSDF::updateTSDF (reference src/core/cuda/TSDF.cu:327-427),
// called once per volume on its own stream by emf::EMFusion::integrateDepth
// (src/core/EMFusion.cpp:865-889).  Every volume of the frame is a row in a
// descriptor table (kernel parameter, constant bank).
//
// k_integrate_rows (the hot kernel; x-resolution a multiple of 8, 16-byte aligned arrays):
//  * persistent grid (a multiple of the SM count); a warp owns whole x-rows (y,z fixed) of a
//    volume, rows are dealt round-robin so neighbouring rows -- which project to neighbouring
//    image rows -- are in flight together;
//  * per row, the x-interval whose voxels can project into the image is solved in closed form
//    (four half-planes of the pinhole frustum, padded by kMarginPx pixels and one voxel), so the
//    ~65 % of a room-sized grid that lies outside the frustum costs one warp-uniform test per row
//    instead of a projection per voxel.  The reference never touches those voxels either;
//  * inside the interval a thread owns 4 consecutive voxels (one float4 of tsdf, one of weights);
//  * the result is bit-identical to the reference build: all arithmetic that reaches memory is the
//    canonical sequence of emf_math.cuh.  The expensive IEEE divisions/square roots are only
//    evaluated where their exact value matters: the pixel index comes from an approximate quotient
//    whenever that quotient is provably on the same side of the rounding boundary, and voxels that
//    an approximate signed distance places safely outside the truncation band take the free-space
//    (value +1, weight +1) or occluded branch without the exact distance;
//  * optionally maintains three bitmaps with one bit per 4-voxel x-segment (all +1 / all 0 /
//    all -1), written with warp ballots, from which bricks.cu derives the brick map the
//    raycast uses to skip march samples whose outcome is known (raycast.cu).
//
// k_integrate_simple: one thread per voxel, any resolution/alignment (ragged volumes).
#include "common.cuh"
#include <stdlib.h>

namespace emfb {

#ifndef EMF_INT_PYR2
#define EMF_INT_PYR2 1
#endif
constexpr int kPyrLevels = 6;   // depth pyramid levels 1..6 (tiles of 2..64 pixels)

struct IntVol {
    float* tsdf;
    float* weights;
    uint32_t* const_bits; // nullable: three bitmaps (all +1 / all 0 / all -1), one bit per 4-voxel x-segment
    int wpr;              // 32-bit words per row in a bitmap
    size_t map_words;     // words per bitmap
    const float* assoc;   // this volume's association image
    size_t assoc_pitch;
    float R[9];           // T_OC
    float t[3];
    int rx, ry, rz;
    float voxel, trunc;
    int first_item;       // prefix sum of work items (rows for k_integrate_rows, CTAs for k_integrate_simple)
    int gate;             // index into IntParams::gate_counts, or -1: integrate unconditionally
    int first_brick;      // prefix sum of 32 x 4 x 4 bricks (k_brick_classify / k_integrate_bricks)
    int nbx, nby;         // bricks per row of bricks / rows of bricks per slice
};

struct IntParams {
    IntVol v[EMF_MAX_VOLUMES];
    int n_vol;
    int total_items;
    const float* depth;
    size_t depth_pitch;
    int w, h;
    float K[9];
    float max_weight;
    const int32_t* gate_counts;   // nullable: device-side visibility counters (raycast composite)
    int gate_thresh;              // a gated volume is integrated iff gate_counts[gate] > gate_thresh
    unsigned long long* stats;    // nullable: [0] updated [1] marked -1 [2] occluded-seen [3] check-only [4] skipped-in-interval
    float g_rel, g_abs;           // |(|pc| / lambda(pixel)) - pc.z| <= g_rel * pc.z + g_abs (classification guard)
    const float2* pyr[kPyrLevels + 1];   // depth (min, max) pyramid, level l = tiles of 2^l pixels (k_integrate_seg); [0] unused
    int pyr_w[kPyrLevels + 1];           // tiles per row of level l
    int* work_counter;                   // k_integrate_seg: next batch of rows (zeroed by k_depth_pyramid)
    const float* inv_lambda;             // k_integrate_seg: 1 / |((x-cx)/fx, (y-cy)/fy, 1)| per pixel, exact (k_depth_pyramid)
    // brick level (k_brick_classify -> k_integrate_bricks)
    int total_bricks;
    uint2* list_mixed;                   // {volume << 22 | brick, classes of its four 8 x 4 x 4 sub-bricks}: some part needs the per-segment treatment
    uint2* list_whole;                   // ... every part decided (free / occluded / behind the camera / outside the image)
    int* brick_counters;                 // [0] mixed [1] whole [2] next item (zeroed by k_depth_pyramid)
};

constexpr int kIntThreads = 256;
constexpr int kSegThreads = 256;
constexpr int kRowBatch = 8;     // rows whose set-up a warp computes at once (one per lane) in k_integrate_seg
constexpr int kSimpleThreads = 128;
constexpr float kMarginPx = 3.0f;    // frustum half-planes are pushed out by this many pixels
constexpr float kMinDepthCull = 0.02f;   // rows that come closer than this to the camera plane are not culled

__device__ __forceinline__ int value_code(float v) {
    return v == 1.0f ? kAllOne : (v == -1.0f ? kAllMinusOne : (v == 0.0f ? kAllZero : kMixed));
}

// __float2int_rn(fdiv(q, qz)) without the IEEE division whenever an approximate quotient is
// provably on the same side of every rounding boundary (error of q * rcp.approx(qz) is below
// 2^-22 |q/qz|; we keep 2^-20 |u| + 2^-20 away from the half-integers).
__device__ __forceinline__ int round_quotient(float q, float qz, float rz) {
    const float u = q * rz;
    const float r = rintf(u);
    const float slack = fmaf(fabsf(u), 9.5367431640625e-7f, 9.5367431640625e-7f);
    if (fabsf(u - r) + slack < 0.5f && fabsf(u) < 1.0e6f) return (int)r;
    return __float2int_rn(fdiv(q, qz));
}

// voxel classes of one lane's 4-voxel segment, one bit per voxel in each mask
struct SegClass { int check, free_, occ, exact, skip; };

template <bool PINHOLE, bool TABLE, bool STATS>
__global__ void __launch_bounds__(kIntThreads, 4) k_integrate_rows(const __grid_constant__ IntParams P) {
    extern __shared__ float s_tab[];   // [0, w): (x - cx) / fx ; [w, w + h): (y - cy) / fy   (exact IEEE quotients)
    if (TABLE) {
        for (int i = threadIdx.x; i < P.w + P.h; i += kIntThreads)
            s_tab[i] = i < P.w ? fdiv(fsub((float)i, P.K[2]), P.K[0]) : fdiv(fsub((float)(i - P.w), P.K[5]), P.K[4]);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = kIntThreads / 32;
    const int gwarp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * warps_per_cta;
    unsigned long long st[5] = {0, 0, 0, 0, 0};

    int vi = 0;
    for (int item = gwarp; item < P.total_items; item += n_warps) {
        while (vi + 1 < P.n_vol && P.v[vi + 1].first_item <= item) ++vi;   // rows are dealt in ascending order
        const IntVol& V = P.v[vi];
        if (V.gate >= 0 && !(__ldg(P.gate_counts + V.gate) > P.gate_thresh)) continue;   // not visible: not integrated
        const int rx = V.rx, ry = V.ry;
        const int row = item - V.first_item;          // z * Ry + y
        const int z = row / ry;
        const int y = row - z * ry;
        const float s = V.voxel;
        // (i - (R-1)/2.f) * voxelSize ; (R-1)*0.5 is exact
        const float hx = fmul((float)(rx - 1), 0.5f);
        const float cy = fmul(fsub((float)y, fmul((float)(ry - 1), 0.5f)), s);
        const float cz = fmul(fsub((float)z, fmul((float)(V.rz - 1), 0.5f)), s);
        const float my0 = fmul(V.R[1], cy), my1 = fmul(V.R[4], cy), my2 = fmul(V.R[7], cy);

        // ---- conservative x-interval of voxels that can project into the image (plain float math;
        //      the padding absorbs its rounding).  pc(x) = A + x * B.
        int xa = 0, xb = rx - 1;
        {
            const float c0 = -hx * s;
            const float ax = V.t[0] + (V.R[0] * c0 + my0 + V.R[2] * cz), bx = V.R[0] * s;
            const float ay = V.t[1] + (V.R[3] * c0 + my1 + V.R[5] * cz), by = V.R[3] * s;
            const float az = V.t[2] + (V.R[6] * c0 + my2 + V.R[8] * cz), bz = V.R[6] * s;
            float qxa, qxb, qya, qyb, qza, qzb;
            if (PINHOLE) {
                qxa = P.K[0] * ax + P.K[2] * az; qxb = P.K[0] * bx + P.K[2] * bz;
                qya = P.K[4] * ay + P.K[5] * az; qyb = P.K[4] * by + P.K[5] * bz;
                qza = az; qzb = bz;
            } else {
                qxa = P.K[0] * ax + P.K[1] * ay + P.K[2] * az; qxb = P.K[0] * bx + P.K[1] * by + P.K[2] * bz;
                qya = P.K[3] * ax + P.K[4] * ay + P.K[5] * az; qyb = P.K[3] * bx + P.K[4] * by + P.K[5] * bz;
                qza = P.K[6] * ax + P.K[7] * ay + P.K[8] * az; qzb = P.K[6] * bx + P.K[7] * by + P.K[8] * bz;
            }
            const float xe = (float)(rx - 1);
            const float zmin = fminf(fminf(az, az + bz * xe), fminf(qza, qza + qzb * xe));
            if (zmin > kMinDepthCull) {   // whole row safely in front of the camera: cull by the four image edges
                float lo = 0.0f, hi = xe;
                const float m0 = 0.5f + kMarginPx, mw = (float)P.w - 0.5f + kMarginPx, mh = (float)P.h - 0.5f + kMarginPx;
                // alpha + beta * x >= 0 for: left, right, top, bottom
                const float al[4] = {qxa + m0 * qza, mw * qza - qxa, qya + m0 * qza, mh * qza - qya};
                const float be[4] = {qxb + m0 * qzb, mw * qzb - qxb, qyb + m0 * qzb, mh * qzb - qyb};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (be[k] > 0.0f) lo = fmaxf(lo, -al[k] / be[k]);
                    else if (be[k] < 0.0f) hi = fminf(hi, -al[k] / be[k]);
                    else if (al[k] < 0.0f) hi = -1.0f;
                }
                if (!(lo <= hi)) continue;
                xa = max(0, (int)floorf(lo) - 1);
                xb = min(rx - 1, (int)ceilf(hi) + 1);
                if (xa > xb) continue;
            }
        }
        xa &= ~127;                     // whole 128-voxel chunks: one bitmap word per warp iteration
        const int64_t row_off = (int64_t)row * rx;
        const ConstDiv div_trunc(V.trunc);
        const float ntrunc = -V.trunc;

        for (int xbase = xa; xbase <= xb; xbase += 128) {   // warp-uniform trip count (ballots below)
            const int x0 = xbase + 4 * lane;
            const bool active = x0 <= xb && x0 < rx;
            int cls_skip = 0, cls_check = 0, cls_free = 0, cls_occ = 0, cls_exact = 0;   // bit j = voxel j
            int pix[4];        // py << 16 | px of the voxels that are looked at
            float dep[4];
            // ---- phase 1: projection (canonical), depth look-up, classification by an approximate signed distance
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                dep[j] = 0.f; pix[j] = 0;
                if (!active) continue;
                const float cx = fmul(fsub((float)(x0 + j), hx), s);
                const float pcx = fadd(V.t[0], ffma(V.R[2], cz, ffma(V.R[0], cx, my0)));
                const float pcy = fadd(V.t[1], ffma(V.R[5], cz, ffma(V.R[3], cx, my1)));
                const float pcz = fadd(V.t[2], ffma(V.R[8], cz, ffma(V.R[6], cx, my2)));
                if (!(pcz > 0.0f)) { cls_check |= 1 << j; continue; }
                float qx, qy, qz;
                if (PINHOLE) {
                    qx = ffma(P.K[2], pcz, fmul(P.K[0], pcx));
                    qy = ffma(P.K[5], pcz, fmul(P.K[4], pcy));
                    qz = pcz;
                } else {
                    qx = dot_yxz(P.K[0], P.K[1], P.K[2], pcx, pcy, pcz);
                    qy = dot_yxz(P.K[3], P.K[4], P.K[5], pcx, pcy, pcz);
                    qz = dot_yxz(P.K[6], P.K[7], P.K[8], pcx, pcy, pcz);
                }
                const float rz = rcp_approx(qz);
                const int px = round_quotient(qx, qz, rz);
                const int py = round_quotient(qy, qz, rz);
                if ((unsigned)px >= (unsigned)P.w || (unsigned)py >= (unsigned)P.h) { cls_skip |= 1 << j; continue; }
                pix[j] = (py << 16) | px;
                const float d = __ldg((const float*)((const char*)P.depth + (size_t)py * P.depth_pitch) + px);
                dep[j] = d;
                if (!(d > 0.0f)) { cls_check |= 1 << j; continue; }
                // |pc| / lambda(pixel) lies within g_rel * pcz of pcz (the pixel ray and the voxel ray differ by at most
                // half a pixel); outside the truncation band by more than that, the exact distance is not needed
                const float sdf_a = d - pcz;
                const float lim = fmaf(pcz, P.g_rel, V.trunc + P.g_abs);
                if (sdf_a > lim) cls_free |= 1 << j;
                else if (sdf_a < -lim) cls_occ |= 1 << j;
                else cls_exact |= 1 << j;
            }
            const int touched = cls_check | cls_free | cls_occ | cls_exact;
            int known = 0;                 // voxels whose final tsdf value this thread knows
            float tv[4] = {0.f, 0.f, 0.f, 0.f};
            if (touched) {
                float* wp = V.weights + row_off + x0;
                float* tp = V.tsdf + row_off + x0;
                float w[4];
                {
                    const float4 w4 = *reinterpret_cast<const float4*>(wp);
                    w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
                }
                if (cls_free | cls_exact) {
                    const float4 t4 = *reinterpret_cast<const float4*>(tp);
                    tv[0] = t4.x; tv[1] = t4.y; tv[2] = t4.z; tv[3] = t4.w;
                    known = 0xF;
                }
                int wrote_t = 0, wrote_w = 0;
                // ---- phase 2a: the classes that need no distance
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int bit = 1 << j;
                    if (cls_check & bit) {
                        // behind the camera or no depth: un-mark never-seen voxels (TSDF.cu:349-353, 369-374)
                        if (w[j] == 0.0f) { tv[j] = 0.0f; wrote_t |= bit; known |= bit; }
                        if (STATS) ++st[3];
                    } else if (cls_free & bit) {
                        // sdf >= trunc: value +1 with weight 1 (free space is never association-weighted)
                        const float ws = fadd(w[j], 1.0f);
                        if (ws > 0.0f) {
                            const float num = ffma(w[j], tv[j], 1.0f);
                            tv[j] = (num == ws && ws <= 3.0e38f) ? 1.0f : fdiv(num, ws);
                            w[j] = fminf(ws, P.max_weight);
                            wrote_t |= bit; wrote_w |= bit;
                            if (STATS) ++st[0];
                        }
                    } else if (cls_occ & bit) {
                        // far behind the surface
                        if (w[j] == 0.0f) { tv[j] = -1.0f; wrote_t |= bit; known |= bit; if (STATS) ++st[1]; }
                        else if (STATS) ++st[2];
                    }
                }
                // ---- phase 2b: voxels near the truncation band: exact signed distance (canonical sequence)
                if (cls_exact) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int bit = 1 << j;
                        if (!(cls_exact & bit)) continue;
                        const int px = pix[j] & 0xffff, py = pix[j] >> 16;
                        const float cx = fmul(fsub((float)(x0 + j), hx), s);
                        const float pcx = fadd(V.t[0], ffma(V.R[2], cz, ffma(V.R[0], cx, my0)));
                        const float pcy = fadd(V.t[1], ffma(V.R[5], cz, ffma(V.R[3], cx, my1)));
                        const float pcz = fadd(V.t[2], ffma(V.R[8], cz, ffma(V.R[6], cx, my2)));
                        float lx, ly;
                        if (TABLE) { lx = s_tab[px]; ly = s_tab[P.w + py]; }
                        else { lx = fdiv(fsub((float)px, P.K[2]), P.K[0]); ly = fdiv(fsub((float)py, P.K[5]), P.K[4]); }
                        const float lambda = fsqrt(fadd(ffma(lx, lx, fmul(ly, ly)), 1.0f));
                        const float inv_lambda = frcp(lambda);
                        const float nrm = norm3(pcx, pcy, pcz);
                        const float sdf = ffma(-nrm, inv_lambda, dep[j]);   // depth - (1/lambda)*|pc| as one FFMA (reference SASS)
                        if (sdf >= ntrunc) {
                            const float q = div_trunc(sdf);
                            const float val = copysignf(fminf(1.0f, fabsf(q)), sdf);
                            float a = 1.0f;
                            if (sdf < V.trunc)
                                a = __ldg((const float*)((const char*)V.assoc + (size_t)py * V.assoc_pitch) + px);
                            const float ws = fadd(w[j], a);
                            if (ws > 0.0f) {
                                tv[j] = fdiv(ffma(w[j], tv[j], fmul(val, a)), ws);
                                w[j] = fminf(ws, P.max_weight);
                                wrote_t |= bit; wrote_w |= bit;
                                if (STATS) ++st[0];
                            }
                        } else if (w[j] == 0.0f) {
                            tv[j] = -1.0f; wrote_t |= bit; if (STATS) ++st[1];
                        } else if (STATS) ++st[2];
                    }
                }
                if (wrote_w) *reinterpret_cast<float4*>(wp) = make_float4(w[0], w[1], w[2], w[3]);
                if (wrote_t) {
                    if (known == 0xF) {
                        *reinterpret_cast<float4*>(tp) = make_float4(tv[0], tv[1], tv[2], tv[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (wrote_t & (1 << j)) tp[j] = tv[j];
                    }
                }
            }
            if (STATS) st[4] += __popc(cls_skip);
            // ---- constant-segment bitmaps: one bit per 4-voxel segment (this lane) in each of three maps
            //      (all +1 / all 0 / all -1).  A word covers this warp's 128-voxel chunk, so it has one owner.
            if (V.const_bits && __any_sync(0xffffffffu, known != 0)) {
                int code = -1;    // common code of the values this lane knows; kMixed if they differ or are not constants
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (known & (1 << j)) {
                        const int c = value_code(tv[j]);
                        code = (code == -1 || code == c) ? c : kMixed;
                    }
                const bool full = (known == 0xF);
                const size_t word = (size_t)row * V.wpr + (xbase >> 7);
#pragma unroll
                for (int m = 1; m <= 3; ++m) {
                    const uint32_t set = __ballot_sync(0xffffffffu, full && code == m);
                    // a partially known segment keeps its old bit only if what was seen agrees with it
                    const uint32_t keep = __ballot_sync(0xffffffffu, !full && (code == -1 || code == m));
                    if (lane == m - 1) {
                        uint32_t* wp32 = V.const_bits + (size_t)(m - 1) * V.map_words + word;
                        *wp32 = keep ? (set | (*wp32 & keep)) : set;
                    }
                }
            }
        }
    }
    if (STATS && P.stats) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            unsigned long long v = st[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && v) atomicAdd(P.stats + k, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// depth (min, max) pyramid: level l holds, per tile of 2^l x 2^l pixels, the smallest and the largest depth.
// A pixel without a measurement (d <= 0 or NaN) poisons its tiles with (0, +inf), so that nothing whose footprint
// contains it can be classified wholesale.  One CTA reduces one 64 x 64 pixel tile through all levels.
// ---------------------------------------------------------------------------------------------
struct PyrParams {
    const float* depth; size_t pitch; int w, h;
    float2* lvl[kPyrLevels + 1];
    int lw[kPyrLevels + 1], lh[kPyrLevels + 1];
    int* work_counter;
    int* brick_counters;  // nullable: 3 ints, zeroed here
    float* inv_lambda;    // W x H, continuous
    float fx, fy, cx, cy;
};

__global__ void __launch_bounds__(256) k_depth_pyramid(const __grid_constant__ PyrParams P) {
    __shared__ float2 s_a[32][33];
    const int tx0 = blockIdx.x * 64, ty0 = blockIdx.y * 64;
    const int t = threadIdx.x;
    if (blockIdx.x == 0 && blockIdx.y == 0 && t == 0) {
        *P.work_counter = 0;
        if (P.brick_counters) { P.brick_counters[0] = 0; P.brick_counters[1] = 0; P.brick_counters[2] = 0; }
    }
    // 1 / lambda of every pixel of this tile, by the canonical sequence of the reference kernel (TSDF.cu:376-380):
    // lambda = sqrt(((x-cx)/fx)^2 + ((y-cy)/fy)^2 + 1), then its IEEE reciprocal
    for (int i = t; i < 64 * 64; i += 256) {
        const int x = tx0 + (i & 63), y = ty0 + (i >> 6);
        if (x < P.w && y < P.h) {
            const float lx = fdiv(fsub((float)x, P.cx), P.fx), ly = fdiv(fsub((float)y, P.cy), P.fy);
            P.inv_lambda[(size_t)y * P.w + x] = frcp(fsqrt(fadd(ffma(lx, lx, fmul(ly, ly)), 1.0f)));
        }
    }
    // level 1 straight from the image: thread -> 4 of the 32 x 32 level-1 tiles
    for (int i = t; i < 32 * 32; i += 256) {
        const int lx = i & 31, ly = i >> 5;
        float mn = INFINITY, mx = -INFINITY;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int x = tx0 + 2 * lx + dx, y = ty0 + 2 * ly + dy;
                if (x < P.w && y < P.h) {
                    const float d = __ldg((const float*)((const char*)P.depth + (size_t)y * P.pitch) + x);
                    if (d > 0.0f) { mn = fminf(mn, d); mx = fmaxf(mx, d); }
                    else { mn = 0.0f; mx = INFINITY; }
                }
            }
        s_a[ly][lx] = make_float2(mn, mx);
        const int gx = (tx0 >> 1) + lx, gy = (ty0 >> 1) + ly;
        if (gx < P.lw[1] && gy < P.lh[1]) P.lvl[1][(size_t)gy * P.lw[1] + gx] = make_float2(mn, mx);
    }
    __syncthreads();
    // levels 2..6 in shared memory (in place: level l occupies the top-left (64 >> l)^2 corner)
    for (int l = 2; l <= kPyrLevels; ++l) {
        const int n = 64 >> l;
        float2 v = make_float2(INFINITY, -INFINITY);
        const int lx = t % n, ly = t / n;
        const bool on = t < n * n;
        if (on) {
            const float2 a = s_a[2 * ly][2 * lx], b = s_a[2 * ly][2 * lx + 1], c = s_a[2 * ly + 1][2 * lx], d = s_a[2 * ly + 1][2 * lx + 1];
            v.x = fminf(fminf(a.x, b.x), fminf(c.x, d.x));
            v.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
        }
        __syncthreads();
        if (on) {
            s_a[ly][lx] = v;
            const int gx = (tx0 >> l) + lx, gy = (ty0 >> l) + ly;
            if (gx < P.lw[l] && gy < P.lh[l]) P.lvl[l][(size_t)gy * P.lw[l] + gx] = v;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// k_integrate_seg: k_integrate_rows with one more level of "decide cheaply, compute exactly only where needed".
//
//  * A lane's 4-voxel segment is first classified AS A WHOLE: its two end voxels are projected (approximately), the
//    depth pyramid gives the smallest and largest measurement over the padded pixel box of the segment (<= 3 x 3
//    tiles of the level whose tile size exceeds half the box), and if even the farthest voxel is in front of the
//    nearest measurement by more than the truncation distance plus a guard -- or the nearest voxel behind the
//    farthest measurement -- all four voxels take the free-space (value +1, weight +1) or occluded branch with
//    no per-voxel projection at all.  Pinhole cameras only (launcher).
//  * The segments that cannot be decided wholesale (near surfaces, depth edges, image border, no measurement) are
//    compacted across the warp: their 4 x n voxels are dealt one per lane, so that the exact per-voxel path -- the
//    canonical arithmetic of the reference, bit for bit -- runs on full warps instead of a few lanes.
//  * Result: bit-identical to k_integrate_rows / the reference kernel.
// ---------------------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(kSegThreads, 4) k_integrate_seg(const __grid_constant__ IntParams P) {
    __shared__ uint8_t s_src[kSegThreads / 32][32];
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warps_per_cta = kSegThreads / 32;
    const int gwarp = blockIdx.x * warps_per_cta + wid;
    const int n_warps = gridDim.x * warps_per_cta;
    unsigned long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // [5] voxels on the exact path [6] / [7] segments decided free / occluded
    const float fw = (float)P.w, fh = (float)P.h;

    // Rows are handed out in batches of kRowBatch consecutive rows from a global counter (zeroed by k_depth_pyramid):
    // dynamic, because the cost of a row ranges from nothing (outside the frustum) to four full chunks.  The per-row
    // set-up -- the row as a line in homogeneous pixel coordinates and the x-interval that can project into the image --
    // is computed for the rows of a batch at once, one row per lane; rows with an empty interval cost nothing more.
    (void)gwarp; (void)n_warps;
    for (;;) {
        int batch = 0;
        if (lane == 0) batch = atomicAdd(P.work_counter, 1);
        batch = __shfl_sync(kFull, batch, 0);
        const int item0 = batch * kRowBatch;
        if (item0 >= P.total_items) break;
        int l_vi = 0, l_row = 0, l_xa = 0, l_xb = -1;
        float l_q[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        {
            const int item = item0 + lane;
            if (lane < kRowBatch && item < P.total_items) {
                int lo = 0, hi = P.n_vol - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (P.v[mid].first_item <= item) lo = mid; else hi = mid - 1;
                }
                l_vi = lo;
                const IntVol& V = P.v[lo];
                const bool gated_out = V.gate >= 0 && !(__ldg(P.gate_counts + V.gate) > P.gate_thresh);   // not visible: not integrated
                if (!gated_out) {
                    const int rx = V.rx, ry = V.ry;
                    l_row = item - V.first_item;          // z * Ry + y
                    const int z = l_row / ry;
                    const int y = l_row - z * ry;
                    const float s = V.voxel;
                    const float hx = (float)(rx - 1) * 0.5f;
                    const float cy = ((float)y - (float)(ry - 1) * 0.5f) * s;
                    const float cz = ((float)z - (float)(V.rz - 1) * 0.5f) * s;
                    // q(x) = qa + x * qb (plain float math: only used for decisions that carry their own safety margins)
                    const float c0 = -hx * s;
                    const float ax = V.t[0] + (V.R[0] * c0 + V.R[1] * cy + V.R[2] * cz), bx = V.R[0] * s;
                    const float ay = V.t[1] + (V.R[3] * c0 + V.R[4] * cy + V.R[5] * cz), by = V.R[3] * s;
                    const float az = V.t[2] + (V.R[6] * c0 + V.R[7] * cy + V.R[8] * cz), bz = V.R[6] * s;
                    const float qxa = P.K[0] * ax + P.K[2] * az, qxb = P.K[0] * bx + P.K[2] * bz;
                    const float qya = P.K[4] * ay + P.K[5] * az, qyb = P.K[4] * by + P.K[5] * bz;
                    l_q[0] = qxa; l_q[1] = qxb; l_q[2] = qya; l_q[3] = qyb; l_q[4] = az; l_q[5] = bz;
                    l_xa = 0; l_xb = rx - 1;
                    const float xe = (float)(rx - 1);
                    const float zmin = fminf(az, az + bz * xe);
                    if (zmin > kMinDepthCull) {   // whole row safely in front of the camera: cull by the four image edges
                        float lo_x = 0.0f, hi_x = xe;
                        const float m0 = 0.5f + kMarginPx, mw = fw - 0.5f + kMarginPx, mh = fh - 0.5f + kMarginPx;
                        const float al[4] = {qxa + m0 * az, mw * az - qxa, qya + m0 * az, mh * az - qya};
                        const float be[4] = {qxb + m0 * bz, mw * bz - qxb, qyb + m0 * bz, mh * bz - qyb};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (be[k] > 0.0f) lo_x = fmaxf(lo_x, __fdividef(-al[k], be[k]) - 0.01f);
                            else if (be[k] < 0.0f) hi_x = fminf(hi_x, __fdividef(-al[k], be[k]) + 0.01f);
                            else if (al[k] < 0.0f) hi_x = -1.0f;
                        }
                        if (lo_x <= hi_x) {
                            l_xa = max(0, (int)floorf(lo_x) - 1);
                            l_xb = min(rx - 1, (int)ceilf(hi_x) + 1);
                        } else {
                            l_xb = -1;
                        }
                    }
                }
            }
        }
        unsigned todo = __ballot_sync(kFull, l_xa <= l_xb);
        while (todo) {
        const int src_lane = __ffs(todo) - 1;
        todo &= todo - 1;
        const int vi = __shfl_sync(kFull, l_vi, src_lane);
        const int row = __shfl_sync(kFull, l_row, src_lane);
        int xa = __shfl_sync(kFull, l_xa, src_lane);
        const int xb = __shfl_sync(kFull, l_xb, src_lane);
        const float qxa = __shfl_sync(kFull, l_q[0], src_lane), qxb = __shfl_sync(kFull, l_q[1], src_lane);
        const float qya = __shfl_sync(kFull, l_q[2], src_lane), qyb = __shfl_sync(kFull, l_q[3], src_lane);
        const float qza = __shfl_sync(kFull, l_q[4], src_lane), qzb = __shfl_sync(kFull, l_q[5], src_lane);
        const IntVol& V = P.v[vi];
        const int rx = V.rx, ry = V.ry;
        const int z = row / ry;
        const int y = row - z * ry;
        const float s = V.voxel;
        // (i - (R-1)/2.f) * voxelSize ; (R-1)*0.5 is exact -- canonical, for the exact per-voxel path
        const float hx = fmul((float)(rx - 1), 0.5f);
        const float cy = fmul(fsub((float)y, fmul((float)(ry - 1), 0.5f)), s);
        const float cz = fmul(fsub((float)z, fmul((float)(V.rz - 1), 0.5f)), s);
        const float my0 = fmul(V.R[1], cy), my1 = fmul(V.R[4], cy), my2 = fmul(V.R[7], cy);
        xa &= ~127;                     // whole 128-voxel chunks: one bitmap word per warp iteration
        const int64_t row_off = (int64_t)row * rx;
        const ConstDiv div_trunc(V.trunc);
        const float ntrunc = -V.trunc;
        const float band = V.trunc + P.g_abs;

        for (int xbase = xa; xbase <= xb; xbase += 128) {   // warp-uniform trip count (collectives below)
            const int x0 = xbase + 4 * lane;
            const bool active = x0 <= xb && x0 < rx;
            // ---- phase A: classify the segment as a whole.  0 skip, 1 free, 2 occluded, 3 per voxel
            int cls = 0;
            if (active) {
                cls = 3;
                const float xf0 = (float)x0, xf3 = (float)(x0 + 3);
                const float z0 = fmaf(xf0, qzb, qza), z3 = fmaf(xf3, qzb, qza);
                const float zlo = fminf(z0, z3), zhi = fmaxf(z0, z3);
                if (zlo > 0.05f) {
                    const float r0 = rcp_approx(z0), r3 = rcp_approx(z3);
                    const float u0 = fmaf(xf0, qxb, qxa) * r0, u3 = fmaf(xf3, qxb, qxa) * r3;
                    const float v0 = fmaf(xf0, qyb, qya) * r0, v3 = fmaf(xf3, qyb, qya) * r3;
                    // pixel box of the four voxels (the projection of a line segment is monotone in each coordinate),
                    // padded by one pixel for the rounding to the nearest pixel and the approximations above
                    const float ulo = fminf(u0, u3) - 1.0f, uhi = fmaxf(u0, u3) + 1.0f;
                    const float vlo = fminf(v0, v3) - 1.0f, vhi = fmaxf(v0, v3) + 1.0f;
                    if (uhi < -0.5f || ulo > fw - 0.5f || vhi < -0.5f || vlo > fh - 0.5f) {
                        cls = 0;   // every voxel projects outside the image: the reference touches nothing
                    } else if (ulo >= 0.0f && vlo >= 0.0f && uhi <= fw - 1.0f && vhi <= fh - 1.0f) {
                        const int iu0 = (int)ulo, iv0 = (int)vlo, iu1 = (int)uhi + 1, iv1 = (int)vhi + 1;   // inclusive, inside the image... iu1 <= w - 1 + 1
                        const int e = max(iu1 - iu0, iv1 - iv0);
#if EMF_INT_PYR2
                        const int L = max(1, 32 - __clz(max(e - 1, 0)));   // tile 2^L >= e  =>  the box spans at most 2 tiles per axis
                        constexpr int kSpan = 2;
#else
                        const int L = 32 - __clz(e >> 1);      // tile 2^L > e / 2  =>  the box spans at most 3 tiles per axis
                        constexpr int kSpan = 3;
#endif
                        if (L <= kPyrLevels) {
                            const float2* __restrict__ lv = P.pyr[L];
                            const int pw = P.pyr_w[L];
                            const int a0 = iu0 >> L, a1 = min(iu1, P.w - 1) >> L, b0 = iv0 >> L, b1 = min(iv1, P.h - 1) >> L;
                            float dmin = INFINITY, dmax = -INFINITY;
#pragma unroll
                            for (int db = 0; db < kSpan; ++db) {
                                const float2* rowp = lv + (size_t)min(b0 + db, b1) * pw;
#pragma unroll
                                for (int da = 0; da < kSpan; ++da) {
                                    const float2 mm = __ldg(rowp + min(a0 + da, a1));
                                    dmin = fminf(dmin, mm.x); dmax = fmaxf(dmax, mm.y);
                                }
                            }
                            // |pc| / lambda(pixel) is within g_rel * pc.z of pc.z for every voxel (launcher)
                            if (dmin - zhi > fmaf(zhi, P.g_rel, band)) cls = 1;
                            else if (zlo - dmax > fmaf(zlo, P.g_rel, band)) cls = 2;
                        }
                    }
                }
            }
            // ---- phase B: wholesale segments
            int known = 0;                 // voxels whose final tsdf value this lane knows
            float tv[4] = {0.f, 0.f, 0.f, 0.f};
            if (STATS) { if (cls == 1) ++st[6]; else if (cls == 2) ++st[7]; }
            if (cls == 1) {
                float* wp = V.weights + row_off + x0;
                float* tp = V.tsdf + row_off + x0;
                const float4 w4 = *reinterpret_cast<const float4*>(wp);
                const float4 t4 = *reinterpret_cast<const float4*>(tp);
                float w[4] = {w4.x, w4.y, w4.z, w4.w};
                tv[0] = t4.x; tv[1] = t4.y; tv[2] = t4.z; tv[3] = t4.w;
                known = 0xF;
                bool any = false;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // sdf >= trunc: value +1 with weight 1 (free space is never association-weighted)
                    const float ws = fadd(w[j], 1.0f);
                    if (ws > 0.0f) {
                        const float num = ffma(w[j], tv[j], 1.0f);
                        tv[j] = (num == ws && ws <= 3.0e38f) ? 1.0f : fdiv(num, ws);
                        w[j] = fminf(ws, P.max_weight);
                        any = true;
                        if (STATS) ++st[0];
                    }
                }
                if (any) {
                    *reinterpret_cast<float4*>(wp) = make_float4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<float4*>(tp) = make_float4(tv[0], tv[1], tv[2], tv[3]);
                }
            } else if (cls == 2) {
                const float4 w4 = *reinterpret_cast<const float4*>(V.weights + row_off + x0);
                float* tp = V.tsdf + row_off + x0;
                const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (w[j] == 0.0f) { tv[j] = -1.0f; known |= 1 << j; if (STATS) ++st[1]; }
                    else if (STATS) ++st[2];
                }
                if (known == 0xF) *reinterpret_cast<float4*>(tp) = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (known & (1 << j)) tp[j] = -1.0f;
                }
            }
            // ---- phase C: the other segments, one voxel per lane
            const unsigned mixed = __ballot_sync(kFull, cls == 3);
            uint32_t set_m[3] = {0u, 0u, 0u};      // segments (by owner lane) whose four voxels end up as constant m
            if (mixed) {
                if (cls == 3) s_src[wid][__popc(mixed & ((1u << lane) - 1u))] = (uint8_t)lane;
                __syncwarp();
                const int ntask = 4 * __popc(mixed);
                for (int base = 0; base < ntask; base += 32) {
                    const int k = base + lane;
                    const bool on = k < ntask;
                    const int src = on ? s_src[wid][k >> 2] : 0;
                    const int x = xbase + 4 * src + (k & 3);
                    float tfin = 2.0f;     // final tsdf of the voxel (2 = not a constant)
                    if (on) {
                        if (STATS) ++st[5];
                        float* wp = V.weights + row_off + x;
                        float* tp = V.tsdf + row_off + x;
                        const float w = *wp;
                        float tcur = *tp;
                        tfin = tcur;
                        const float cx = fmul(fsub((float)x, hx), s);
                        const float pcx = fadd(V.t[0], ffma(V.R[2], cz, ffma(V.R[0], cx, my0)));
                        const float pcy = fadd(V.t[1], ffma(V.R[5], cz, ffma(V.R[3], cx, my1)));
                        const float pcz = fadd(V.t[2], ffma(V.R[8], cz, ffma(V.R[6], cx, my2)));
                        if (!(pcz > 0.0f)) {
                            if (w == 0.0f) { *tp = 0.0f; tfin = 0.0f; }
                            if (STATS) ++st[3];
                        } else {
                            const float qx = ffma(P.K[2], pcz, fmul(P.K[0], pcx));
                            const float qy = ffma(P.K[5], pcz, fmul(P.K[4], pcy));
                            const float rz = rcp_approx(pcz);
                            const int px = round_quotient(qx, pcz, rz);
                            const int py = round_quotient(qy, pcz, rz);
                            if ((unsigned)px >= (unsigned)P.w || (unsigned)py >= (unsigned)P.h) {
                                if (STATS) ++st[4];
                            } else {
                                const float d = __ldg((const float*)((const char*)P.depth + (size_t)py * P.depth_pitch) + px);
                                if (!(d > 0.0f)) {
                                    if (w == 0.0f) { *tp = 0.0f; tfin = 0.0f; }
                                    if (STATS) ++st[3];
                                } else if (d - pcz > fmaf(pcz, P.g_rel, band)) {
                                    // clearly in front of the measurement (same guard as the segment test): free space
                                    const float ws = fadd(w, 1.0f);
                                    if (ws > 0.0f) {
                                        const float num = ffma(w, tcur, 1.0f);
                                        tcur = (num == ws && ws <= 3.0e38f) ? 1.0f : fdiv(num, ws);
                                        *tp = tcur; *wp = fminf(ws, P.max_weight);
                                        tfin = tcur;
                                        if (STATS) ++st[0];
                                    }
                                } else if (pcz - d > fmaf(pcz, P.g_rel, band)) {
                                    // clearly behind it: occluded
                                    if (w == 0.0f) { *tp = -1.0f; tfin = -1.0f; if (STATS) ++st[1]; }
                                    else if (STATS) ++st[2];
                                } else {
                                    const float inv_lambda = __ldg(P.inv_lambda + (size_t)py * P.w + px);   // exact, per pixel
                                    const float nrm = norm3(pcx, pcy, pcz);
                                    const float sdf = ffma(-nrm, inv_lambda, d);   // depth - (1/lambda)*|pc| as one FFMA (reference SASS)
                                    if (sdf >= ntrunc) {
                                        const float q = div_trunc(sdf);
                                        const float val = copysignf(fminf(1.0f, fabsf(q)), sdf);
                                        float a = 1.0f;
                                        if (sdf < V.trunc)
                                            a = __ldg((const float*)((const char*)V.assoc + (size_t)py * V.assoc_pitch) + px);
                                        const float ws = fadd(w, a);
                                        if (ws > 0.0f) {
                                            tcur = fdiv(ffma(w, tcur, fmul(val, a)), ws);
                                            *tp = tcur; *wp = fminf(ws, P.max_weight);
                                            tfin = tcur;
                                            if (STATS) ++st[0];
                                        }
                                    } else if (w == 0.0f) {
                                        *tp = -1.0f; tfin = -1.0f;
                                        if (STATS) ++st[1];
                                    } else if (STATS) ++st[2];
                                }
                            }
                        }
                    }
                    if (V.const_bits) {   // the four voxels of a segment sit in four consecutive lanes
                        int c = on ? value_code(tfin) : kMixed;
                        const int c1 = __shfl_xor_sync(kFull, c, 1); c = (c == c1) ? c : kMixed;
                        const int c2 = __shfl_xor_sync(kFull, c, 2); c = (c == c2) ? c : kMixed;
#pragma unroll
                        for (int m = 1; m <= 3; ++m)
                            set_m[m - 1] |= __reduce_or_sync(kFull, (on && (lane & 3) == 0 && c == m) ? (1u << src) : 0u);
                    }
                }
                __syncwarp();
            }
            // ---- constant-segment bitmaps: one bit per 4-voxel segment in each of three maps (all +1 / all 0 / all -1).
            //      A word covers this warp's 128-voxel chunk, so it has one owner.
            if (V.const_bits && __any_sync(kFull, cls != 0)) {
                int code = -1;    // common code of the values this lane knows; kMixed if they differ or are not constants
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (known & (1 << j)) {
                        const int c = value_code(tv[j]);
                        code = (code == -1 || code == c) ? c : kMixed;
                    }
                const bool full = (known == 0xF);
                const size_t word = (size_t)row * V.wpr + (xbase >> 7);
#pragma unroll
                for (int m = 1; m <= 3; ++m) {
                    const uint32_t set = __ballot_sync(kFull, cls != 3 && full && code == m) | set_m[m - 1];
                    // a partially known segment keeps its old bit only if what was seen agrees with it
                    const uint32_t keep = __ballot_sync(kFull, cls != 3 && !full && (code == -1 || code == m));
                    if (lane == m - 1) {
                        uint32_t* wp32 = V.const_bits + (size_t)(m - 1) * V.map_words + word;
                        *wp32 = keep ? (set | (*wp32 & keep)) : set;
                    }
                }
            }
        }
    }
    }   // batches
    if (STATS && P.stats) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned long long v = st[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            if (lane == 0 && v) atomicAdd(P.stats + k, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Brick level: one more step of "decide cheaply, compute exactly only where needed".
//
// k_integrate_seg still pays ~100 instructions per 4-voxel segment to find out that the segment is free space or occluded,
// and runs its branches on partly filled warps (12.9 / 18 / 23 of 32 lanes).  But three quarters of the voxels it touches sit
// in bricks that are free space or occluded as a whole.  So:
//  * k_brick_classify -- one thread per brick of 32 x 4 x 4 voxels (rows of 128 bytes): the eight corner voxels are
//    projected, the depth pyramid gives the smallest and largest measurement over the padded pixel box of the brick, and the
//    same guarded test as for a segment decides: outside the image (nothing to do), free space, occluded, behind the
//    camera, or mixed.  Bricks are appended to two lists (mixed / decided) with warp-aggregated atomics.
//  * k_integrate_bricks -- persistent warps take one brick at a time, the mixed ones first (they are the long items).  A
//    decided brick is a pure stream: per lane one float4 of tsdf and one of weights per z-slice, all 32 lanes on the same
//    branch, no projection at all.  A mixed brick runs k_integrate_seg's segment classification and exact per-voxel path
//    slice by slice (32 segments = 8 per row x 4 rows).
// Results: bit-identical to k_integrate_seg / the reference kernel (same canonical per-voxel arithmetic; a brick is only
// decided wholesale when every voxel of it provably takes that branch).
// ---------------------------------------------------------------------------------------------
#ifndef EMF_BRICK_SUB
#define EMF_BRICK_SUB 0
#endif
constexpr int kBrickX = 32, kBrickY = 4, kBrickZ = 4;
enum : int { kBrOut = 0, kBrFree = 1, kBrOcc = 2, kBrMixed = 3, kBrBehind = 4 };
constexpr int kBrickVolShift = 22;      // item.x = volume << 22 | brick index inside the volume; item.y = 4 x 3 bits: class of each sub-brick
constexpr int kClassifyThreads = 256;
constexpr int kBrickThreads = 256;

// class of the box of voxel centres [x0, x1] x [y0, y1] x [z0, z1] (metres, volume frame)
__device__ __forceinline__ int classify_box(const IntParams& P, const IntVol& V, float x0, float x1, float y0, float y1, float z0, float z1) {
    float zlo = INFINITY, zhi = -INFINITY, ulo = INFINITY, uhi = -INFINITY, vlo = INFINITY, vhi = -INFINITY;
    float pz[8], qx[8], qy[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float cx = (q & 1) ? x1 : x0, cy = (q & 2) ? y1 : y0, cz = (q & 4) ? z1 : z0;
        const float px = V.t[0] + (V.R[0] * cx + V.R[1] * cy + V.R[2] * cz);
        const float py = V.t[1] + (V.R[3] * cx + V.R[4] * cy + V.R[5] * cz);
        pz[q] = V.t[2] + (V.R[6] * cx + V.R[7] * cy + V.R[8] * cz);
        qx[q] = P.K[0] * px + P.K[2] * pz[q];
        qy[q] = P.K[4] * py + P.K[5] * pz[q];
        zlo = fminf(zlo, pz[q]); zhi = fmaxf(zhi, pz[q]);
    }
    const float zmargin = 1.0e-3f + 1.0e-5f * (fabsf(V.t[2]) + fabsf(x0) + fabsf(x1) + fabsf(y0) + fabsf(y1) + fabsf(z0) + fabsf(z1));
    if (zhi < -zmargin) return kBrBehind;                      // pc.z <= 0 for every voxel (TSDF.cu:349-353)
    if (!(zlo > 0.05f)) return kBrMixed;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float r = 1.0f / pz[q];
        const float u = qx[q] * r, v = qy[q] * r;
        ulo = fminf(ulo, u); uhi = fmaxf(uhi, u); vlo = fminf(vlo, v); vhi = fmaxf(vhi, v);
    }
    // pixel box of the voxels (the projection of a convex body lies in the hull of its corners'), padded by one pixel for
    // the rounding to the nearest pixel and the arithmetic above
    ulo -= 1.0f; uhi += 1.0f; vlo -= 1.0f; vhi += 1.0f;
    const float fw = (float)P.w, fh = (float)P.h;
    if (uhi < -0.5f || ulo > fw - 0.5f || vhi < -0.5f || vlo > fh - 0.5f) return kBrOut;   // the reference touches nothing
    if (!(ulo >= 0.0f && vlo >= 0.0f && uhi <= fw - 1.0f && vhi <= fh - 1.0f)) return kBrMixed;
    const int iu0 = (int)ulo, iv0 = (int)vlo, iu1 = min((int)uhi + 1, P.w - 1), iv1 = min((int)vhi + 1, P.h - 1);
    const int e = max(iu1 - iu0, iv1 - iv0) + 1;
    int L = max(1, 32 - __clz(max(e - 1, 0)));      // tile 2^L >= e  =>  the box spans at most 2 tiles per axis
    L = min(L, kPyrLevels);                          // (larger boxes: all the top-level tiles they touch)
    const float2* __restrict__ lv = P.pyr[L];
    const int pw = P.pyr_w[L];
    const int a0 = iu0 >> L, a1 = iu1 >> L, b0 = iv0 >> L, b1 = iv1 >> L;
    float dmin = INFINITY, dmax = -INFINITY;
    for (int bb = b0; bb <= b1; ++bb)
        for (int aa = a0; aa <= a1; ++aa) {
            const float2 mm = __ldg(lv + (size_t)bb * pw + aa);
            dmin = fminf(dmin, mm.x); dmax = fmaxf(dmax, mm.y);
        }
    // |pc| / lambda(pixel) is within g_rel * pc.z of pc.z for every voxel (launcher)
    const float band = V.trunc + P.g_abs;
    if (dmin - zhi > fmaf(zhi, P.g_rel, band)) return kBrFree;
    if (zlo - dmax > fmaf(zlo, P.g_rel, band)) return kBrOcc;
    return kBrMixed;
}

__global__ void __launch_bounds__(kClassifyThreads) k_brick_classify(const __grid_constant__ IntParams P) {
    constexpr unsigned kFull = 0xffffffffu;
    const int gb = blockIdx.x * kClassifyThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int kind = 0;                  // 0: nothing to do, 1: every part decided, 2: some part needs the per-segment treatment
    uint2 item = make_uint2(0u, 0u);
    if (gb < P.total_bricks) {
        int lo = 0, hi = P.n_vol - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (P.v[mid].first_brick <= gb) lo = mid; else hi = mid - 1;
        }
        const IntVol& V = P.v[lo];
        const bool gated_out = V.gate >= 0 && P.gate_counts && !(__ldg(P.gate_counts + V.gate) > P.gate_thresh);   // not visible: not integrated
        if (!gated_out) {
            const int b = gb - V.first_brick;
            const int bz = b / (V.nbx * V.nby), brem = b - bz * V.nbx * V.nby;
            const int by = brem / V.nbx, bx = brem - by * V.nbx;
            item.x = ((uint32_t)lo << kBrickVolShift) | (uint32_t)b;
            const float s = V.voxel;
            // voxel centres of the brick's corners (plain float math: every decision carries its own margin)
            const float x0 = ((float)(bx * kBrickX) - 0.5f * (float)(V.rx - 1)) * s;
            const float y0 = ((float)(by * kBrickY) - 0.5f * (float)(V.ry - 1)) * s, y1 = y0 + (float)(kBrickY - 1) * s;
            const float z0 = ((float)(bz * kBrickZ) - 0.5f * (float)(V.rz - 1)) * s, z1 = z0 + (float)(kBrickZ - 1) * s;
            const int whole = classify_box(P, V, x0, x0 + (float)(kBrickX - 1) * s, y0, y1, z0, z1);
            uint32_t sub = 0u;
            if (whole != kBrMixed) {
                sub = (uint32_t)whole * 0x249u;        // the same class in all four 3-bit fields
                kind = whole == kBrOut ? 0 : 1;
            } else if (!EMF_BRICK_SUB) {
                sub = (uint32_t)kBrMixed * 0x249u;
                kind = 2;
            } else {
                // a brick is a row of four sub-bricks of 8 x 4 x 4 voxels: their pixel boxes are a quarter as wide
                // (measured on the bench scene: 72 % instead of 89 % of the objects' bricks keep a mixed part, but the frame's
                //  integrate gets no faster -- the time is in the per-voxel path of the mixed parts -- and the classification
                //  costs 27 instead of 13 us: off by default)
                bool any_mixed = false, any_work = false;
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    const float xa = x0 + (float)(8 * q) * s;
                    const int c = classify_box(P, V, xa, xa + 7.0f * s, y0, y1, z0, z1);
                    sub |= (uint32_t)c << (3 * q);
                    any_mixed = any_mixed || c == kBrMixed;
                    any_work = any_work || c != kBrOut;
                }
                kind = any_mixed ? 2 : (any_work ? 1 : 0);
            }
            item.y = sub;
        }
    }
    const unsigned mm = __ballot_sync(kFull, kind == 2);
    const unsigned mw = __ballot_sync(kFull, kind == 1);
    int base_m = 0, base_w = 0;
    if (lane == 0) {
        if (mm) base_m = atomicAdd(P.brick_counters + 0, __popc(mm));
        if (mw) base_w = atomicAdd(P.brick_counters + 1, __popc(mw));
    }
    base_m = __shfl_sync(kFull, base_m, 0); base_w = __shfl_sync(kFull, base_w, 0);
    const unsigned below = (1u << lane) - 1u;
    if (kind == 2) P.list_mixed[base_m + __popc(mm & below)] = item;
    else if (kind == 1) P.list_whole[base_w + __popc(mw & below)] = item;
}

// ---- A/B (EMF_INT_TMA=1, off by default): the depth pixels a mixed brick slice can touch are staged in shared memory with
// the TMA engine's bulk copies (cp.async.bulk, one per image row of the slice's pixel box, completion on an mbarrier), and the
// per-voxel path gathers from there.  Measured on the bench frame: see DESIGN.md section 9 (no gain: the gathers hit L1 / L2 and
// the kernel is bound by issue slots; the tile costs occupancy).
#ifndef EMF_INT_TMA
#define EMF_INT_TMA 0
#endif
constexpr int kTmaRows = 12, kTmaCols = 96;      // (8 warps x 12 x 96 floats = 36 KB of static shared memory)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool STATS>
__global__ void __launch_bounds__(kBrickThreads, EMF_INT_TMA ? 3 : 4) k_integrate_bricks(const __grid_constant__ IntParams P) {
    __shared__ uint8_t s_src[kBrickThreads / 32][32];
#if EMF_INT_TMA
    __shared__ __align__(128) float s_tile[kBrickThreads / 32][kTmaRows][kTmaCols];
    __shared__ __align__(8) unsigned long long s_bar[kBrickThreads / 32];
    uint32_t tma_phase = 0;
    if ((threadIdx.x & 31) == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[threadIdx.x >> 5])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
#endif
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int seg = lane & 7, yy = lane >> 3;          // this lane's 4-voxel segment inside a 32 x 4 slice of a brick
    unsigned long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const float fw = (float)P.w, fh = (float)P.h;
    const int n_m = P.brick_counters[0], n_w = P.brick_counters[1];
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(P.brick_counters + 2, 1);
        i = __shfl_sync(kFull, i, 0);
        if (i >= n_m + n_w) break;
        const uint2 item = i < n_m ? __ldg(P.list_mixed + i) : __ldg(P.list_whole + (i - n_m));
        const int cls = (int)((item.y >> (3 * (seg >> 1))) & 7u);      // class of this lane's sub-brick
        const IntVol& V = P.v[(item.x >> kBrickVolShift) & 127u];
        if (V.gate >= 0 && P.gate_counts && !(__ldg(P.gate_counts + V.gate) > P.gate_thresh)) continue;   // (classified before the gate was known)
        const int b = (int)(item.x & ((1u << kBrickVolShift) - 1u));
        const int rx = V.rx, ry = V.ry;
        const int bz = b / (V.nbx * V.nby), brem = b - bz * V.nbx * V.nby;
        const int by = brem / V.nbx, bx = brem - by * V.nbx;
        const int x0 = bx * kBrickX + 4 * seg, y = by * kBrickY + yy, z0 = bz * kBrickZ;
        const int64_t plane = (int64_t)rx * ry;
        const int64_t off0 = ((int64_t)z0 * ry + y) * rx + x0;
        if (i >= n_m) {
            // ---- every sub-brick decided: a pure stream, no projection at all
            if (cls == kBrFree) {
                // sdf >= trunc for every voxel: value +1 with weight 1 (free space is never association-weighted)
                float4 w4[kBrickZ], t4[kBrickZ];
#pragma unroll
                for (int k = 0; k < kBrickZ; ++k) {
                    w4[k] = *reinterpret_cast<const float4*>(V.weights + off0 + k * plane);
                    t4[k] = *reinterpret_cast<const float4*>(V.tsdf + off0 + k * plane);
                }
#pragma unroll
                for (int k = 0; k < kBrickZ; ++k) {
                    float w[4] = {w4[k].x, w4[k].y, w4[k].z, w4[k].w}, tv[4] = {t4[k].x, t4[k].y, t4[k].z, t4[k].w};
                    bool any = false;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float ws = fadd(w[j], 1.0f);
                        if (ws > 0.0f) {
                            const float num = ffma(w[j], tv[j], 1.0f);
                            tv[j] = (num == ws && ws <= 3.0e38f) ? 1.0f : fdiv(num, ws);
                            w[j] = fminf(ws, P.max_weight);
                            any = true;
                            if (STATS) ++st[0];
                        }
                    }
                    if (any) {
                        *reinterpret_cast<float4*>(V.weights + off0 + k * plane) = make_float4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<float4*>(V.tsdf + off0 + k * plane) = make_float4(tv[0], tv[1], tv[2], tv[3]);
                    }
                }
                if (STATS) st[6] += kBrickZ;
            } else if (cls == kBrOcc || cls == kBrBehind) {
                // far behind the surface: never-seen voxels are marked -1 (TSDF.cu:397-399); behind the camera: un-marked to 0 (:349-353)
                const float mark = cls == kBrOcc ? -1.0f : 0.0f;
                float4 w4[kBrickZ];
#pragma unroll
                for (int k = 0; k < kBrickZ; ++k) w4[k] = *reinterpret_cast<const float4*>(V.weights + off0 + k * plane);
#pragma unroll
                for (int k = 0; k < kBrickZ; ++k) {
                    const float w[4] = {w4[k].x, w4[k].y, w4[k].z, w4[k].w};
                    float* tp = V.tsdf + off0 + k * plane;
                    int known = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (w[j] == 0.0f) { known |= 1 << j; if (STATS && cls == kBrOcc) ++st[1]; }
                        else if (STATS && cls == kBrOcc) ++st[2];
                        if (STATS && cls == kBrBehind) ++st[3];
                    }
                    if (known == 0xF) *reinterpret_cast<float4*>(tp) = make_float4(mark, mark, mark, mark);
                    else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (known & (1 << j)) tp[j] = mark;
                    }
                }
                if (STATS && cls == kBrOcc) st[7] += kBrickZ;
            }
            continue;
        }
        // ---- mixed brick: k_integrate_seg's treatment, slice by slice
        const float s = V.voxel;
        const float hx = fmul((float)(rx - 1), 0.5f);
        const ConstDiv div_trunc(V.trunc);
        const float ntrunc = -V.trunc;
        const float band = V.trunc + P.g_abs;
        const int xbase = bx * kBrickX;
#pragma unroll 1
        for (int k = 0; k < kBrickZ; ++k) {
            const int z = z0 + k;
            const int64_t row_off = ((int64_t)z * ry + y) * rx;
            // this lane's row as a line in homogeneous pixel coordinates: q(x) = qa + x * qb (plain float math: only used
            // for decisions that carry their own safety margins)
            float qxa, qxb, qya, qyb, qza, qzb;
            {
                const float cyp = ((float)y - (float)(ry - 1) * 0.5f) * s;
                const float czp = ((float)z - (float)(V.rz - 1) * 0.5f) * s;
                const float c0 = -((float)(rx - 1) * 0.5f) * s;
                const float ax = V.t[0] + (V.R[0] * c0 + V.R[1] * cyp + V.R[2] * czp), bxx = V.R[0] * s;
                const float ay = V.t[1] + (V.R[3] * c0 + V.R[4] * cyp + V.R[5] * czp), byy = V.R[3] * s;
                const float az = V.t[2] + (V.R[6] * c0 + V.R[7] * cyp + V.R[8] * czp), bzz = V.R[6] * s;
                qxa = P.K[0] * ax + P.K[2] * az; qxb = P.K[0] * bxx + P.K[2] * bzz;
                qya = P.K[4] * ay + P.K[5] * az; qyb = P.K[4] * byy + P.K[5] * bzz;
                qza = az; qzb = bzz;
            }
            // ---- phase A: classify the segment as a whole.  0 skip, 1 free, 2 occluded, 3 per voxel, 4 behind the camera
            //      (the lanes of a sub-brick that k_brick_classify decided already know)
            int cls_s = cls == kBrMixed ? 3 : cls;          // (kBrOut / kBrFree / kBrOcc / kBrBehind = 0 / 1 / 2 / 4)
#if EMF_INT_TMA
            int bu0 = 1 << 20, bu1 = -(1 << 20), bv0 = 1 << 20, bv1 = -(1 << 20);      // pixel box of this lane's segment, if undecided
            bool nobox = false;
#endif
            if (cls == kBrMixed) {
                const float xf0 = (float)x0, xf3 = (float)(x0 + 3);
                const float zz0 = fmaf(xf0, qzb, qza), zz3 = fmaf(xf3, qzb, qza);
                const float zlo = fminf(zz0, zz3), zhi = fmaxf(zz0, zz3);
                if (zlo > 0.05f) {
                    const float r0 = rcp_approx(zz0), r3 = rcp_approx(zz3);
                    const float u0 = fmaf(xf0, qxb, qxa) * r0, u3 = fmaf(xf3, qxb, qxa) * r3;
                    const float v0 = fmaf(xf0, qyb, qya) * r0, v3 = fmaf(xf3, qyb, qya) * r3;
                    const float ulo = fminf(u0, u3) - 1.0f, uhi = fmaxf(u0, u3) + 1.0f;
                    const float vlo = fminf(v0, v3) - 1.0f, vhi = fmaxf(v0, v3) + 1.0f;
                    if (uhi < -0.5f || ulo > fw - 0.5f || vhi < -0.5f || vlo > fh - 0.5f) {
                        cls_s = 0;   // every voxel projects outside the image: the reference touches nothing
                    } else if (ulo >= 0.0f && vlo >= 0.0f && uhi <= fw - 1.0f && vhi <= fh - 1.0f) {
                        const int iu0 = (int)ulo, iv0 = (int)vlo, iu1 = (int)uhi + 1, iv1 = (int)vhi + 1;
#if EMF_INT_TMA
                        bu0 = iu0; bu1 = min(iu1, P.w - 1); bv0 = iv0; bv1 = min(iv1, P.h - 1);
#endif
                        const int e = max(iu1 - iu0, iv1 - iv0);
                        const int L = max(1, 32 - __clz(max(e - 1, 0)));   // tile 2^L >= e  =>  the box spans at most 2 tiles per axis
                        if (L <= kPyrLevels) {
                            const float2* __restrict__ lv = P.pyr[L];
                            const int pw = P.pyr_w[L];
                            const int a0 = iu0 >> L, a1 = min(iu1, P.w - 1) >> L, b0 = iv0 >> L, b1 = min(iv1, P.h - 1) >> L;
                            float dmin = INFINITY, dmax = -INFINITY;
#pragma unroll
                            for (int db = 0; db < 2; ++db) {
                                const float2* rowp = lv + (size_t)min(b0 + db, b1) * pw;
#pragma unroll
                                for (int da = 0; da < 2; ++da) {
                                    const float2 mm = __ldg(rowp + min(a0 + da, a1));
                                    dmin = fminf(dmin, mm.x); dmax = fmaxf(dmax, mm.y);
                                }
                            }
                            if (dmin - zhi > fmaf(zhi, P.g_rel, band)) cls_s = 1;
                            else if (zlo - dmax > fmaf(zlo, P.g_rel, band)) cls_s = 2;
                        }
                    }
                }
            }
            // ---- phase B: wholesale segments
            if (STATS) { if (cls_s == 1) ++st[6]; else if (cls_s == 2) ++st[7]; }
            if (cls_s == 1) {
                float* wp = V.weights + row_off + x0;
                float* tp = V.tsdf + row_off + x0;
                const float4 w4 = *reinterpret_cast<const float4*>(wp);
                const float4 t4 = *reinterpret_cast<const float4*>(tp);
                float w[4] = {w4.x, w4.y, w4.z, w4.w}, tv[4] = {t4.x, t4.y, t4.z, t4.w};
                bool any = false;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float ws = fadd(w[j], 1.0f);
                    if (ws > 0.0f) {
                        const float num = ffma(w[j], tv[j], 1.0f);
                        tv[j] = (num == ws && ws <= 3.0e38f) ? 1.0f : fdiv(num, ws);
                        w[j] = fminf(ws, P.max_weight);
                        any = true;
                        if (STATS) ++st[0];
                    }
                }
                if (any) {
                    *reinterpret_cast<float4*>(wp) = make_float4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<float4*>(tp) = make_float4(tv[0], tv[1], tv[2], tv[3]);
                }
            } else if (cls_s == 2) {
                const float4 w4 = *reinterpret_cast<const float4*>(V.weights + row_off + x0);
                float* tp = V.tsdf + row_off + x0;
                const float w[4] = {w4.x, w4.y, w4.z, w4.w};
                int known = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (w[j] == 0.0f) { known |= 1 << j; if (STATS) ++st[1]; }
                    else if (STATS) ++st[2];
                }
                if (known == 0xF) *reinterpret_cast<float4*>(tp) = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (known & (1 << j)) tp[j] = -1.0f;
                }
            } else if (cls_s == 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(V.weights + row_off + x0);
                float* tp = V.tsdf + row_off + x0;
                const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { if (w[j] == 0.0f) tp[j] = 0.0f; if (STATS) ++st[3]; }
            }
            // ---- phase C: the other segments, one voxel per lane (the canonical arithmetic of the reference, bit for bit)
            const unsigned mixed = __ballot_sync(kFull, cls_s == 3);
#if EMF_INT_TMA
            // the depth pixels the undecided segments of this slice can touch: one bulk copy per image row into shared memory
            int U0 = __reduce_min_sync(kFull, cls_s == 3 ? bu0 : (1 << 20)), U1 = __reduce_max_sync(kFull, cls_s == 3 ? bu1 : -(1 << 20));
            int V0 = __reduce_min_sync(kFull, cls_s == 3 ? bv0 : (1 << 20)), V1 = __reduce_max_sync(kFull, cls_s == 3 ? bv1 : -(1 << 20));
            U0 &= ~3;
            const int tw = ((U1 - U0 + 1) + 3) & ~3, th = V1 - V0 + 1;
            const bool staged = mixed && U1 >= U0 && tw <= kTmaCols && th >= 1 && th <= kTmaRows && U0 + tw <= P.w &&
                                (P.depth_pitch & 15) == 0 && (((uintptr_t)P.depth) & 15) == 0;
            if (staged) {
                const uint32_t bar = smem_u32(&s_bar[wid]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(tw * 4 * th)) : "memory");
                __syncwarp();
                if (lane < th) {
                    const char* src = (const char*)P.depth + (size_t)(V0 + lane) * P.depth_pitch + (size_t)U0 * 4;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(&s_tile[wid][lane][0])), "l"(src), "r"((uint32_t)(tw * 4)), "r"(bar) : "memory");
                }
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(ok) : "r"(bar), "r"(tma_phase) : "memory");
                tma_phase ^= 1u;
            }
#endif
            if (mixed) {
                if (cls_s == 3) s_src[wid][__popc(mixed & ((1u << lane) - 1u))] = (uint8_t)lane;
                __syncwarp();
                const int ntask = 4 * __popc(mixed);
                for (int base = 0; base < ntask; base += 32) {
                    const int kk = base + lane;
                    if (kk < ntask) {
                        const int src = s_src[wid][kk >> 2];
                        const int x = xbase + 4 * (src & 7) + (kk & 3);
                        const int yv = by * kBrickY + (src >> 3);
                        if (STATS) ++st[5];
                        const int64_t voff = ((int64_t)z * ry + yv) * rx + x;
                        float* wp = V.weights + voff;
                        float* tp = V.tsdf + voff;
                        const float w = *wp;
                        float tcur = *tp;
                        // (i - (R-1)/2.f) * voxelSize ; (R-1)*0.5 is exact
                        const float cx = fmul(fsub((float)x, hx), s);
                        const float cy = fmul(fsub((float)yv, fmul((float)(ry - 1), 0.5f)), s);
                        const float cz = fmul(fsub((float)z, fmul((float)(V.rz - 1), 0.5f)), s);
                        const float my0 = fmul(V.R[1], cy), my1 = fmul(V.R[4], cy), my2 = fmul(V.R[7], cy);
                        const float pcx = fadd(V.t[0], ffma(V.R[2], cz, ffma(V.R[0], cx, my0)));
                        const float pcy = fadd(V.t[1], ffma(V.R[5], cz, ffma(V.R[3], cx, my1)));
                        const float pcz = fadd(V.t[2], ffma(V.R[8], cz, ffma(V.R[6], cx, my2)));
                        if (!(pcz > 0.0f)) {
                            if (w == 0.0f) *tp = 0.0f;
                            if (STATS) ++st[3];
                        } else {
                            const float qx = ffma(P.K[2], pcz, fmul(P.K[0], pcx));
                            const float qy = ffma(P.K[5], pcz, fmul(P.K[4], pcy));
                            const float rz = rcp_approx(pcz);
                            const int px = round_quotient(qx, pcz, rz);
                            const int py = round_quotient(qy, pcz, rz);
                            if ((unsigned)px >= (unsigned)P.w || (unsigned)py >= (unsigned)P.h) {
                                if (STATS) ++st[4];
                            } else {
#if EMF_INT_TMA
                                const float d = (staged && px >= U0 && px < U0 + tw && py >= V0 && py <= V1)
                                                    ? s_tile[wid][py - V0][px - U0]
                                                    : __ldg((const float*)((const char*)P.depth + (size_t)py * P.depth_pitch) + px);
#else
                                const float d = __ldg((const float*)((const char*)P.depth + (size_t)py * P.depth_pitch) + px);
#endif
                                if (!(d > 0.0f)) {
                                    if (w == 0.0f) *tp = 0.0f;
                                    if (STATS) ++st[3];
                                } else if (d - pcz > fmaf(pcz, P.g_rel, band)) {
                                    // clearly in front of the measurement (same guard as the segment test): free space
                                    const float ws = fadd(w, 1.0f);
                                    if (ws > 0.0f) {
                                        const float num = ffma(w, tcur, 1.0f);
                                        tcur = (num == ws && ws <= 3.0e38f) ? 1.0f : fdiv(num, ws);
                                        *tp = tcur; *wp = fminf(ws, P.max_weight);
                                        if (STATS) ++st[0];
                                    }
                                } else if (pcz - d > fmaf(pcz, P.g_rel, band)) {
                                    // clearly behind it: occluded
                                    if (w == 0.0f) { *tp = -1.0f; if (STATS) ++st[1]; }
                                    else if (STATS) ++st[2];
                                } else {
                                    const float inv_lambda = __ldg(P.inv_lambda + (size_t)py * P.w + px);   // exact, per pixel
                                    const float nrm = norm3(pcx, pcy, pcz);
                                    const float sdf = ffma(-nrm, inv_lambda, d);   // depth - (1/lambda)*|pc| as one FFMA (reference SASS)
                                    if (sdf >= ntrunc) {
                                        const float q = div_trunc(sdf);
                                        const float val = copysignf(fminf(1.0f, fabsf(q)), sdf);
                                        float a = 1.0f;
                                        if (sdf < V.trunc)
                                            a = __ldg((const float*)((const char*)V.assoc + (size_t)py * V.assoc_pitch) + px);
                                        const float ws = fadd(w, a);
                                        if (ws > 0.0f) {
                                            tcur = fdiv(ffma(w, tcur, fmul(val, a)), ws);
                                            *tp = tcur; *wp = fminf(ws, P.max_weight);
                                            if (STATS) ++st[0];
                                        }
                                    } else if (w == 0.0f) {
                                        *tp = -1.0f;
                                        if (STATS) ++st[1];
                                    } else if (STATS) ++st[2];
                                }
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    if (STATS && P.stats) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned long long v = st[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            if (lane == 0 && v) atomicAdd(P.stats + k, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// any resolution / alignment: one thread per voxel, straight canonical arithmetic
// ---------------------------------------------------------------------------------------------
template <bool PINHOLE>
__global__ void __launch_bounds__(kSimpleThreads) k_integrate_simple(const __grid_constant__ IntParams P) {
    int lo = 0, hi = P.n_vol - 1;
    const int b = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (P.v[mid].first_item <= b) lo = mid; else hi = mid - 1;
    }
    const IntVol& V = P.v[lo];
    if (V.gate >= 0 && !(__ldg(P.gate_counts + V.gate) > P.gate_thresh)) return;
    const int rx = V.rx, ry = V.ry;
    const int64_t n_vox = (int64_t)rx * ry * V.rz;
    const int64_t i = (int64_t)(b - V.first_item) * kSimpleThreads + threadIdx.x;
    if (i >= n_vox) return;
    const int64_t row = i / rx;
    const int x = (int)(i - row * rx);
    const int z = (int)(row / ry);
    const int y = (int)(row - (int64_t)z * ry);
    const float s = V.voxel;
    const float cx = fmul(fsub((float)x, fmul((float)(rx - 1), 0.5f)), s);
    const float cy = fmul(fsub((float)y, fmul((float)(ry - 1), 0.5f)), s);
    const float cz = fmul(fsub((float)z, fmul((float)(V.rz - 1), 0.5f)), s);
    const float pcx = fadd(V.t[0], dot_yxz(V.R[0], V.R[1], V.R[2], cx, cy, cz));
    const float pcy = fadd(V.t[1], dot_yxz(V.R[3], V.R[4], V.R[5], cx, cy, cz));
    const float pcz = fadd(V.t[2], dot_yxz(V.R[6], V.R[7], V.R[8], cx, cy, cz));
    float* wp = V.weights + i;
    float* tp = V.tsdf + i;
    if (!(pcz > 0.0f)) { if (*wp == 0.0f) *tp = 0.0f; return; }
    float qx, qy, qz;
    if (PINHOLE) {
        qx = ffma(P.K[2], pcz, fmul(P.K[0], pcx)); qy = ffma(P.K[5], pcz, fmul(P.K[4], pcy)); qz = pcz;
    } else {
        qx = dot_yxz(P.K[0], P.K[1], P.K[2], pcx, pcy, pcz);
        qy = dot_yxz(P.K[3], P.K[4], P.K[5], pcx, pcy, pcz);
        qz = dot_yxz(P.K[6], P.K[7], P.K[8], pcx, pcy, pcz);
    }
    const int px = __float2int_rn(fdiv(qx, qz)), py = __float2int_rn(fdiv(qy, qz));
    if (px < 0 || px >= P.w || py < 0 || py >= P.h) return;
    const float d = __ldg((const float*)((const char*)P.depth + (size_t)py * P.depth_pitch) + px);
    if (!(d > 0.0f)) { if (*wp == 0.0f) *tp = 0.0f; return; }
    const float lx = fdiv(fsub((float)px, P.K[2]), P.K[0]);
    const float ly = fdiv(fsub((float)py, P.K[5]), P.K[4]);
    const float lambda = fsqrt(fadd(ffma(lx, lx, fmul(ly, ly)), 1.0f));
    const float sdf = ffma(-norm3(pcx, pcy, pcz), frcp(lambda), d);
    const float w = *wp;
    if (sdf >= -V.trunc) {
        const float val = copysignf(fminf(1.0f, fabsf(fdiv(sdf, V.trunc))), sdf);
        float a = 1.0f;
        if (sdf < V.trunc) a = __ldg((const float*)((const char*)V.assoc + (size_t)py * V.assoc_pitch) + px);
        const float ws = fadd(w, a);
        if (ws > 0.0f) {
            *tp = fdiv(ffma(w, *tp, fmul(val, a)), ws);
            *wp = fminf(ws, P.max_weight);
        }
    } else if (w == 0.0f) {
        *tp = -1.0f;
    }
}

static bool brick_path_enabled() {      // EMF_INT_BRICKS=0: A/B against k_integrate_seg
    static int v = -1;
    if (v < 0) { const char* e = getenv("EMF_INT_BRICKS"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        g_sm_count = n;
    }
    return g_sm_count;
}

constexpr int kBrickCap = 3 << 20;      // bricks per list the full workspace has room for (1024^3 + 64 x 128^3 = 2.36 M)
size_t pyramid_layout(int w, int h, size_t off[kPyrLevels + 1], int lw[kPyrLevels + 1], int lh[kPyrLevels + 1]) {
    size_t total = 0;
    off[0] = 0; lw[0] = w; lh[0] = h;
    for (int l = 1; l <= kPyrLevels; ++l) {
        lw[l] = (w + (1 << l) - 1) >> l; lh[l] = (h + (1 << l) - 1) >> l;
        off[l] = total;
        total += (((size_t)lw[l] * lh[l] * sizeof(float2)) + 255) & ~(size_t)255;
    }
    total += (((size_t)w * h * sizeof(float)) + 255) & ~(size_t)255;   // 1 / lambda per pixel
    total += 256;   // work counter of k_integrate_seg (last 256 bytes)
    return total;
}

int launch_integrate(int n_vol, const emf_volume* vols, const emf_pose* T_oc, const float* K, const emf_image* depth,
                     const emf_image* assoc, float max_weight, const int32_t* gate_counts, const int* gates,
                     int gate_thresh, unsigned long long* stats, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream, int phase = 0) {
    // phase 0: everything; 1: only what depends on the depth image and the poses alone (depth pyramid, brick classification
    // of EVERY volume: the visibility gate is applied by the integrate kernel) -- can run next to the raycast; 2: the rest
    if (n_vol <= 0 || !vols || !T_oc || !K || !assoc) return EMF_ERR_INVALID;
    if (n_vol > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    if (!image_ok(depth, 4)) return EMF_ERR_INVALID;
    IntParams P;  // ~11 KB descriptor table, passed by value as the kernel parameter
    bool rows_ok = true;
    for (int i = 0; i < n_vol; ++i) {
        const emf_volume& v = vols[i];
        if (!v.tsdf || !v.weights || !res_ok(v.res) || !image_ok(&assoc[i], 4) || !same_size(&assoc[i], depth))
            return EMF_ERR_INVALID;
        rows_ok = rows_ok && (v.res[0] % 4 == 0) && aligned16(v.tsdf) && aligned16(v.weights);
        if (v.const_bits && (v.res[0] % 4 != 0 || !aligned16(v.tsdf) || !aligned16(v.weights))) return EMF_ERR_INVALID;
    }
    int64_t items = 0;
    for (int i = 0; i < n_vol; ++i) {
        const emf_volume& v = vols[i];
        IntVol& d = P.v[i];
        d.tsdf = v.tsdf; d.weights = v.weights;
        d.const_bits = rows_ok ? v.const_bits : nullptr;   // (the one-thread-per-voxel path keeps no maps: callers pass NULL)
        d.wpr = emf_bitmap_words_per_row(v.res[0]);
        d.map_words = (size_t)d.wpr * v.res[1] * v.res[2];
        d.assoc = (const float*)assoc[i].ptr; d.assoc_pitch = assoc[i].pitch;
        for (int k = 0; k < 9; ++k) d.R[k] = T_oc[i].R[k];
        for (int k = 0; k < 3; ++k) d.t[k] = T_oc[i].t[k];
        d.rx = v.res[0]; d.ry = v.res[1]; d.rz = v.res[2];
        d.voxel = v.voxel_size; d.trunc = v.truncdist;
        d.first_item = (int)items;
        d.gate = (gates && gate_counts) ? gates[i] : -1;
        const int64_t n_vox = (int64_t)v.res[0] * v.res[1] * v.res[2];
        items += rows_ok ? (int64_t)v.res[1] * v.res[2] : (n_vox + kSimpleThreads - 1) / kSimpleThreads;
        if (items > 0x7fffffff) return EMF_ERR_UNSUPPORTED;
    }
    P.n_vol = n_vol; P.total_items = (int)items;
    P.depth = (const float*)depth->ptr; P.depth_pitch = depth->pitch; P.w = depth->width; P.h = depth->height;
    for (int k = 0; k < 9; ++k) P.K[k] = K[k];
    P.max_weight = max_weight;
    P.gate_counts = gate_counts; P.gate_thresh = gate_thresh;
    P.stats = stats;
    const bool pin = is_pinhole(K);
    // lambda(pixel) vs the voxel's own ray: the two differ by at most half a pixel in each direction, and
    // d ln(lambda) / da = a / lambda^2 <= 1/2, so | ln(lambda_voxel / lambda_pixel) | <= (1/|fx| + 1/|fy|) / 4.
    // A camera matrix that is not a plain pinhole gets an infinite guard: every voxel takes the exact path.
    P.g_rel = pin ? 1.25f * 0.25f * (1.0f / fabsf(K[0]) + 1.0f / fabsf(K[4])) + 4.0e-6f : INFINITY;
    P.g_abs = 1.0e-4f;
    if (!rows_ok) {
        if (phase == 1) return EMF_OK;
        if (pin) k_integrate_simple<true><<<(unsigned)items, kSimpleThreads, 0, stream>>>(P);
        else k_integrate_simple<false><<<(unsigned)items, kSimpleThreads, 0, stream>>>(P);
        return launch_status();
    }
    const size_t tab_bytes = (size_t)(P.w + P.h) * sizeof(float);
    const bool table = tab_bytes <= 40 * 1024;
    const size_t smem = table ? tab_bytes : 0;
    // ---- segment-level path: needs the depth pyramid workspace, a plain pinhole camera and the pixel-ray table
    {
        size_t off[kPyrLevels + 1]; int lw[kPyrLevels + 1], lh[kPyrLevels + 1];
        const size_t need = pyramid_layout(P.w, P.h, off, lw, lh);
        if (workspace && workspace_bytes >= need && pin && ((uintptr_t)workspace & 15) == 0) {
            PyrParams Q;
            Q.depth = P.depth; Q.pitch = P.depth_pitch; Q.w = P.w; Q.h = P.h;
            P.pyr[0] = nullptr; P.pyr_w[0] = 0; Q.lvl[0] = nullptr; Q.lw[0] = 0; Q.lh[0] = 0;
            for (int l = 1; l <= kPyrLevels; ++l) {
                Q.lvl[l] = (float2*)((char*)workspace + off[l]); Q.lw[l] = lw[l]; Q.lh[l] = lh[l];
                P.pyr[l] = Q.lvl[l]; P.pyr_w[l] = lw[l];
            }
            Q.work_counter = (int*)((char*)workspace + need - 256);
            P.work_counter = Q.work_counter;
            // ---- brick level: every volume made of whole 32 x 4 x 4 bricks, no bitmaps to keep, lists fit the workspace
            int64_t n_bricks = 0;
            bool bricks_ok = brick_path_enabled() && workspace_bytes > need + 256;
            for (int i = 0; i < n_vol && bricks_ok; ++i) {
                const emf_volume& v = vols[i];
                bricks_ok = v.res[0] % kBrickX == 0 && v.res[1] % kBrickY == 0 && v.res[2] % kBrickZ == 0 && !v.const_bits;
                P.v[i].nbx = v.res[0] / kBrickX; P.v[i].nby = v.res[1] / kBrickY;
                P.v[i].first_brick = (int)n_bricks;
                const int64_t nb = (int64_t)P.v[i].nbx * P.v[i].nby * (v.res[2] / kBrickZ);
                bricks_ok = bricks_ok && nb < ((int64_t)1 << kBrickVolShift) && n_vol <= 128;
                n_bricks += nb;
            }
            const size_t list_cap = workspace_bytes > need + 256 ? (workspace_bytes - need - 256) / (2 * sizeof(uint2)) : 0;
            bricks_ok = bricks_ok && n_bricks > 0 && (size_t)n_bricks <= list_cap && n_bricks < 0x7fffffff;
            Q.brick_counters = bricks_ok ? (int*)((char*)workspace + need) : nullptr;
            Q.inv_lambda = (float*)((char*)workspace + need - 256 - ((((size_t)P.w * P.h * sizeof(float)) + 255) & ~(size_t)255));
            P.inv_lambda = Q.inv_lambda;
            Q.fx = K[0]; Q.fy = K[4]; Q.cx = K[2]; Q.cy = K[5];
            const dim3 pgrid((P.w + 63) / 64, (P.h + 63) / 64);
            if (phase != 2) k_depth_pyramid<<<pgrid, 256, 0, stream>>>(Q);
            if (bricks_ok) {
                P.total_bricks = (int)n_bricks;
                P.brick_counters = Q.brick_counters;
                P.list_mixed = (uint2*)((char*)workspace + need + 256);
                P.list_whole = P.list_mixed + list_cap;
                if (phase != 2) {
                    IntParams C = P;
                    if (phase == 1) C.gate_counts = nullptr;      // classify everything: the gate is not known yet
                    k_brick_classify<<<(unsigned)((n_bricks + kClassifyThreads - 1) / kClassifyThreads), kClassifyThreads, 0, stream>>>(C);
                }
                if (phase == 1) return launch_status();
                static int bocc_s = 0, bocc_n = 0;
                int& bocc = stats ? bocc_s : bocc_n;
                if (bocc == 0) {
                    if (stats) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bocc, k_integrate_bricks<true>, kBrickThreads, 0);
                    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bocc, k_integrate_bricks<false>, kBrickThreads, 0);
                    if (bocc <= 0) bocc = 2;
                }
                const unsigned bblocks = (unsigned)((int64_t)sm_count() * bocc);
                if (stats) k_integrate_bricks<true><<<bblocks, kBrickThreads, 0, stream>>>(P);
                else k_integrate_bricks<false><<<bblocks, kBrickThreads, 0, stream>>>(P);
                return launch_status();
            }
            if (phase == 1) return launch_status();
            static int occ_s = 0, occ_n = 0;
            int& occ = stats ? occ_s : occ_n;
            if (occ == 0) {
                if (stats) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_integrate_seg<true>, kSegThreads, 0);
                else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_integrate_seg<false>, kSegThreads, 0);
                if (occ <= 0) occ = 2;
            }
            int64_t blocks = (int64_t)sm_count() * occ;
            const int64_t min_blocks = (items + kSegThreads / 32 - 1) / (kSegThreads / 32);
            if (blocks > min_blocks) blocks = min_blocks;
            if (stats) k_integrate_seg<true><<<(unsigned)blocks, kSegThreads, 0, stream>>>(P);
            else k_integrate_seg<false><<<(unsigned)blocks, kSegThreads, 0, stream>>>(P);
            return launch_status();
        }
    }
    if (phase == 1) return EMF_OK;
    for (int l = 0; l <= kPyrLevels; ++l) { P.pyr[l] = nullptr; P.pyr_w[l] = 0; }
    P.work_counter = nullptr; P.inv_lambda = nullptr;
    const dim3 block(kIntThreads);
    // persistent grid: every SM filled to the kernel's occupancy, rows dealt round-robin to warps
#define EMF_LAUNCH_ROWS(PIN, TAB)                                                                          \
    do {                                                                                                   \
        static int occ_s = 0, occ_n = 0;                                                                   \
        int& occ = stats ? occ_s : occ_n;                                                                  \
        if (occ == 0) {                                                                                    \
            if (stats) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_integrate_rows<PIN, TAB, true>, kIntThreads, smem); \
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_integrate_rows<PIN, TAB, false>, kIntThreads, smem);     \
            if (occ <= 0) occ = 2;                                                                         \
        }                                                                                                  \
        int64_t blocks = (int64_t)sm_count() * occ;                                                        \
        const int64_t min_blocks = (items + kIntThreads / 32 - 1) / (kIntThreads / 32);                    \
        if (blocks > min_blocks) blocks = min_blocks;                                                      \
        const dim3 grid((unsigned)blocks);                                                                 \
        if (stats) k_integrate_rows<PIN, TAB, true><<<grid, block, smem, stream>>>(P);                     \
        else k_integrate_rows<PIN, TAB, false><<<grid, block, smem, stream>>>(P);                          \
    } while (0)
    if (pin) { if (table) EMF_LAUNCH_ROWS(true, true); else EMF_LAUNCH_ROWS(true, false); }
    else { if (table) EMF_LAUNCH_ROWS(false, true); else EMF_LAUNCH_ROWS(false, false); }
#undef EMF_LAUNCH_ROWS
    return launch_status();
}

}  // namespace emfb

extern "C" EMF_API size_t emf_integrate_workspace_bytes(int width, int height) {
    if (width <= 0 || height <= 0) return 0;
    size_t off[emfb::kPyrLevels + 1]; int lw[emfb::kPyrLevels + 1], lh[emfb::kPyrLevels + 1];
    return emfb::pyramid_layout(width, height, off, lw, lh) + 256 + (size_t)2 * emfb::kBrickCap * sizeof(uint2);
}

extern "C" EMF_API int emf_integrate_volumes(int n_vol, const emf_volume* vols, const emf_pose* T_oc, const float K[9],
                                     const emf_image* depth, const emf_image* assoc, float max_weight,
                                     emf_stream_t stream) {
    return emfb::launch_integrate(n_vol, vols, T_oc, K, depth, assoc, max_weight, nullptr, nullptr, 0, nullptr, nullptr, 0,
                                  (cudaStream_t)stream);
}

extern "C" EMF_API int emf_integrate_volumes_gated(int n_vol, const emf_volume* vols, const emf_pose* T_oc,
                                           const float K[9], const emf_image* depth, const emf_image* assoc,
                                           float max_weight, const int32_t* gate_counts, const int* gates,
                                           int gate_thresh, uint64_t* stats, emf_stream_t stream) {
    return emfb::launch_integrate(n_vol, vols, T_oc, K, depth, assoc, max_weight, gate_counts, gates, gate_thresh,
                                  (unsigned long long*)stats, nullptr, 0, (cudaStream_t)stream);
}

extern "C" EMF_API int emf_integrate_volumes_ws(int n_vol, const emf_volume* vols, const emf_pose* T_oc, const float K[9],
                                        const emf_image* depth, const emf_image* assoc, float max_weight,
                                        const int32_t* gate_counts, const int* gates, int gate_thresh, uint64_t* stats,
                                        void* workspace, size_t workspace_bytes, emf_stream_t stream) {
    return emfb::launch_integrate(n_vol, vols, T_oc, K, depth, assoc, max_weight, gate_counts, gates, gate_thresh,
                                  (unsigned long long*)stats, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" EMF_API int emf_integrate_volumes_phase(int n_vol, const emf_volume* vols, const emf_pose* T_oc, const float K[9],
                                           const emf_image* depth, const emf_image* assoc, float max_weight,
                                           const int32_t* gate_counts, const int* gates, int gate_thresh, uint64_t* stats,
                                           void* workspace, size_t workspace_bytes, int phase, emf_stream_t stream) {
    if (phase < 0 || phase > 2) return EMF_ERR_INVALID;
    return emfb::launch_integrate(n_vol, vols, T_oc, K, depth, assoc, max_weight, gate_counts, gates, gate_thresh,
                                  (unsigned long long*)stats, workspace, workspace_bytes, (cudaStream_t)stream, phase);
}

extern "C" EMF_API int emf_update_tsdf(const emf_image* depth, const emf_image* assoc_weights, float* tsdf, float* weights,
                               const emf_pose* T_oc, const float K[9], const int res[3], float voxel_size,
                               float truncdist, float max_weight, emf_stream_t stream) {
    if (!res || !assoc_weights || !T_oc) return EMF_ERR_INVALID;
    emf_volume v = {};
    v.tsdf = tsdf; v.weights = weights;
    v.res[0] = res[0]; v.res[1] = res[1]; v.res[2] = res[2];
    v.voxel_size = voxel_size; v.truncdist = truncdist; v.id = 0;
    return emfb::launch_integrate(1, &v, T_oc, K, depth, assoc_weights, max_weight, nullptr, nullptr, 0, nullptr, nullptr, 0,
                                  (cudaStream_t)stream);
}
