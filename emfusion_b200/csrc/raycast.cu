// raycast.cu -- per-pixel volume raycast for one or many volumes in ONE launch,
// the gradient pass, and the depth-ordered composite.
//
// Replaces emf::cuda::TSDF::raycastTSDF (reference src/core/cuda/TSDF.cu:466-601),
// emf::ObjTSDF::raycast's two full-volume weight-masking passes
// (src/core/ObjTSDF.cpp:209-210), emf::TSDF::updateGradients
// (src/core/TSDF.cpp:120-123) and the ~8K+6 OpenCV launches of
// emf::EMFusion::raycast's composite (src/core/EMFusion.cpp:760-794).
//
// Differences in mechanism, none in result:
//  * the weight trilinear of a march sample is only evaluated when the sample
//    is a back-face candidate (f < 0 && f' > 0) -- its only consumer;
//  * the object weight mask (fgProb > 0.5) is applied at the 8 corners of a
//    weight gather instead of materialising raycastWeights per frame;
//  * normals come from on-the-fly forward differences when no gradient volume
//    is supplied (same single fp32 subtraction per component => same bits);
//  * a volume is only traced inside the screen rectangle of its box;
//  * the three IEEE divisions by the voxel size per march sample reuse one refined reciprocal
//    (ConstDiv, emf_math.cuh) -- the same instruction sequence the reference build executes;
//  * with a brick map (emf_volume::brick_map, bricks.cu) a warp whose rays all carry exactly +1, 0 or -1
//    "jumps": where the map certifies that the next n samples of every ray would return that same value
//    again -- such a sample changes nothing of the march state -- they are not taken, and every ray advances
//    its ray parameter by the same n fp32 additions in closed form (seq_add.h).
#include "common.cuh"
#include "seq_add.h"
#include <math.h>

namespace emfb {

struct RayVol {
    const float* tsdf;
    const float* weights;
    const float* fg_probs;   // nullable
    const int32_t* fg_box;   // nullable: inclusive voxel bounds of {fgProb > 0.5} (device memory)
    const float* grads;      // nullable (float3 per voxel)
    const uint8_t* bmap;     // nullable: brick map (bricks.cu), one byte per 8^3 brick
    int nbx, nby;            // bricks per row / rows per slice
    float* ray; size_t ray_pitch;
    float* vert; size_t vert_pitch;
    float* norm; size_t norm_pitch;
    uint8_t* mask; size_t mask_pitch;
    float R[9];              // T_CO
    float t[3];
    int rx, ry, rz;
    float voxel, trunc;
    float thr1[3], thr2[3];  // smallest v with fadd(v, pad) >= R per axis, pad = 1 / 2 (bounds tests without the addition)
    int x0, y0, x1, y1;      // screen rect (exclusive upper)
    int tiles_x;             // tiles per rect row
    int first_block;
    int tube;                // 1: tube skipping (large volume, Rx % 4 == 0, 16-byte aligned)
};

struct RayParams {
    RayVol v[EMF_MAX_VOLUMES];
    int n_vol;
    int w, h;
    float K[9];
    int32_t* hit_voxel;      // optional (single-volume API)
    int write_all;           // 1: batched semantics (ray/mask written for every pixel of the rect)
    unsigned long long* stats;   // optional: [0] tsdf samples taken [1] samples skipped by jumps
                                 //           [2] jumps [3] weight samples
};

constexpr int kWarpW = 8, kWarpH = 4;     // pixels of one warp
constexpr int kTileW = 2 * kWarpW, kTileH = 2 * kWarpH;   // CTA tile = 2 x 2 warps
constexpr int kRayThreads = kTileW * kTileH;

// trilinear TSDF sample with 32-bit element offsets (volumes are < 2^31 voxels: res_ok)
__device__ __forceinline__ float trilinear32(const float* __restrict__ vol, int rx, int plane, int lx, int ly, int lz,
                                             float vx, float vy, float vz) {
    const float ax = fsub(vx, (float)lx), ay = fsub(vy, (float)ly), az = fsub(vz, (float)lz);
    const float bx = fsub(1.0f, ax), by = fsub(1.0f, ay), bz = fsub(1.0f, az);
    const float* r00 = vol + (uint32_t)(lz * plane + ly * rx + lx);
    const float* r01 = r00 + rx;
    const float* r10 = r00 + plane;
    const float* r11 = r10 + rx;
    const float c00 = lerp1(bx, __ldg(r00), ax, __ldg(r00 + 1));
    const float c01 = lerp1(bx, __ldg(r01), ax, __ldg(r01 + 1));
    const float c10 = lerp1(bx, __ldg(r10), ax, __ldg(r10 + 1));
    const float c11 = lerp1(bx, __ldg(r11), ax, __ldg(r11 + 1));
    return lerp1(bz, lerp1(by, c00, ay, c01), az, lerp1(by, c10, ay, c11));
}

__device__ __forceinline__ float weight_at(const RayVol& V, int64_t idx) {
    float w = __ldg(V.weights + idx);
    if (V.fg_probs) w = (__ldg(V.fg_probs + idx) > 0.5f) ? w : 0.0f;
    return w;
}

__device__ __forceinline__ float trilinear_weight(const RayVol& V, float vx, float vy, float vz) {
    const TriSetup s(vx, vy, vz, V.rx, V.ry);
    const int64_t b00 = s.base, b01 = b00 + V.rx, b10 = b00 + (int64_t)V.ry * V.rx, b11 = b10 + V.rx;
    return s.combine(weight_at(V, b00), weight_at(V, b00 + 1), weight_at(V, b01), weight_at(V, b01 + 1),
                     weight_at(V, b10), weight_at(V, b10 + 1), weight_at(V, b11), weight_at(V, b11 + 1));
}

// forward-difference gradient at integer voxel (x,y,z); zero on the last plane of each axis
// (reference src/core/cuda/TSDF.cu:436-447 after the setTo(0) of src/core/TSDF.cpp:121)
__device__ __forceinline__ void grad_at(const RayVol& V, int x, int y, int z, float g[3]) {
    if (V.grads) {
        const float* p = V.grads + 3 * (((int64_t)z * V.ry + y) * V.rx + x);
        g[0] = __ldg(p); g[1] = __ldg(p + 1); g[2] = __ldg(p + 2);
        return;
    }
    if (x >= V.rx - 1 || y >= V.ry - 1 || z >= V.rz - 1) { g[0] = g[1] = g[2] = 0.f; return; }
    const float* p = V.tsdf + ((int64_t)z * V.ry + y) * V.rx + x;
    const float f = __ldg(p);
    g[0] = fsub(__ldg(p + 1), f);
    g[1] = fsub(__ldg(p + V.rx), f);
    g[2] = fsub(__ldg(p + (int64_t)V.ry * V.rx), f);
}

__device__ __forceinline__ void trilinear_grad(const RayVol& V, float vx, float vy, float vz, float out[3]) {
    const TriSetup s(vx, vy, vz, V.rx, V.ry);
    const int lx = __float2int_rz(vx), ly = __float2int_rz(vy), lz = __float2int_rz(vz);
    float g[8][3];
#pragma unroll
    for (int c = 0; c < 8; ++c) grad_at(V, lx + (c & 1), ly + ((c >> 1) & 1), lz + (c >> 2), g[c]);
#pragma unroll
    for (int k = 0; k < 3; ++k)
        out[k] = s.combine(g[0][k], g[1][k], g[2][k], g[3][k], g[4][k], g[5][k], g[6][k], g[7][k]);
}

// v < 0 || v + pad >= R on every axis, with the rounded additions folded into per-volume thresholds
// (for the non-negative thresholds and the finite, never -0.0 coordinates of a marching ray, `v < 0 || v >= thr` is one
// unsigned comparison of the bit patterns: a set sign bit compares above every threshold)
__device__ __forceinline__ bool out_of_thr(float vx, float vy, float vz, const float* thr) {
    return __float_as_uint(vx) >= __float_as_uint(thr[0]) || __float_as_uint(vy) >= __float_as_uint(thr[1]) ||
           __float_as_uint(vz) >= __float_as_uint(thr[2]);
}

// march state of one ray
struct Ray {
    float tcur, step, f;        // ray parameter, current step, previous sample
    float vx, vy, vz;           // sample position (voxel coordinates) of tcur
    float out_t, hvx, hvy, hvz, hmx, hmy, hmz;
    bool hit;
};
struct RayConst {
    float dx, dy, dz, ox, oy, oz, hxh, hyh, hzh, s, half_s, tmax;
    bool sane;   // the divisor and every |o + dir t| of this ray are inside ConstDiv's exponent window
};

// v = (R-1)/2 + (o + dir t) / s.  The three IEEE divisions share one exponent-range test (ConstDiv::fast is the
// unguarded sequence): all numerators are far below 2^64 for any finite ray (c.sane), so only a numerator that is tiny
// (or exactly zero) sends the sample through the guarded operator.
__device__ __forceinline__ void ray_position(Ray& r, const RayConst& c, const ConstDiv& div_s) {
    const float nx = ffma(c.dx, r.tcur, c.ox), ny = ffma(c.dy, r.tcur, c.oy), nz = ffma(c.dz, r.tcur, c.oz);
    if (c.sane && fminf(fminf(fabsf(nx), fabsf(ny)), fabsf(nz)) > 5.5e-20f) {
        r.vx = fadd(c.hxh, div_s.fast(nx)); r.vy = fadd(c.hyh, div_s.fast(ny)); r.vz = fadd(c.hzh, div_s.fast(nz));
    } else {
        r.vx = fadd(c.hxh, div_s(nx)); r.vy = fadd(c.hyh, div_s(ny)); r.vz = fadd(c.hzh, div_s(nz));
    }
}

// one march step of the reference algorithm (TSDF.cu:523-572); returns true when the ray is finished
template <bool STATS>
__device__ __forceinline__ bool march_step(Ray& r, const RayConst& c, const ConstDiv& div_s, const RayVol& V, int rx, int plane,
                                           unsigned long long* st) {
    r.tcur = fadd(r.tcur, r.step);
    if (!(r.tcur <= c.tmax)) return true;
    ray_position(r, c, div_s);
    if (out_of_thr(r.vx, r.vy, r.vz, V.thr2)) return false;
    const int lx = __float2int_rz(r.vx), ly = __float2int_rz(r.vy), lz = __float2int_rz(r.vz);
    const float fn = trilinear32(V.tsdf, rx, plane, lx, ly, lz, r.vx, r.vy, r.vz);
    if (STATS) ++st[0];
    // back face (TSDF.cu:532): the weight sample is only needed for this test
    if (r.f < 0.0f && fn > 0.0f) {
        if (STATS) ++st[3];
        if (trilinear_weight(V, r.vx, r.vy, r.vz) > 0.0f) return true;
    }
    if (fabsf(fn) < 1.0f) r.step = c.s;
    if (fabsf(fn) < 0.8f) r.step = c.half_s;
    if (r.f > 0.0f && fn < 0.0f) {   // front face (TSDF.cu:540)
        const float ts = fsub(r.tcur, fdiv(fmul(r.f, r.step), fsub(fn, r.f)));
        const float mx = fmul(c.dx, ts), my = fmul(c.dy, ts), mz = fmul(c.dz, ts);
        const float sx = fadd(c.hxh, div_s(fadd(c.ox, mx)));
        const float sy = fadd(c.hyh, div_s(fadd(c.oy, my)));
        const float sz = fadd(c.hzh, div_s(fadd(c.oz, mz)));
        if (out_of_thr(sx, sy, sz, V.thr2)) return false;   // reference `continue`: f keeps its old value
        if (trilinear_weight(V, sx, sy, sz) > 0.0f) {
            r.hit = true; r.out_t = ts;
            r.hvx = sx; r.hvy = sy; r.hvz = sz; r.hmx = mx; r.hmy = my; r.hmz = mz;
            return true;
        }
    }
    r.f = fn;
    return false;
}

// ---- two samples in flight (PAIR): the march is a chain of dependent gathers (position -> 8 loads -> 7 lerps -> decision),
// and at 64 registers only 8 warps per scheduler hide it.  The sample after next is therefore fetched SPECULATIVELY under
// the assumption that the next one neither ends the ray nor changes the step (true for > 95 % of the samples: free space and
// never-observed space keep the step); its loads are issued before the first sample is resolved.  Each sample is then
// resolved with exactly the arithmetic of march_step, in order; a speculative sample whose assumption failed is discarded.
struct Samp {
    float vx, vy, vz;
    float c[8];
    bool in;
};
__device__ __forceinline__ void samp_fetch(Samp& q, float t, const RayConst& c, const ConstDiv& div_s, const RayVol& V, int rx, int plane) {
    const float nx = ffma(c.dx, t, c.ox), ny = ffma(c.dy, t, c.oy), nz = ffma(c.dz, t, c.oz);
    if (c.sane && fminf(fminf(fabsf(nx), fabsf(ny)), fabsf(nz)) > 5.5e-20f) {
        q.vx = fadd(c.hxh, div_s.fast(nx)); q.vy = fadd(c.hyh, div_s.fast(ny)); q.vz = fadd(c.hzh, div_s.fast(nz));
    } else {
        q.vx = fadd(c.hxh, div_s(nx)); q.vy = fadd(c.hyh, div_s(ny)); q.vz = fadd(c.hzh, div_s(nz));
    }
    q.in = !out_of_thr(q.vx, q.vy, q.vz, V.thr2);
    if (q.in) {
        const int lx = __float2int_rz(q.vx), ly = __float2int_rz(q.vy), lz = __float2int_rz(q.vz);
        const float* r00 = V.tsdf + (uint32_t)(lz * plane + ly * rx + lx);
        const float* r01 = r00 + rx;
        const float* r10 = r00 + plane;
        const float* r11 = r10 + rx;
        q.c[0] = __ldg(r00); q.c[1] = __ldg(r00 + 1); q.c[2] = __ldg(r01); q.c[3] = __ldg(r01 + 1);
        q.c[4] = __ldg(r10); q.c[5] = __ldg(r10 + 1); q.c[6] = __ldg(r11); q.c[7] = __ldg(r11 + 1);
    }
}
__device__ __forceinline__ float samp_value(const Samp& q) {
    const float ax = fsub(q.vx, (float)__float2int_rz(q.vx)), ay = fsub(q.vy, (float)__float2int_rz(q.vy)),
                az = fsub(q.vz, (float)__float2int_rz(q.vz));
    const float bx = fsub(1.0f, ax), by = fsub(1.0f, ay), bz = fsub(1.0f, az);
    const float c00 = lerp1(bx, q.c[0], ax, q.c[1]), c01 = lerp1(bx, q.c[2], ax, q.c[3]);
    const float c10 = lerp1(bx, q.c[4], ax, q.c[5]), c11 = lerp1(bx, q.c[6], ax, q.c[7]);
    return lerp1(bz, lerp1(by, c00, ay, c01), az, lerp1(by, c10, ay, c11));
}
// the part of march_step after the TSDF sample fn at ray parameter r.tcur / position q is known; true = ray finished
template <bool STATS>
__device__ __forceinline__ bool samp_resolve(Ray& r, const RayConst& c, const ConstDiv& div_s, const RayVol& V, const Samp& q, float fn,
                                             unsigned long long* st) {
    if (STATS) ++st[0];
    if (r.f < 0.0f && fn > 0.0f) {
        if (STATS) ++st[3];
        if (trilinear_weight(V, q.vx, q.vy, q.vz) > 0.0f) return true;
    }
    if (fabsf(fn) < 1.0f) r.step = c.s;
    if (fabsf(fn) < 0.8f) r.step = c.half_s;
    if (r.f > 0.0f && fn < 0.0f) {
        const float ts = fsub(r.tcur, fdiv(fmul(r.f, r.step), fsub(fn, r.f)));
        const float mx = fmul(c.dx, ts), my = fmul(c.dy, ts), mz = fmul(c.dz, ts);
        const float sx = fadd(c.hxh, div_s(fadd(c.ox, mx)));
        const float sy = fadd(c.hyh, div_s(fadd(c.oy, my)));
        const float sz = fadd(c.hzh, div_s(fadd(c.oz, mz)));
        if (out_of_thr(sx, sy, sz, V.thr2)) return false;
        if (trilinear_weight(V, sx, sy, sz) > 0.0f) {
            r.hit = true; r.out_t = ts;
            r.hvx = sx; r.hvy = sy; r.hvz = sz; r.hmx = mx; r.hmy = my; r.hmz = mz;
            return true;
        }
    }
    r.f = fn;
    return false;
}
// at most max_iters pair iterations (<= 2 max_iters samples); true = the ray is finished
template <bool STATS>
__device__ __forceinline__ bool march_pairs(Ray& r, const RayConst& c, const ConstDiv& div_s, const RayVol& V, int rx, int plane,
                                            unsigned long long* st, int max_iters = 0x7fffffff) {
    for (int it = 0; it < max_iters; ++it) {
        const float t1 = fadd(r.tcur, r.step);
        if (!(t1 <= c.tmax)) return true;
        const float t2 = fadd(t1, r.step);
        const float step0 = r.step;
        Samp a, b;
        samp_fetch(a, t1, c, div_s, V, rx, plane);
        const bool have2 = t2 <= c.tmax;
        b.in = false;
        if (have2) samp_fetch(b, t2, c, div_s, V, rx, plane);
        r.tcur = t1;
        if (a.in && samp_resolve<STATS>(r, c, div_s, V, a, samp_value(a), st)) return true;
        if (!have2 || r.step != step0) continue;     // the speculation failed (or there is no second sample)
        r.tcur = t2;
        if (b.in && samp_resolve<STATS>(r, c, div_s, V, b, samp_value(b), st)) return true;
    }
    return false;
}

// ---- tube skipping (large volumes; compiled in, OFF by default: EMF_RAY_TUBE -- see DESIGN.md section 4.3 for why).
// The rays of a warp -- an 8 x 4 pixel patch -- run inside a tube a few voxels wide.  A
// march sample whose eight corners all hold exactly +1 (observed free space) returns exactly +1 again and changes nothing
// of the march state but the ray parameter; the same goes for exactly 0 (never observed) once the step is half a voxel.
// When every ray of the warp carries that constant, the warp reads the voxels of the box its next n samples can touch --
// one row of the box per lane, 128-bit loads: ~3 % of the loads and instructions those samples would cost -- and if all of
// them hold the constant, every ray advances its parameter by the same n fp32 additions in closed form (seq_add.h) and
// nothing else happens.  Otherwise the warp marches a round of samples and tries again later.  No auxiliary structure is
// kept: the certificate is the volume itself, read coalesced.
constexpr int kTubeSamples = 32;     // samples per skip (per ray)
constexpr int kTubeRound = 8;        // pair iterations of a marching round
template <bool STATS>
__device__ __forceinline__ void march_tube(Ray& r, const RayConst& c, const ConstDiv& div_s, const RayVol& V, int rx, int plane,
                                           bool done, unsigned long long* st) {
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const float inv_s = 1.0f / c.s;
    int backoff = 0;
    while (!__all_sync(kFull, done)) {
        if (backoff == 0) {
            const bool one = done || r.f == 1.0f;
            const bool zero = done || (__float_as_uint(r.f) == 0u && r.step == c.half_s);
            const bool all1 = __all_sync(kFull, one);
            const bool all0 = !all1 && __all_sync(kFull, zero);
            bool skipped = false;
            if (all1 || all0) {
                int n = kTubeSamples;
                if (!done) {
                    // samples left before the ray passes tmax, conservatively (the additions round) ...
                    float room = (c.tmax - r.tcur) / r.step * 0.999f - 1.0f;
                    // ... and as many as keep the box of an oblique ray narrow (<= ~10 voxels in x, ~5 in y)
                    const float per = r.step * inv_s;
                    room = fminf(room, fminf(10.0f / fmaxf(fabsf(c.dx) * per, 1e-6f), 5.0f / fmaxf(fabsf(c.dy) * per, 1e-6f)));
                    n = room >= (float)kTubeSamples ? kTubeSamples : (room > 0.0f ? (int)room : 0);
                }
                n = __reduce_min_sync(kFull, n);
                if (n >= 8) {
                    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
                    if (!done) {
                        const float ta = r.tcur + r.step, tb = r.tcur + r.step * (float)n;
                        const float pa[3] = {c.hxh + (c.ox + c.dx * ta) * inv_s, c.hyh + (c.oy + c.dy * ta) * inv_s, c.hzh + (c.oz + c.dz * ta) * inv_s};
                        const float pb[3] = {c.hxh + (c.ox + c.dx * tb) * inv_s, c.hyh + (c.oy + c.dy * tb) * inv_s, c.hzh + (c.oz + c.dz * tb) * inv_s};
#pragma unroll
                        for (int k = 0; k < 3; ++k) {   // base voxels floor(v) .. floor(v) + 1, one voxel of slack for the approximations
                            lo[k] = (int)floorf(fminf(pa[k], pb[k])) - 1;
                            hi[k] = (int)floorf(fmaxf(pa[k], pb[k])) + 2;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) { lo[k] = __reduce_min_sync(kFull, lo[k]); hi[k] = __reduce_max_sync(kFull, hi[k]); }
                    // (samples outside the volume's inner bounds are skipped by the march itself: only voxels that exist matter)
                    const int x0 = max(lo[0], 0) & ~3, x1 = min(hi[0], V.rx - 1);
                    const int y0 = max(lo[1], 0), y1 = min(hi[1], V.ry - 1), z0 = max(lo[2], 0), z1 = min(hi[2], V.rz - 1);
                    const int nq = x1 >= x0 ? (x1 - x0) / 4 + 1 : 0, ny = y1 - y0 + 1, rows = ny * (z1 - z0 + 1);
                    if (nq >= 1 && nq <= 6 && ny >= 1 && rows >= 1 && rows <= 12 * 32) {
                        const uint32_t want = all1 ? 0x3f800000u : 0u;
                        bool ok = true;
                        for (int rr = lane; rr < rows; rr += 32) {
                            const int zz = z0 + rr / ny, yy = y0 + rr - (rr / ny) * ny;
                            const uint4* p = reinterpret_cast<const uint4*>(V.tsdf + ((size_t)zz * V.ry + yy) * V.rx + x0);
                            for (int q = 0; q < nq; ++q) {
                                const uint4 v = __ldg(p + q);
                                ok = ok && v.x == want && v.y == want && v.z == want && v.w == want;
                            }
                        }
                        if (__all_sync(kFull, ok)) {
                            if (!done) {
                                r.tcur = emf_seq_add(r.tcur, r.step, n);
                                if (STATS) { st[1] += n; ++st[2]; }
                            }
                            skipped = true;
                        }
                    }
                }
            }
            if (skipped) continue;
            backoff = (all1 || all0) ? 2 : 1;
        } else {
            --backoff;
        }
        if (!done) done = march_pairs<STATS>(r, c, div_s, V, rx, plane, st, kTubeRound);
    }
}

// JUMP = false: every lane marches its own ray (the lean default).
// JUMP = true : the warp stays in lockstep so that jumps through certified constant regions can be agreed on with warp
//               collectives (needs emf_volume::brick_map on at least one volume of the launch).
#ifndef EMF_RAY_PAIR
#define EMF_RAY_PAIR 1
#endif
#ifndef EMF_RAY_TUBE
#define EMF_RAY_TUBE 0
#endif
#ifndef EMF_RAY_MINB
#define EMF_RAY_MINB 8
#endif
template <bool STATS, bool JUMP>
__global__ void __launch_bounds__(kRayThreads, EMF_RAY_MINB) k_raycast(const __grid_constant__ RayParams P) {
    unsigned long long st[4] = {0, 0, 0, 0};
    int lo = 0, hi = P.n_vol - 1;
    const int b = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (P.v[mid].first_block <= b) lo = mid; else hi = mid - 1;
    }
    const RayVol& V = P.v[lo];
    const int lb = b - V.first_block;
    const int ty = lb / V.tiles_x, tx = lb - ty * V.tiles_x;
    // warp = kWarpW x kWarpH pixel patch inside the CTA tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xr = V.x0 + tx * kTileW + (warp & 1) * kWarpW + (lane % kWarpW);
    const int yr = V.y0 + ty * kTileH + (warp >> 1) * kWarpH + (lane / kWarpW);
    const bool valid = xr < V.x1 && yr < V.y1;
    if (!JUMP && !valid && !V.tube) return;   // (tube skipping uses warp collectives: idle lanes stay)
    const int x = min(xr, V.x1 - 1), y = min(yr, V.y1 - 1);   // (JUMP: lanes outside the rectangle idle in the loop)

    float* ray_px = (float*)((char*)V.ray + (size_t)y * V.ray_pitch) + x;
    uint8_t* mask_px = V.mask + (size_t)y * V.mask_pitch + x;
    Ray r;
    r.out_t = 0.0f; r.hit = false;
    r.hvx = r.hvy = r.hvz = r.hmx = r.hmy = r.hmz = 0.f;
    RayConst c;
    c.s = V.voxel;
    const float s = c.s;
    const float ux = fdiv(fsub((float)x, P.K[2]), P.K[0]);
    const float uy = fdiv(fsub((float)y, P.K[5]), P.K[4]);
    // rot_CO * (ux, uy, 1)
    const float rayx = fadd(V.R[2], ffma(V.R[0], ux, fmul(V.R[1], uy)));
    const float rayy = fadd(V.R[5], ffma(V.R[3], ux, fmul(V.R[4], uy)));
    const float rayz = fadd(V.R[8], ffma(V.R[6], ux, fmul(V.R[7], uy)));
    const float rn = norm3(rayx, rayy, rayz);
    c.dx = fdiv(rayx, rn); c.dy = fdiv(rayy, rn); c.dz = fdiv(rayz, rn);
    // boxBounds = (volSize - 1) / 2 * voxelSize with INTEGER division (TSDF.cu:490)
    const float bx = fmul((float)((V.rx - 1) / 2), s);
    const float by = fmul((float)((V.ry - 1) / 2), s);
    const float bz = fmul((float)((V.rz - 1) / 2), s);
    c.ox = V.t[0]; c.oy = V.t[1]; c.oz = V.t[2];
    const float tin = fmaxf(fmaxf(fdiv(fsub(c.dx > 0.f ? -bx : bx, c.ox), c.dx), fdiv(fsub(c.dy > 0.f ? -by : by, c.oy), c.dy)),
                            fdiv(fsub(c.dz > 0.f ? -bz : bz, c.oz), c.dz));
    const float tout = fminf(fminf(fdiv(fsub(c.dx > 0.f ? bx : -bx, c.ox), c.dx), fdiv(fsub(c.dy > 0.f ? by : -by, c.oy), c.dy)),
                             fdiv(fsub(c.dz > 0.f ? bz : -bz, c.oz), c.dz));
    r.tcur = fadd(s, tin);
    c.tmax = fsub(tout, s);
    const float old = (P.write_all || !valid) ? 0.0f : *ray_px;   // in/out far clip (TSDF.cu:496-500)
    if (old != 0.0f) c.tmax = fminf(old, c.tmax);

    c.hxh = fmul((float)(V.rx - 1), 0.5f); c.hyh = fmul((float)(V.ry - 1), 0.5f); c.hzh = fmul((float)(V.rz - 1), 0.5f);
    c.half_s = fmul(s, 0.5f);
    bool culled = false;
    if (V.fg_box) {
        // ObjTSDF::raycast: a hit needs a positive masked weight, i.e. a corner voxel with fgProb > 0.5.  Both the march
        // sample and the refined crossing lie on the ray, so a ray can only hit while it is inside the box of those voxels
        // (padded by 3 voxels against rounding): clip the march there; a ray that misses the box is not marched at all.
        const float inv_s = 1.0f / s;
        const float p0[3] = {c.hxh + c.ox * inv_s, c.hyh + c.oy * inv_s, c.hzh + c.oz * inv_s};
        const float dv[3] = {c.dx * inv_s, c.dy * inv_s, c.dz * inv_s};
        float ta = -INFINITY, tb = INFINITY;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float lo_k = (float)__ldg(V.fg_box + k) - 4.0f, hi_k = (float)__ldg(V.fg_box + 3 + k) + 4.0f;
            if (fabsf(dv[k]) > 1e-12f) {
                const float t0 = (lo_k - p0[k]) / dv[k], t1 = (hi_k - p0[k]) / dv[k];
                ta = fmaxf(ta, fminf(t0, t1)); tb = fminf(tb, fmaxf(t0, t1));
            } else if (p0[k] < lo_k || p0[k] > hi_k) {
                tb = -INFINITY;
            }
        }
        if (__ldg(V.fg_box) > __ldg(V.fg_box + 3) || !(ta <= tb)) culled = true;
        else c.tmax = fminf(c.tmax, tb);
    }
    const ConstDiv div_s(s);
    c.sane = div_s.ok && fabsf(c.ox) + fabsf(c.oy) + fabsf(c.oz) + fabsf(r.tcur) + fabsf(c.tmax) + V.trunc < 1.0e18f;
    const int rx = V.rx, plane = V.rx * V.ry;
    r.step = V.trunc;
    r.vx = r.vy = r.vz = r.f = 0.f;
    bool done = !valid || culled || r.tcur >= c.tmax;
    if (!done) {
        for (;;) {   // coarse skip (TSDF.cu:509-515)
            ray_position(r, c, div_s);
            if (out_of_thr(r.vx, r.vy, r.vz, V.thr1) && r.tcur < c.tmax) r.tcur = fadd(r.step, r.tcur);
            else break;
        }
        // still outside => the reference's march loop cannot run (tcur >= tmax): defined as no hit
        if (!out_of_thr(r.vx, r.vy, r.vz, V.thr1) && r.vx == r.vx && r.vy == r.vy && r.vz == r.vz) {
            r.f = trilinear(V.tsdf, V.rx, V.ry, r.vx, r.vy, r.vz);
            if (fabsf(r.f) < 1.0f) r.step = s;
            if (fabsf(r.f) < 0.8f) r.step = c.half_s;
        } else {
            done = true;
        }
    }
    if (!JUMP) {
        if (V.tube) {
            march_tube<STATS>(r, c, div_s, V, rx, plane, done, st);
            if (!valid) return;
        } else if (!done) {
#if EMF_RAY_PAIR
            march_pairs<STATS>(r, c, div_s, V, rx, plane, st);
#else
            while (!march_step<STATS>(r, c, div_s, V, rx, plane, st)) {}
#endif
        }
    } else {
        constexpr unsigned kFull = 0xffffffffu;
        const uint8_t* __restrict__ bmap = V.bmap;
        const float frx = (float)V.rx, fry = (float)V.ry, frz = (float)V.rz;
        // jump geometry: voxels advanced per metre of ray parameter along each axis (and the largest of them)
        const float ax_m = fdiv(fabsf(c.dx), s), ay_m = fdiv(fabsf(c.dy), s), az_m = fdiv(fabsf(c.dz), s);
        const float vox_per_m = fmaxf(ax_m, fmaxf(ay_m, az_m));
        int wait = 0;   // (warp-uniform) march steps to take before the brick map is consulted again after a failed attempt
        while (!__all_sync(kFull, done)) {
            // ---- jump: (vx, vy, vz) is the sample position of tcur.  Where the brick map certifies that every sample the
            //      next n march steps would take returns exactly f, those steps change nothing but tcur.  The warp jumps
            //      together (n = the smallest count any of its marching rays is certified for) and so stays in lockstep.
            if (bmap) {
                const bool isconst = done || r.f == 1.0f || r.f == 0.0f || r.f == -1.0f;
                if (__all_sync(kFull, isconst)) {
                    if (wait == 0) {
                        int n = 0x7fffffff;
                        if (!done) {
                            n = 0;
                            if (r.vx >= 0.0f && r.vy >= 0.0f && r.vz >= 0.0f && r.vx < frx && r.vy < fry && r.vz < frz) {
                                const int bxi = __float2int_rz(r.vx) >> 3, byi = __float2int_rz(r.vy) >> 3, bzi = __float2int_rz(r.vz) >> 3;
                                const unsigned e = __ldg(bmap + (unsigned)((bzi * V.nby + byi) * V.nbx + bxi));
                                const unsigned want = r.f == 1.0f ? 1u : (r.f == 0.0f ? 2u : 3u);
                                if ((e >> 4) == want) {
                                    const unsigned D = e & 7u;
                                    const float inv_step = __fdividef(0.999f, r.step);
                                    float nf = 0.0f;
                                    if (D >= 2u) {
                                        // every brick within D - 1 bricks holds the constant: any sample whose base voxel moves
                                        // less than 8 (D - 1) - 1 voxels (Chebyshev) from here is certified
                                        nf = ((float)((D - 1u) * 8u) - 1.25f) * __fdividef(inv_step, vox_per_m);
                                    } else if (e & 8u) {
                                        // the 2 x 2 x 2 block of bricks starting at this one holds the constant: certified while
                                        // the base voxel stays in [8 b, 8 b + 14] on every axis
                                        const float lox = r.vx - (float)(bxi << 3), loy = r.vy - (float)(byi << 3), loz = r.vz - (float)(bzi << 3);
                                        const float rx_ = (c.dx > 0.0f ? 15.0f - lox : lox) - 0.05f;
                                        const float ry_ = (c.dy > 0.0f ? 15.0f - loy : loy) - 0.05f;
                                        const float rz_ = (c.dz > 0.0f ? 15.0f - loz : loz) - 0.05f;
                                        nf = fminf(fminf(__fdividef(rx_, fmaxf(ax_m, 1e-12f)), __fdividef(ry_, fmaxf(ay_m, 1e-12f))),
                                                   __fdividef(rz_, fmaxf(az_m, 1e-12f))) * inv_step;
                                    }
                                    n = nf >= 1.0f ? (nf < 4096.0f ? (int)nf : 4096) : 0;
                                }
                            }
                        }
                        n = __reduce_min_sync(kFull, n);
                        if (n >= 1) {
                            if (!done) {
                                r.tcur = emf_seq_add(r.tcur, r.step, n);
                                if (STATS) { st[1] += n; ++st[2]; }
                                if (!(r.tcur <= c.tmax)) done = true;    // the ray ends inside the certified region: no hit
                                else ray_position(r, c, div_s);
                            }
                            continue;
                        }
                        wait = 1;
                    } else {
                        --wait;
                    }
                }
            }
            if (!done) done = march_step<STATS>(r, c, div_s, V, rx, plane, st);
        }
        if (!valid) return;
    }
    const bool hit = r.hit;
    const float out_t = r.out_t, hvx = r.hvx, hvy = r.hvy, hvz = r.hvz, hmx = r.hmx, hmy = r.hmy, hmz = r.hmz;

    if (hit) {
        float g[3];
        trilinear_grad(V, hvx, hvy, hvz, g);
        float* vp = (float*)((char*)V.vert + (size_t)y * V.vert_pitch) + 3 * x;
        float* np = (float*)((char*)V.norm + (size_t)y * V.norm_pitch) + 3 * x;
        // transpose(rot_CO) * (t* dir)
        vp[0] = dot_yxz(V.R[0], V.R[3], V.R[6], hmx, hmy, hmz);
        vp[1] = dot_yxz(V.R[1], V.R[4], V.R[7], hmx, hmy, hmz);
        vp[2] = dot_yxz(V.R[2], V.R[5], V.R[8], hmx, hmy, hmz);
        const float gn = norm3(g[0], g[1], g[2]);
        const float nx = fdiv(g[0], gn), ny = fdiv(g[1], gn), nz = fdiv(g[2], gn);
        np[0] = dot_yxz(V.R[0], V.R[3], V.R[6], nx, ny, nz);
        np[1] = dot_yxz(V.R[1], V.R[4], V.R[7], nx, ny, nz);
        np[2] = dot_yxz(V.R[2], V.R[5], V.R[8], nx, ny, nz);
        *ray_px = out_t;
        *mask_px = 1;
        if (P.hit_voxel) {
            int32_t* hv = P.hit_voxel + 3 * ((size_t)y * P.w + x);
            hv[0] = __float2int_rz(hvx); hv[1] = __float2int_rz(hvy); hv[2] = __float2int_rz(hvz);
        }
    } else if (P.write_all) {
        *ray_px = 0.0f;
        *mask_px = 0;
    }
    if (STATS && P.stats) {
#pragma unroll
        for (int k = 0; k < 4; ++k) if (st[k]) atomicAdd(P.stats + k, st[k]);
    }
}

// smallest float v with fl(v + pad) >= R: `v + pad >= R` (reference bounds tests, TSDF.cu:510,526,547) <=> v >= threshold,
// because rounding is monotone
static float pad_threshold(int R, float pad) {
    const float fr = (float)R;
    float v = fr - pad;
    for (;;) {
        const float p = nextafterf(v, -INFINITY);
        volatile float sum = p + pad;
        if (sum >= fr) v = p; else break;
    }
    for (;;) {   // (and never too low)
        volatile float sum = v + pad;
        if (sum >= fr) break;
        v = nextafterf(v, INFINITY);
    }
    return v;
}

static int fill_ray_vol(RayVol& d, const emf_volume& v, const emf_pose& T, const emf_image* ray,
                        const emf_image* vert, const emf_image* norm, const emf_image* mask, const int* rect,
                        int w, int h) {
    if (!v.tsdf || !v.weights || !res_ok(v.res)) return EMF_ERR_INVALID;
    if (!image_ok(ray, 4) || !image_ok(vert, 12) || !image_ok(norm, 12) || !image_ok(mask, 1)) return EMF_ERR_INVALID;
    if (ray->width != w || ray->height != h || !same_size(ray, vert) || !same_size(ray, norm) || !same_size(ray, mask))
        return EMF_ERR_INVALID;
    d.tsdf = v.tsdf; d.weights = v.weights; d.fg_probs = v.fg_probs; d.grads = v.grads;
    d.fg_box = v.fg_probs ? v.fg_box : nullptr;
    d.bmap = (v.brick_map && v.const_bits && v.res[0] % 4 == 0) ? v.brick_map : nullptr;
    d.nbx = (v.res[0] + 7) / 8; d.nby = (v.res[1] + 7) / 8;
    d.ray = (float*)ray->ptr; d.ray_pitch = ray->pitch;
    d.vert = (float*)vert->ptr; d.vert_pitch = vert->pitch;
    d.norm = (float*)norm->ptr; d.norm_pitch = norm->pitch;
    d.mask = (uint8_t*)mask->ptr; d.mask_pitch = mask->pitch;
    for (int k = 0; k < 9; ++k) d.R[k] = T.R[k];
    for (int k = 0; k < 3; ++k) d.t[k] = T.t[k];
    d.rx = v.res[0]; d.ry = v.res[1]; d.rz = v.res[2];
    d.voxel = v.voxel_size; d.trunc = v.truncdist;
    for (int k = 0; k < 3; ++k) { d.thr1[k] = pad_threshold(v.res[k], 1.0f); d.thr2[k] = pad_threshold(v.res[k], 2.0f); }
    int x0 = 0, y0 = 0, x1 = w, y1 = h;
    if (rect) {
        x0 = rect[0] < 0 ? 0 : rect[0]; y0 = rect[1] < 0 ? 0 : rect[1];
        x1 = rect[2] > w ? w : rect[2]; y1 = rect[3] > h ? h : rect[3];
        if (x1 < x0) x1 = x0;
        if (y1 < y0) y1 = y0;
    }
    d.x0 = x0; d.y0 = y0; d.x1 = x1; d.y1 = y1;
    d.tiles_x = (x1 - x0 + kTileW - 1) / kTileW;
#if EMF_RAY_TUBE
    // long rays through a large grid: the background.  (Object rays are clipped to a few dozen samples: nothing to skip.)
    d.tube = ((int64_t)v.res[0] * v.res[1] * v.res[2] >= ((int64_t)1 << 24) && v.res[0] % 4 == 0 && aligned16(v.tsdf)) ? 1 : 0;
#else
    d.tube = 0;
#endif
    return EMF_OK;
}

// ---------------------------------------------------------------------------------------------
// gradient volume (only for consumers that want the materialised float3 array)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_grads(const float* __restrict__ tsdf, float* __restrict__ grads,
                                               int rx, int ry, int rz) {
    const int64_t n = (int64_t)rx * ry * rz;
    const int64_t plane = (int64_t)rx * ry;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / rx;
        const int x = (int)(i - row * rx);
        const int z = (int)(row / ry);
        const int y = (int)(row - (int64_t)z * ry);
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (x < rx - 1 && y < ry - 1 && z < rz - 1) {
            const float f = __ldg(tsdf + i);
            gx = fsub(__ldg(tsdf + i + 1), f);
            gy = fsub(__ldg(tsdf + i + rx), f);
            gz = fsub(__ldg(tsdf + i + plane), f);
        }
        float* g = grads + 3 * i;
        g[0] = gx; g[1] = gy; g[2] = gz;
    }
}

// ---------------------------------------------------------------------------------------------
// composite (reference src/core/EMFusion.cpp:760-794)
// ---------------------------------------------------------------------------------------------
struct CompObj {
    const float* ray; size_t ray_pitch;
    const float* vert; size_t vert_pitch;
    const float* norm; size_t norm_pitch;
    const uint8_t* mask; size_t mask_pitch;
    int x0, y0, x1, y1;
    int id;
};
struct CompParams {
    CompObj o[EMF_MAX_VOLUMES];
    int n_obj;
    int w, h, boundary;
    const float* bg_ray; size_t bg_ray_pitch;
    const float* bg_vert; size_t bg_vert_pitch;
    const float* bg_norm; size_t bg_norm_pitch;
    const uint8_t* bg_mask; size_t bg_mask_pitch;
    float* ray; size_t ray_pitch;
    float* vert; size_t vert_pitch;
    float* norm; size_t norm_pitch;
    uint8_t* seg; size_t seg_pitch;
    int32_t* vis_count;
};

__global__ void __launch_bounds__(256) k_composite(const __grid_constant__ CompParams P) {
    __shared__ int s_cnt[EMF_MAX_VOLUMES];
    for (int k = threadIdx.x; k < P.n_obj; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < P.w && y < P.h) {
        float r = 0.0f;
        int win = -1;
        for (int k = 0; k < P.n_obj; ++k) {
            const CompObj& O = P.o[k];
            if (x < O.x0 || x >= O.x1 || y < O.y0 || y >= O.y1) continue;
            if (!O.mask[(size_t)y * O.mask_pitch + x]) continue;
            const float t = *((const float*)((const char*)O.ray + (size_t)y * O.ray_pitch) + x);
            if (r <= 0.0f || t < r) { r = t; win = k; }
        }
        int seg = 0;
        if (win >= 0) seg = P.o[win].id > 255 ? 255 : P.o[win].id;   // CV_8U saturate
        const bool bgm = P.bg_mask[(size_t)y * P.bg_mask_pitch + x] != 0;
        if (bgm) {
            const float bt = *((const float*)((const char*)P.bg_ray + (size_t)y * P.bg_ray_pitch) + x);
            if (fsub(r, bt) > 0.05f) seg = 0;
        }
        *((float*)((char*)P.ray + (size_t)y * P.ray_pitch) + x) = r;
        P.seg[(size_t)y * P.seg_pitch + x] = (uint8_t)seg;
        const float* vs; const float* ns;
        if (seg == 0) {
            vs = (const float*)((const char*)P.bg_vert + (size_t)y * P.bg_vert_pitch) + 3 * x;
            ns = (const float*)((const char*)P.bg_norm + (size_t)y * P.bg_norm_pitch) + 3 * x;
        } else {
            const CompObj& O = P.o[win];
            vs = (const float*)((const char*)O.vert + (size_t)y * O.vert_pitch) + 3 * x;
            ns = (const float*)((const char*)O.norm + (size_t)y * O.norm_pitch) + 3 * x;
        }
        float* vo = (float*)((char*)P.vert + (size_t)y * P.vert_pitch) + 3 * x;
        float* no = (float*)((char*)P.norm + (size_t)y * P.norm_pitch) + 3 * x;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
        if (seg != 0 || bgm) { v0 = vs[0]; v1 = vs[1]; v2 = vs[2]; n0 = ns[0]; n1 = ns[1]; n2 = ns[2]; }
        vo[0] = v0; vo[1] = v1; vo[2] = v2; no[0] = n0; no[1] = n1; no[2] = n2;
        if (seg != 0 && P.o[win].id == seg && x >= P.boundary && x < P.w - P.boundary && y >= P.boundary &&
            y < P.h - P.boundary)
            atomicAdd(&s_cnt[win], 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < P.n_obj; k += blockDim.x)
        if (s_cnt[k]) atomicAdd(P.vis_count + k, s_cnt[k]);
}

// ---------------------------------------------------------------------------------------------
// multi-GPU: merge of the per-rank pre-composites on the rank that owns the background.
// Every rank composites its own objects (k_composite against an empty background: list order is preserved inside a
// shard); here the per-rank winners are merged in (raylength, list index) order -- what the reference's sequential
// "strictly nearer, or first in the list" loop (src/core/EMFusion.cpp:760-771) computes over all objects -- followed by
// the background rule, the fill from the background and the visibility counts (:773-794).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxParts = 16;
struct MergeParams {
    const float* ray[kMaxParts]; const float* vert[kMaxParts]; const float* norm[kMaxParts]; const uint8_t* seg[kMaxParts];
    size_t ray_pitch[kMaxParts], vert_pitch[kMaxParts], norm_pitch[kMaxParts], seg_pitch[kMaxParts];
    int n_parts, n_obj, w, h, boundary;
    int16_t lut[256];          // segmentation id -> list index (-1: none)
    float* bg_ray; size_t bg_ray_pitch;
    float* bg_vert; size_t bg_vert_pitch;
    float* bg_norm; size_t bg_norm_pitch;
    uint8_t* bg_mask; size_t bg_mask_pitch;
    // replicated background: part p holds rows [p * band_rows, (p + 1) * band_rows) of the background's raycast
    int band_rows;             // 0: the bg_* images are complete inputs; > 0: they are assembled here from the bands
    const float* b_ray[kMaxParts]; const float* b_vert[kMaxParts]; const float* b_norm[kMaxParts]; const uint8_t* b_mask[kMaxParts];
    float* o_ray; size_t o_ray_pitch;
    float* o_vert; size_t o_vert_pitch;
    float* o_norm; size_t o_norm_pitch;
    uint8_t* o_seg; size_t o_seg_pitch;
    int32_t* vis_count;
};

__global__ void __launch_bounds__(256) k_composite_merge(const __grid_constant__ MergeParams P) {
    __shared__ int s_cnt[EMF_MAX_VOLUMES];
    for (int k = threadIdx.x; k < P.n_obj; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < P.w && y < P.h) {
        float r = 0.0f;
        int best = 1 << 30, win = -1, seg = 0;
        for (int p = 0; p < P.n_parts; ++p) {
            const int sg = P.seg[p][(size_t)y * P.seg_pitch[p] + x];
            if (!sg) continue;
            const int idx = P.lut[sg] < 0 ? (1 << 29) : P.lut[sg];
            const float t = *((const float*)((const char*)P.ray[p] + (size_t)y * P.ray_pitch[p]) + x);
            // sequential rule in list order == lexicographic minimum over (raylength, list index); a winner whose
            // raylength is <= 0 is replaced by any later object
            const bool take = win < 0 || (r <= 0.0f && idx > best) || t < r || (t == r && idx < best && !(r <= 0.0f));
            if (take) { r = t; best = idx; win = p; seg = sg; }
        }
        float* bgv = (float*)((char*)P.bg_vert + (size_t)y * P.bg_vert_pitch) + 3 * x;
        float* bgn = (float*)((char*)P.bg_norm + (size_t)y * P.bg_norm_pitch) + 3 * x;
        float* bgr = (float*)((char*)P.bg_ray + (size_t)y * P.bg_ray_pitch) + x;
        uint8_t* bgmp = P.bg_mask + (size_t)y * P.bg_mask_pitch + x;
        if (P.band_rows > 0) {   // assemble the background's raycast from the band of the rank that traced this row
            const int p = min(y / P.band_rows, P.n_parts - 1);
            const size_t i = (size_t)(y - p * P.band_rows) * P.w + x;
            const uint8_t m = P.b_mask[p][i];
            *bgmp = m; *bgr = P.b_ray[p][i];
            if (m) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { bgv[k] = P.b_vert[p][3 * i + k]; bgn[k] = P.b_norm[p][3 * i + k]; }
            }
        }
        const bool bgm = *bgmp != 0;
        if (bgm) {
            const float bt = *bgr;
            if (fsub(r, bt) > 0.05f) seg = 0;
        }
        *((float*)((char*)P.o_ray + (size_t)y * P.o_ray_pitch) + x) = r;
        P.o_seg[(size_t)y * P.o_seg_pitch + x] = (uint8_t)seg;
        const float* vs; const float* ns;
        if (seg == 0) {
            vs = bgv;
            ns = bgn;
        } else {
            vs = (const float*)((const char*)P.vert[win] + (size_t)y * P.vert_pitch[win]) + 3 * x;
            ns = (const float*)((const char*)P.norm[win] + (size_t)y * P.norm_pitch[win]) + 3 * x;
        }
        float* vo = (float*)((char*)P.o_vert + (size_t)y * P.o_vert_pitch) + 3 * x;
        float* no = (float*)((char*)P.o_norm + (size_t)y * P.o_norm_pitch) + 3 * x;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
        if (seg != 0 || bgm) { v0 = vs[0]; v1 = vs[1]; v2 = vs[2]; n0 = ns[0]; n1 = ns[1]; n2 = ns[2]; }
        vo[0] = v0; vo[1] = v1; vo[2] = v2; no[0] = n0; no[1] = n1; no[2] = n2;
        if (seg != 0 && P.lut[seg] >= 0 && x >= P.boundary && x < P.w - P.boundary && y >= P.boundary && y < P.h - P.boundary)
            atomicAdd(&s_cnt[P.lut[seg]], 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < P.n_obj; k += blockDim.x)
        if (s_cnt[k]) atomicAdd(P.vis_count + k, s_cnt[k]);
}

}  // namespace emfb

using namespace emfb;

extern "C" EMF_API int emf_raycast_tsdf(const float* tsdf, const float* grads, const float* weights, const float* fg_probs,
                                const emf_image* raylengths, const emf_image* vertices, const emf_image* normals,
                                const emf_image* mask, const emf_pose* T_co, const float K[9], const int res[3],
                                float voxel_size, float truncdist, int32_t* hit_voxel, emf_stream_t stream) {
    if (!T_co || !K || !res || !raylengths) return EMF_ERR_INVALID;
    emf_volume v = {};
    v.tsdf = (float*)tsdf; v.weights = (float*)weights; v.grads = grads; v.fg_probs = fg_probs;
    v.res[0] = res[0]; v.res[1] = res[1]; v.res[2] = res[2];
    v.voxel_size = voxel_size; v.truncdist = truncdist; v.id = 0;
    RayParams P;
    const int w = raylengths->width, h = raylengths->height;
    const int rc = fill_ray_vol(P.v[0], v, *T_co, raylengths, vertices, normals, mask, nullptr, w, h);
    if (rc != EMF_OK) return rc;
    P.v[0].first_block = 0;
    P.n_vol = 1; P.w = w; P.h = h;
    for (int k = 0; k < 9; ++k) P.K[k] = K[k];
    P.hit_voxel = hit_voxel; P.write_all = 0; P.stats = nullptr;
    const int blocks = P.v[0].tiles_x * ((h + kTileH - 1) / kTileH);
    k_raycast<false, false><<<blocks, kRayThreads, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}

extern "C" EMF_API int emf_raycast_volumes(int n_vol, const emf_volume* vols, const emf_pose* T_co, const float K[9],
                                   const int* rects, const emf_image* ray_out, const emf_image* vert_out,
                                   const emf_image* norm_out, const emf_image* mask_out, uint64_t* stats,
                                   emf_stream_t stream) {
    if (n_vol <= 0 || !vols || !T_co || !K || !ray_out || !vert_out || !norm_out || !mask_out) return EMF_ERR_INVALID;
    if (n_vol > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    RayParams P;
    const int w = ray_out[0].width, h = ray_out[0].height;
    int64_t blocks = 0;
    for (int i = 0; i < n_vol; ++i) {
        const int rc = fill_ray_vol(P.v[i], vols[i], T_co[i], &ray_out[i], &vert_out[i], &norm_out[i], &mask_out[i],
                                    rects ? rects + 4 * i : nullptr, w, h);
        if (rc != EMF_OK) return rc;
        P.v[i].first_block = (int)blocks;
        blocks += (int64_t)P.v[i].tiles_x * ((P.v[i].y1 - P.v[i].y0 + kTileH - 1) / kTileH);
    }
    if (blocks == 0) return EMF_OK;
    P.n_vol = n_vol; P.w = w; P.h = h;
    for (int k = 0; k < 9; ++k) P.K[k] = K[k];
    P.hit_voxel = nullptr; P.write_all = 1; P.stats = (unsigned long long*)stats;
    bool jump = false;
    for (int i = 0; i < n_vol; ++i) jump = jump || P.v[i].bmap != nullptr;
    const cudaStream_t cs = (cudaStream_t)stream;
    if (jump) {
        if (stats) k_raycast<true, true><<<(unsigned)blocks, kRayThreads, 0, cs>>>(P);
        else k_raycast<false, true><<<(unsigned)blocks, kRayThreads, 0, cs>>>(P);
    } else {
        if (stats) k_raycast<true, false><<<(unsigned)blocks, kRayThreads, 0, cs>>>(P);
        else k_raycast<false, false><<<(unsigned)blocks, kRayThreads, 0, cs>>>(P);
    }
    return launch_status();
}

extern "C" EMF_API int emf_compute_tsdf_grads(const float* tsdf, float* grads, const int res[3], emf_stream_t stream) {
    if (!tsdf || !grads || !res_ok(res)) return EMF_ERR_INVALID;
    const int64_t n = (int64_t)res[0] * res[1] * res[2];
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    k_grads<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(tsdf, grads, res[0], res[1], res[2]);
    return launch_status();
}

extern "C" EMF_API int emf_raycast_composite(int n_obj, const int* ids, const int* rects, const emf_image* obj_ray,
                                     const emf_image* obj_vert, const emf_image* obj_norm, const emf_image* obj_mask,
                                     const emf_image* bg_ray, const emf_image* bg_vert, const emf_image* bg_norm,
                                     const emf_image* bg_mask, int boundary, const emf_image* ray,
                                     const emf_image* vert, const emf_image* norm, const emf_image* seg,
                                     int32_t* vis_count, emf_stream_t stream) {
    if (n_obj < 0 || n_obj > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    if (!image_ok(bg_ray, 4) || !image_ok(bg_vert, 12) || !image_ok(bg_norm, 12) || !image_ok(bg_mask, 1) ||
        !image_ok(ray, 4) || !image_ok(vert, 12) || !image_ok(norm, 12) || !image_ok(seg, 1))
        return EMF_ERR_INVALID;
    if (n_obj > 0 && (!ids || !obj_ray || !obj_vert || !obj_norm || !obj_mask || !vis_count)) return EMF_ERR_INVALID;
    CompParams P;
    const int w = ray->width, h = ray->height;
    for (int k = 0; k < n_obj; ++k) {
        if (!image_ok(&obj_ray[k], 4) || !image_ok(&obj_vert[k], 12) || !image_ok(&obj_norm[k], 12) ||
            !image_ok(&obj_mask[k], 1))
            return EMF_ERR_INVALID;
        CompObj& o = P.o[k];
        o.ray = (const float*)obj_ray[k].ptr; o.ray_pitch = obj_ray[k].pitch;
        o.vert = (const float*)obj_vert[k].ptr; o.vert_pitch = obj_vert[k].pitch;
        o.norm = (const float*)obj_norm[k].ptr; o.norm_pitch = obj_norm[k].pitch;
        o.mask = (const uint8_t*)obj_mask[k].ptr; o.mask_pitch = obj_mask[k].pitch;
        o.x0 = 0; o.y0 = 0; o.x1 = w; o.y1 = h;
        if (rects) {
            o.x0 = rects[4 * k] < 0 ? 0 : rects[4 * k]; o.y0 = rects[4 * k + 1] < 0 ? 0 : rects[4 * k + 1];
            o.x1 = rects[4 * k + 2] > w ? w : rects[4 * k + 2]; o.y1 = rects[4 * k + 3] > h ? h : rects[4 * k + 3];
        }
        o.id = ids[k];
    }
    P.n_obj = n_obj; P.w = w; P.h = h; P.boundary = boundary;
    P.bg_ray = (const float*)bg_ray->ptr; P.bg_ray_pitch = bg_ray->pitch;
    P.bg_vert = (const float*)bg_vert->ptr; P.bg_vert_pitch = bg_vert->pitch;
    P.bg_norm = (const float*)bg_norm->ptr; P.bg_norm_pitch = bg_norm->pitch;
    P.bg_mask = (const uint8_t*)bg_mask->ptr; P.bg_mask_pitch = bg_mask->pitch;
    P.ray = (float*)ray->ptr; P.ray_pitch = ray->pitch;
    P.vert = (float*)vert->ptr; P.vert_pitch = vert->pitch;
    P.norm = (float*)norm->ptr; P.norm_pitch = norm->pitch;
    P.seg = (uint8_t*)seg->ptr; P.seg_pitch = seg->pitch;
    P.vis_count = vis_count;
    if (n_obj > 0) cudaMemsetAsync(vis_count, 0, sizeof(int32_t) * n_obj, (cudaStream_t)stream);
    const dim3 grid((w + 31) / 32, (h + 7) / 8);
    k_composite<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}

extern "C" EMF_API int emf_composite_merge(int n_parts, const emf_image* part_ray, const emf_image* part_vert,
                                   const emf_image* part_norm, const emf_image* part_seg, int n_obj, const int* ids,
                                   const emf_image* bg_ray, const emf_image* bg_vert, const emf_image* bg_norm,
                                   const emf_image* bg_mask, int boundary, const emf_image* ray, const emf_image* vert,
                                   const emf_image* norm, const emf_image* seg, int32_t* vis_count, int band_rows,
                                   const void* const* band_ray, const void* const* band_vert, const void* const* band_norm,
                                   const void* const* band_mask, emf_stream_t stream) {
    if (n_parts <= 0 || n_parts > kMaxParts || n_obj < 0 || n_obj > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    if (band_rows < 0 || (band_rows > 0 && (!band_ray || !band_vert || !band_norm || !band_mask))) return EMF_ERR_INVALID;
    if (!part_ray || !part_vert || !part_norm || !part_seg || (n_obj > 0 && (!ids || !vis_count))) return EMF_ERR_INVALID;
    if (!image_ok(bg_ray, 4) || !image_ok(bg_vert, 12) || !image_ok(bg_norm, 12) || !image_ok(bg_mask, 1) ||
        !image_ok(ray, 4) || !image_ok(vert, 12) || !image_ok(norm, 12) || !image_ok(seg, 1))
        return EMF_ERR_INVALID;
    MergeParams P;
    const int w = ray->width, h = ray->height;
    for (int p = 0; p < n_parts; ++p) {
        if (!image_ok(&part_ray[p], 4) || !image_ok(&part_vert[p], 12) || !image_ok(&part_norm[p], 12) ||
            !image_ok(&part_seg[p], 1) || part_ray[p].width != w || part_ray[p].height != h)
            return EMF_ERR_INVALID;
        P.ray[p] = (const float*)part_ray[p].ptr; P.ray_pitch[p] = part_ray[p].pitch;
        P.vert[p] = (const float*)part_vert[p].ptr; P.vert_pitch[p] = part_vert[p].pitch;
        P.norm[p] = (const float*)part_norm[p].ptr; P.norm_pitch[p] = part_norm[p].pitch;
        P.seg[p] = (const uint8_t*)part_seg[p].ptr; P.seg_pitch[p] = part_seg[p].pitch;
    }
    for (int k = 0; k < 256; ++k) P.lut[k] = -1;
    for (int k = n_obj - 1; k >= 0; --k) P.lut[ids[k] > 255 ? 255 : (ids[k] < 0 ? 0 : ids[k])] = (int16_t)k;   // first in list wins a shared id
    P.lut[0] = -1;
    P.n_parts = n_parts; P.n_obj = n_obj; P.w = w; P.h = h; P.boundary = boundary;
    P.bg_ray = (float*)bg_ray->ptr; P.bg_ray_pitch = bg_ray->pitch;
    P.bg_vert = (float*)bg_vert->ptr; P.bg_vert_pitch = bg_vert->pitch;
    P.bg_norm = (float*)bg_norm->ptr; P.bg_norm_pitch = bg_norm->pitch;
    P.bg_mask = (uint8_t*)bg_mask->ptr; P.bg_mask_pitch = bg_mask->pitch;
    P.band_rows = band_rows;
    for (int p = 0; p < n_parts && band_rows > 0; ++p) {
        if (!band_ray[p] || !band_vert[p] || !band_norm[p] || !band_mask[p]) return EMF_ERR_INVALID;
        P.b_ray[p] = (const float*)band_ray[p]; P.b_vert[p] = (const float*)band_vert[p];
        P.b_norm[p] = (const float*)band_norm[p]; P.b_mask[p] = (const uint8_t*)band_mask[p];
    }
    P.o_ray = (float*)ray->ptr; P.o_ray_pitch = ray->pitch;
    P.o_vert = (float*)vert->ptr; P.o_vert_pitch = vert->pitch;
    P.o_norm = (float*)norm->ptr; P.o_norm_pitch = norm->pitch;
    P.o_seg = (uint8_t*)seg->ptr; P.o_seg_pitch = seg->pitch;
    P.vis_count = vis_count;
    if (n_obj > 0) cudaMemsetAsync(vis_count, 0, sizeof(int32_t) * n_obj, (cudaStream_t)stream);
    const dim3 grid((w + 31) / 32, (h + 7) / 8);
    k_composite_merge<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}
