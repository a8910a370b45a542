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
#include <stdlib.h>

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
    int tiles_y;             // tile rows; > 0: blocks take the tiles ring by ring from the rim of the rectangle inwards (0: row-major)
    int first_block;
    int tube;                // 1: tube skipping (large volume, Rx % 4 == 0, 16-byte aligned)
    // ray-space certificate (k_ray_certify): bit (slab, tile) = every voxel a march sample of that 8 x 4 pixel tile can
    // touch while its base voxel lies in that slab holds exactly +1
    const uint32_t* cert;    // nullable: word [slab * cert_G + tile / 32], bit tile % 32
    int cert_G, cert_txc;    // words per slab; tiles per rect row
    int cert_axis, cert_sign, cert_L;   // slabs are ranges of 2^L voxels along this volume axis; every ray advances along it in this direction
    int cert_nw;             // words of slab bits that k_ray_certify wrote (the others read as 0)
};

struct RayParams {
    RayVol v[EMF_MAX_VOLUMES];
    int n_vol;
    int w, h;
    float K[9];
    int32_t* hit_voxel;      // optional (single-volume API)
    int write_all;           // 1: batched semantics (ray/mask written for every pixel of the rect)
    // longest-first schedule of the first volume's tiles (k_ray_schedule below); nullptr = row-major
    uint32_t* sched;             // [0] tile count the order is valid for [1] tiles placed before the other volumes' blocks
                                 // [4 ..) order[n_max], then cost[n_max] (pair iterations of the tile's longest ray, this launch)
    int sched_n0, sched_max;
    unsigned long long* timeline;   // diagnostics (EMF_RAY_TIMELINE=1: the stats pointer is a timeline buffer, 4 words per warp after 32; production kernel)
    int hist;                    // diagnostics (EMF_RAY_HIST=1, needs 32 counters): stats[8 + min(15, warp iterations / 32)]++, stats[24] = max
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

// (a real function: it is needed for a handful of samples per ray, and inlining it at every site of the march made the
//  kernel larger than the instruction cache -- "no instruction" became the top stall reason)
#ifndef EMF_RAY_WEIGHT_INLINE
#define EMF_RAY_WEIGHT_INLINE 1      // measured: 0.608 ms (inline) vs 0.653 ms (call) for the frame's raycast
#endif
#if EMF_RAY_WEIGHT_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
float trilinear_weight_fn(const float* __restrict__ weights, const float* __restrict__ fg_probs, int rx, int ry,
                                                  float vx, float vy, float vz) {
    const TriSetup s(vx, vy, vz, rx, ry);
    const int64_t b[4] = {s.base, s.base + rx, s.base + (int64_t)ry * rx, s.base + (int64_t)ry * rx + rx};
    float w[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { w[2 * k] = __ldg(weights + b[k]); w[2 * k + 1] = __ldg(weights + b[k] + 1); }
    if (fg_probs) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            w[2 * k] = (__ldg(fg_probs + b[k]) > 0.5f) ? w[2 * k] : 0.0f;
            w[2 * k + 1] = (__ldg(fg_probs + b[k] + 1) > 0.5f) ? w[2 * k + 1] : 0.0f;
        }
    }
    return s.combine(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ float trilinear_weight(const RayVol& V, float vx, float vy, float vz) {
    return trilinear_weight_fn(V.weights, V.fg_probs, V.rx, V.ry, vx, vy, vz);
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
// the guarded divisions (a numerator that is tiny or exactly zero, or a ray outside ConstDiv's window): rare, out of line
#ifndef EMF_RAY_GDIV_INLINE
#define EMF_RAY_GDIV_INLINE 0
#endif
#if EMF_RAY_GDIV_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
float3 guarded_div3(float nx, float ny, float nz, float s) {
    const ConstDiv div_s(s);
    return make_float3(div_s(nx), div_s(ny), div_s(nz));
}
__device__ __forceinline__ void samp_pos(Samp& q, float t, const RayConst& c, const ConstDiv& div_s, const RayVol& V) {
    const float nx = ffma(c.dx, t, c.ox), ny = ffma(c.dy, t, c.oy), nz = ffma(c.dz, t, c.oz);
    if (c.sane && fminf(fminf(fabsf(nx), fabsf(ny)), fabsf(nz)) > 5.5e-20f) {
        q.vx = fadd(c.hxh, div_s.fast(nx)); q.vy = fadd(c.hyh, div_s.fast(ny)); q.vz = fadd(c.hzh, div_s.fast(nz));
    } else {
        const float3 d = guarded_div3(nx, ny, nz, c.s);
        q.vx = fadd(c.hxh, d.x); q.vy = fadd(c.hyh, d.y); q.vz = fadd(c.hzh, d.z);
    }
    q.in = !out_of_thr(q.vx, q.vy, q.vz, V.thr2);
}
__device__ __forceinline__ void samp_load(Samp& q, const RayVol& V, int rx, int plane) {
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
__device__ __forceinline__ void samp_fetch(Samp& q, float t, const RayConst& c, const ConstDiv& div_s, const RayVol& V, int rx, int plane) {
    samp_pos(q, t, c, div_s, V);
    samp_load(q, V, rx, plane);
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
    if (STATS) { ++st[0]; if (r.f == 1.0f && fn == 1.0f) ++st[6]; }
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
// The rare outcomes of a sample -- a sign change against the ray's previous value -- out of line: the back-face test
// (TSDF.cu:532) and the front-face refinement (TSDF.cu:540-566) with their weight gathers.  The sample's position is
// recomputed from its ray parameter (a pure function of it).  Returns true when the ray is finished; r.f is updated as the
// reference's loop would (left alone on `continue`).
#ifndef EMF_RAY_EVENT_INLINE
#define EMF_RAY_EVENT_INLINE 1
#endif
template <bool STATS>
#if EMF_RAY_EVENT_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
bool pair_event(Ray& r, const RayConst& c, const ConstDiv& div_s, const RayVol& V, float fn, unsigned long long* st) {
    Samp q;
    samp_pos(q, r.tcur, c, div_s, V);
    if (r.f < 0.0f) {                       // (fn > 0) back face
        if (STATS) ++st[3];
        if (trilinear_weight(V, q.vx, q.vy, q.vz) > 0.0f) return true;
    } else {                                // f > 0, fn < 0: front face (the step already follows fn)
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

// one general march step out of line (ray ends, samples outside the march bounds, guarded divisions: a few per ray)
template <bool STATS>
__device__ __noinline__ bool march_step_slow(Ray& r, const RayConst& c, const RayVol& V, unsigned long long* st) {
    const ConstDiv div_s(c.s);
    return march_step<STATS>(r, c, div_s, V, V.rx, V.rx * V.ry, st);
}

// The march, two samples per iteration.  Fast path: both samples exist, both positions take the unguarded division, both lie
// inside the march bounds (one combined test each) -- then 16 gathers are issued together, the second sample speculatively
// (it is discarded if the first one ends the ray, keeps f, or changes the step).  Everything else goes through ONE general
// step and the loop goes on.  Same samples, same arithmetic, same order as march_step.  true = the ray is finished.
template <bool STATS>
__device__ __forceinline__ bool march_pairs(Ray& r, const RayConst& c, const ConstDiv& div_s, const RayVol& V, int rx, int plane,
                                            unsigned long long* st, int max_iters = 0x7fffffff, int* iters = nullptr) {
    const float* __restrict__ vol = V.tsdf;
    // (the march state in scalars: the out-of-line general step takes a COPY of the ray, so that nothing of the hot loop's
    //  state has its address taken and ends up in local memory)
    float tcur = r.tcur, step = r.step, f = r.f;
    bool done = false;
    for (int it = 0; it < max_iters; ++it) {
        if (iters) *iters = it;
        if (STATS) { const unsigned am = __activemask(); if ((int)(threadIdx.x & 31) == __ffs(am) - 1) { ++st[4]; st[5] += __popc(am); } }
        const float t1 = fadd(tcur, step);
        const float t2 = fadd(t1, step);
        const float n1x = ffma(c.dx, t1, c.ox), n1y = ffma(c.dy, t1, c.oy), n1z = ffma(c.dz, t1, c.oz);
        const float n2x = ffma(c.dx, t2, c.ox), n2y = ffma(c.dy, t2, c.oy), n2z = ffma(c.dz, t2, c.oz);
        bool fast = c.sane && t2 <= c.tmax &&
                    fminf(fminf(fminf(fabsf(n1x), fabsf(n1y)), fabsf(n1z)), fminf(fminf(fabsf(n2x), fabsf(n2y)), fabsf(n2z))) > 5.5e-20f;
        const float v1x = fadd(c.hxh, div_s.fast(n1x)), v1y = fadd(c.hyh, div_s.fast(n1y)), v1z = fadd(c.hzh, div_s.fast(n1z));
        const float v2x = fadd(c.hxh, div_s.fast(n2x)), v2y = fadd(c.hyh, div_s.fast(n2y)), v2z = fadd(c.hzh, div_s.fast(n2z));
        fast = fast && !out_of_thr(v1x, v1y, v1z, V.thr2) && !out_of_thr(v2x, v2y, v2z, V.thr2);
        if (!fast) {
            Ray tmp = r;
            tmp.tcur = tcur; tmp.step = step; tmp.f = f;
            done = march_step_slow<STATS>(tmp, c, V, st);
            tcur = tmp.tcur; step = tmp.step; f = tmp.f;
            if (done) { r = tmp; break; }
            continue;
        }
        // ---- base voxels, fractions, 16 gathers
        const int l1x = __float2int_rz(v1x), l1y = __float2int_rz(v1y), l1z = __float2int_rz(v1z);
        const int l2x = __float2int_rz(v2x), l2y = __float2int_rz(v2y), l2z = __float2int_rz(v2z);
        const float* p1 = vol + (uint32_t)(l1z * plane + l1y * rx + l1x);
        const float* p2 = vol + (uint32_t)(l2z * plane + l2y * rx + l2x);
        const float a000 = __ldg(p1), a001 = __ldg(p1 + 1), a010 = __ldg(p1 + rx), a011 = __ldg(p1 + rx + 1);
        const float a100 = __ldg(p1 + plane), a101 = __ldg(p1 + plane + 1), a110 = __ldg(p1 + plane + rx), a111 = __ldg(p1 + plane + rx + 1);
        const float b000 = __ldg(p2), b001 = __ldg(p2 + 1), b010 = __ldg(p2 + rx), b011 = __ldg(p2 + rx + 1);
        const float b100 = __ldg(p2 + plane), b101 = __ldg(p2 + plane + 1), b110 = __ldg(p2 + plane + rx), b111 = __ldg(p2 + plane + rx + 1);
        float fn;
        {
            const float ax = fsub(v1x, (float)l1x), ay = fsub(v1y, (float)l1y), az = fsub(v1z, (float)l1z);
            const float bx = fsub(1.0f, ax), by = fsub(1.0f, ay), bz = fsub(1.0f, az);
            fn = lerp1(bz, lerp1(by, lerp1(bx, a000, ax, a001), ay, lerp1(bx, a010, ax, a011)), az,
                       lerp1(by, lerp1(bx, a100, ax, a101), ay, lerp1(bx, a110, ax, a111)));
        }
        // ---- first sample (the order of march_step: back-face test, step update, front-face test, f = fn)
        const float step0 = step;
        tcur = t1;
        if (STATS) { ++st[0]; if (f == 1.0f && fn == 1.0f) ++st[6]; }
        if ((f < 0.0f && fn > 0.0f) || (f > 0.0f && fn < 0.0f)) {
            const bool back = f < 0.0f;
            if (!back) { if (fabsf(fn) < 1.0f) step = c.s; if (fabsf(fn) < 0.8f) step = c.half_s; }
            r.tcur = tcur; r.step = step; r.f = f;
            done = pair_event<STATS>(r, c, div_s, V, fn, st);
            f = r.f;
            if (done) break;
            if (back) { if (fabsf(fn) < 1.0f) step = c.s; if (fabsf(fn) < 0.8f) step = c.half_s; }
        } else {
            if (fabsf(fn) < 1.0f) step = c.s;
            if (fabsf(fn) < 0.8f) step = c.half_s;
            f = fn;
        }
        if (step != step0) continue;                 // the speculation failed: the second sample is somewhere else
        // ---- second sample
        {
            const float ax = fsub(v2x, (float)l2x), ay = fsub(v2y, (float)l2y), az = fsub(v2z, (float)l2z);
            const float bx = fsub(1.0f, ax), by = fsub(1.0f, ay), bz = fsub(1.0f, az);
            fn = lerp1(bz, lerp1(by, lerp1(bx, b000, ax, b001), ay, lerp1(bx, b010, ax, b011)), az,
                       lerp1(by, lerp1(bx, b100, ax, b101), ay, lerp1(bx, b110, ax, b111)));
        }
        tcur = t2;
        if (STATS) { ++st[0]; if (f == 1.0f && fn == 1.0f) ++st[6]; }
        if ((f < 0.0f && fn > 0.0f) || (f > 0.0f && fn < 0.0f)) {
            const bool back = f < 0.0f;
            if (!back) { if (fabsf(fn) < 1.0f) step = c.s; if (fabsf(fn) < 0.8f) step = c.half_s; }
            r.tcur = tcur; r.step = step; r.f = f;
            done = pair_event<STATS>(r, c, div_s, V, fn, st);
            f = r.f;
            if (done) break;
            if (back) { if (fabsf(fn) < 1.0f) step = c.s; if (fabsf(fn) < 0.8f) step = c.half_s; }
        } else {
            if (fabsf(fn) < 1.0f) step = c.s;
            if (fabsf(fn) < 0.8f) step = c.half_s;
            f = fn;
        }
    }
    r.tcur = tcur; r.step = step; r.f = f;
    return done;
}

// ---- certified skipping (k_ray_certify below).  While a ray carries exactly +1, a march sample whose eight corners all
// hold exactly +1 returns exactly +1 again ((1 - a) + a rounds to 1 for every fraction a) and changes nothing of the march
// state but the ray parameter.  The certificate says, per 8 x 4 pixel tile (= one warp) and per slab of 2^L voxels along
// the volume axis the camera looks along, that EVERY voxel any sample of any ray of the tile can touch while its base
// voxel is in that slab holds +1.  The ray position along that axis is monotone in the ray parameter (every rounded
// operation of ray_position is), so a run of certified slabs is skipped in one go: the n additions of the ray parameter
// in closed form (inside one binade RN(t + step) = t + inc with a constant inc: seq_add.h), then the position of the LAST
// skipped sample is evaluated exactly as the sampler would and checked to lie in the run -- the samples before it then
// do, too.  Bit-identical to marching through.
// The warp works in phases, decided with ballots (its rays share the tile's bits): rays whose next sample is certified wait
// until every ray of the warp is in that state (or finished), then all skip together; otherwise the others take samples.
// Skipping lane by lane instead makes nearly every iteration of the warp pay for both paths (measured: no faster than marching).
constexpr int kCertWords = 8;                      // 32-bit words of slab bits per tile
constexpr int kCertSlabs = 32 * kCertWords;

// Advances t by up to `total` executions of `t = t + step` (round to nearest even each), returns the new t and how many
// were made in `made` (>= 1 when total >= 1).  Inside one binade every float is a multiple of u = 2^(e-23), so
// RN(t + step) = t + inc with a constant inc once two consecutive additions agree (seq_add.h); the number of additions that
// stay below 2^(e+1) is an exact integer quotient of bit-pattern differences.  A few rounds cover a few binades.
__device__ __forceinline__ float cert_hop(float t, float step, int total, int& made) {
    made = 0;
#pragma unroll 1
    for (int round = 0; round < 4 && made < total; ++round) {
        const float t1 = __fadd_rn(t, step);
        ++made;
        const uint32_t u0 = __float_as_uint(t), u1 = __float_as_uint(t1);
        const uint32_t b1 = u1 >> 23;
        const int left = total - made;
        if (left < 1 || (u0 >> 23) != b1 || b1 == 0u || b1 >= 254u || !(step > 0.0f) || !(step <= t)) { t = t1; continue; }
        const float t2 = __fadd_rn(t1, step);
        const uint32_t u2 = __float_as_uint(t2);
        const uint32_t B = u1 - u0;                                  // inc in ulps of the binade
        if ((u2 >> 23) != b1 || u2 - u1 != B || B == 0u) { t = t1; continue; }
        const uint32_t A = ((b1 + 1u) << 23) - u1 - 1u;              // additions k after t1 stay in the binade iff k * B <= A
        int k = (int)__fdividef((float)A, (float)B);                 // A < 2^23: both exact as floats; quotient to an ulp
        if ((uint64_t)(uint32_t)k * B > A) --k;
        else if ((uint64_t)(uint32_t)(k + 1) * B <= A) ++k;
        k = min(k, left);
        if (k < 1) { t = t1; continue; }
        t = __uint_as_float(u1 + (uint32_t)k * B);                    // = t1 + k * inc, exactly
        made += k;
    }
    return t;
}

#ifndef EMF_RAY_TUBE
#define EMF_RAY_TUBE 0
#endif
#if EMF_RAY_TUBE
// ---- tube skipping (large volumes; compiled only with -DEMF_RAY_TUBE=1 -- see DESIGN.md section 4.3 for why).
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

#endif   // EMF_RAY_TUBE

// Everything of the reference kernel before its march loop (TSDF.cu:477-521): the ray, the slab test, the coarse skip and
// the first sample.  false = the ray has nothing to march (misses the box / the foreground box, or never gets inside).
// the ray and its interval [t_first, c.tmax] inside the volume (TSDF.cu:477-503); false = nothing to march
__device__ __forceinline__ bool ray_consts(const RayVol& V, const float* K, int x, int y, float old, RayConst& c, float& t_first) {
    c.s = V.voxel;
    const float s = c.s;
    const float ux = fdiv(fsub((float)x, K[2]), K[0]);
    const float uy = fdiv(fsub((float)y, K[5]), K[4]);
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
    t_first = fadd(s, tin);
    c.tmax = fsub(tout, s);
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
    c.sane = div_s.ok && fabsf(c.ox) + fabsf(c.oy) + fabsf(c.oz) + fabsf(t_first) + fabsf(c.tmax) + V.trunc < 1.0e18f;
    return !(culled || t_first >= c.tmax);
}
__device__ __forceinline__ bool ray_begin(const RayVol& V, const float* K, int x, int y, float old, Ray& r, RayConst& c) {
    r.out_t = 0.0f; r.hit = false;
    r.hvx = r.hvy = r.hvz = r.hmx = r.hmy = r.hmz = 0.f;
    r.step = V.trunc;
    r.vx = r.vy = r.vz = r.f = 0.f;
    r.tcur = 0.f;
    if (!ray_consts(V, K, x, y, old, c, r.tcur)) return false;
    const float s = c.s;
    const ConstDiv div_s(s);
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
        return true;
    }
    return false;
}

// MODE 0: every lane marches its own ray, two samples in flight (objects; volumes without a certificate).
// (certified skipping -- the volume carries RayVol::cert -- has its own kernel: k_raycast_cert)
// MODE 2: the warp stays in lockstep so that jumps through certified constant regions can be agreed on with warp
//         collectives (needs emf_volume::brick_map on at least one volume of the launch).
// One loop per kernel: with all of them inlined into one kernel the code outgrew the instruction cache.
#ifndef EMF_RAY_PAIR
#define EMF_RAY_PAIR 1
#endif
#ifndef EMF_RAY_TUBE
#define EMF_RAY_TUBE 0
#endif
#ifndef EMF_CERT_L
#define EMF_CERT_L 2      // log2 of the smallest slab thickness (voxels)
#endif
#ifndef EMF_RAY_MINB
#define EMF_RAY_MINB 7      // 72 registers, no spills, 28 warps per SM: 1.114 ms per frame; 8 (64 registers, 36 B of spills) 1.128; 6 1.126
#endif
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
template <bool STATS, int MODE>
__global__ void __launch_bounds__(kRayThreads, EMF_RAY_MINB) k_raycast(const __grid_constant__ RayParams P) {
    constexpr bool JUMP = MODE == 2;
    unsigned long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // [4], [5]: warp iterations of the march loop and the lanes active in them
    const unsigned long long t_begin = ((STATS && P.hist == 2) || P.timeline) ? global_ns() : 0ull;
    int lo = 0, hi = P.n_vol - 1;
    int b = blockIdx.x;
    const bool sched = MODE == 0 && P.sched != nullptr;
    if (sched) {
        // The launch ends with its longest dependent chains: tiles whose rays run along the rim of the frustum or a shadow seam
        // take 3-4 x the median.  The first volume's tiles are therefore taken longest first -- by the cost the previous
        // launch recorded (frames are coherent; the order changes no result) --; the other volumes' (short) blocks follow the
        // first sched[1] of them (default: all -- measured on the bench stream: objects after the background 1.113 ms/frame,
        // after its longest quarter 1.163, row-major 1.137; a young model with long seams gains more: 0.80 -> 0.59 ms).
        const int n0 = P.sched_n0, rest = (int)gridDim.x - n0;
        const bool ok = (int)P.sched[0] == n0;
        const int nf = ok ? min((int)P.sched[1], n0) : n0;
        int t0;
        if (b < nf) t0 = b;
        else if (b < nf + rest) { t0 = -1; b = n0 + (b - nf); }
        else t0 = b - rest;
        if (t0 >= 0) b = ok ? (int)P.sched[4 + t0] : t0;
    }
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (P.v[mid].first_block <= b) lo = mid; else hi = mid - 1;
    }
    const RayVol& V = P.v[lo];
    const int lb = b - V.first_block;
    int ty = lb / V.tiles_x, tx = lb - ty * V.tiles_x;
    // warp = kWarpW x kWarpH pixel patch inside the CTA tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xr = V.x0 + tx * kTileW + (warp & 1) * kWarpW + (lane % kWarpW);
    const int yr = V.y0 + ty * kTileH + (warp >> 1) * kWarpH + (lane / kWarpW);
    const bool valid = xr < V.x1 && yr < V.y1;
    if (MODE == 0 && !valid && !(EMF_RAY_TUBE && V.tube)) return;   // (the other modes use warp collectives: idle lanes stay)
    const int x = min(xr, V.x1 - 1), y = min(yr, V.y1 - 1);   // (JUMP: lanes outside the rectangle idle in the loop)

    float* ray_px = (float*)((char*)V.ray + (size_t)y * V.ray_pitch) + x;
    uint8_t* mask_px = V.mask + (size_t)y * V.mask_pitch + x;
    Ray r;
    RayConst c;
    const float old = (P.write_all || !valid) ? 0.0f : *ray_px;   // in/out far clip (TSDF.cu:496-500)
    bool done = !ray_begin(V, P.K, x, y, old, r, c) || !valid;
    const ConstDiv div_s(c.s);
    const int rx = V.rx, plane = V.rx * V.ry;
    if (MODE == 0) {
#if EMF_RAY_TUBE
        if (V.tube) {
            march_tube<STATS>(r, c, div_s, V, rx, plane, done, st);
            if (!valid) return;
        } else
#endif
        if (!done) {
#if EMF_RAY_PAIR
            int iters = 0;
            march_pairs<STATS>(r, c, div_s, V, rx, plane, st, 0x7fffffff, sched ? &iters : nullptr);
            if (sched && lo == 0) {
                const unsigned am = __activemask();
                const int m = __reduce_max_sync(am, iters);
                if ((int)(threadIdx.x & 31) == __ffs(am) - 1) atomicMax(P.sched + 4 + P.sched_max + lb, (uint32_t)m);
            }
#else
            while (!march_step<STATS>(r, c, div_s, V, rx, plane, st)) {}
#endif
        }
    } else {
        constexpr unsigned kFull = 0xffffffffu;
        const uint8_t* __restrict__ bmap = V.bmap;
        const float frx = (float)V.rx, fry = (float)V.ry, frz = (float)V.rz;
        // jump geometry: voxels advanced per metre of ray parameter along each axis (and the largest of them)
        const float s = c.s;
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
    if (!STATS && P.timeline) {
        const unsigned am = __activemask();
        if ((int)(threadIdx.x & 31) == __ffs(am) - 1) {
            unsigned long long* rec = P.timeline + 32 + 4 * ((size_t)blockIdx.x * (kRayThreads / 32) + (threadIdx.x >> 5));
            rec[0] = t_begin; rec[1] = global_ns(); rec[2] = 0; rec[3] = (unsigned)lo;
        }
    }
    if (STATS && P.stats) {
#pragma unroll
        for (int k = 0; k < 8; ++k) if (st[k]) atomicAdd(P.stats + k, st[k]);
        if (P.hist && V.cert && (threadIdx.x & 31) == 0) {
            atomicAdd(P.stats + 8 + min(15ull, st[4] / 32ull), 1ull);
            atomicMax(P.stats + 24, st[4]);
        }
        if (P.hist == 2 && (threadIdx.x & 31) == 0) {   // timeline: 4 words per warp after the 32 counters
            unsigned smid;
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            unsigned long long* rec = P.stats + 32 + 4 * ((size_t)blockIdx.x * (kRayThreads / 32) + (threadIdx.x >> 5));
            rec[0] = t_begin; rec[1] = global_ns(); rec[2] = st[4]; rec[3] = ((unsigned long long)smid << 32) | (unsigned)lo;
        }
    }
}

// STABLE counting sort of the first volume's tiles by the cost the launch just recorded, longest class first, for the NEXT
// launch.  Eight coarse classes (64 pair iterations each) and row-major order inside a class: tiles that run at the same
// time stay neighbours in the image and share their voxels in L1 / L2.  One CTA; a thread owns a contiguous chunk of tiles.
constexpr int kSchedClasses = 8;
__device__ __forceinline__ int sched_class(uint32_t cost) { return kSchedClasses - 1 - (int)min((uint32_t)(kSchedClasses - 1), cost >> 6); }
__global__ void __launch_bounds__(1024) k_ray_schedule(uint32_t* __restrict__ sched, int n, int n_max, int front_pct) {
    __shared__ int warp_tot[kSchedClasses][32];
    __shared__ int class_base[kSchedClasses];
    uint32_t* order = sched + 4;
    uint32_t* cost = sched + 4 + n_max;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int chunk = (n + (int)blockDim.x - 1) / (int)blockDim.x;
    const int i0 = min(n, tid * chunk), i1 = min(n, i0 + chunk);
    int cnt[kSchedClasses];
#pragma unroll
    for (int k = 0; k < kSchedClasses; ++k) cnt[k] = 0;
    for (int i = i0; i < i1; ++i) {
        const int c = sched_class(cost[i]);
#pragma unroll
        for (int k = 0; k < kSchedClasses; ++k) cnt[k] += (c == k);
    }
    int pre[kSchedClasses];          // exclusive prefix of this thread inside its warp, per class
#pragma unroll
    for (int k = 0; k < kSchedClasses; ++k) {
        int v = cnt[k];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += u; }
        pre[k] = v - cnt[k];
        if (lane == 31) warp_tot[k][wid] = v;
    }
    __syncthreads();
    if (wid < kSchedClasses) {       // warp k scans the 32 warp totals of class k
        const int t = warp_tot[wid][lane];
        int v = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += u; }
        warp_tot[wid][lane] = v - t;
        if (lane == 31) class_base[wid] = v;          // (the class total, turned into a base below)
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int k = 0; k < kSchedClasses; ++k) { const int t = class_base[k]; class_base[k] = acc; acc += t; }
        sched[0] = (uint32_t)n;
        sched[1] = (uint32_t)(((long long)n * front_pct) / 100);
    }
    __syncthreads();
    int pos[kSchedClasses];
#pragma unroll
    for (int k = 0; k < kSchedClasses; ++k) pos[k] = class_base[k] + warp_tot[k][wid] + pre[k];
    for (int i = i0; i < i1; ++i) {
        const int c = sched_class(cost[i]);
#pragma unroll
        for (int k = 0; k < kSchedClasses; ++k) if (c == k) order[pos[k]++] = (uint32_t)i;
        cost[i] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// k_raycast_cert: the certified march of one volume (the background), FOUR LANES PER RAY.
//
// What bounds the background's raycast is not its number of samples but its longest DEPENDENT CHAIN: rays along the rim of
// the frustum and along the shadow volumes of objects run through seams of observed / never observed / occluded space for
// up to ~1200 samples that are no no-ops, and a sample is position -> 8 gathers -> 7 lerps -> decision.  One ray per
// lane marches that chain two samples at a time (0.7 us per pair on an otherwise idle SM: the whole launch waits for it).
// Here the four lanes of a group take the ray's next EIGHT samples at once, two each (lane j: samples 2j and 2j + 1, at t
// after 2j + 1 / 2j + 2 additions of the step), under the assumption that the step does not change.  The first sample
// that changes the march state otherwise than by f = fn -- a different step, a sign change that ends the ray or keeps f,
// a sample outside the march bounds, the end of the ray -- is the "stop": the samples before it are accepted wholesale
// (state after them: t and f of the last one), the stop sample is resolved by the sequential rule by all lanes of the
// group in step (they hold the same state), what lies behind it is discarded.  Same samples, same arithmetic, same order
// as the reference loop (TSDF.cu:523-572): bit-identical, the chain per sample four times shorter, and as many warp
// instructions per accepted sample as with one ray per lane.
// A CTA is one 8 x 4 pixel tile of the certificate (k_ray_certify below), a warp one row of it.  Skipping: see cert_hop
// above -- the warp works in phases decided with ballots: rays whose next sample is certified wait until every ray of
// the warp is in that state (or finished), then all skip together.  Tiles are handed out rim first.
// ---------------------------------------------------------------------------------------------
struct CertRayParams {
    RayVol v;
    float K[9];
    int w, h;
    int hist;
    unsigned long long* stats;
};


struct RayEvent { int code; float ts, sx, sy, sz; };   // 0: go on, f = fn; 1: finished (back face); 2: go on, f unchanged; 3: hit

// a sign change between the previous sample f and this one fn (TSDF.cu:532, 540-566); `step` is the step AFTER its update
__device__ __noinline__ RayEvent ray_event(const RayVol& V, float f, float fn, float tcur, float step, float vx, float vy, float vz,
                                           float dx, float dy, float dz) {
    RayEvent e;
    e.code = 0; e.ts = e.sx = e.sy = e.sz = 0.0f;
    if (f < 0.0f && fn > 0.0f) {     // back face: ends the ray if the weight there is positive
        if (trilinear_weight_fn(V.weights, V.fg_probs, V.rx, V.ry, vx, vy, vz) > 0.0f) e.code = 1;
        return e;
    }
    const ConstDiv div_s(V.voxel);
    const float hxh = fmul((float)(V.rx - 1), 0.5f), hyh = fmul((float)(V.ry - 1), 0.5f), hzh = fmul((float)(V.rz - 1), 0.5f);
    const float ts = fsub(tcur, fdiv(fmul(f, step), fsub(fn, f)));
    const float sx = fadd(hxh, div_s(fadd(V.t[0], fmul(dx, ts))));
    const float sy = fadd(hyh, div_s(fadd(V.t[1], fmul(dy, ts))));
    const float sz = fadd(hzh, div_s(fadd(V.t[2], fmul(dz, ts))));
    if (out_of_thr(sx, sy, sz, V.thr2)) { e.code = 2; return e; }     // reference `continue`: f keeps its old value
    if (trilinear_weight_fn(V.weights, V.fg_probs, V.rx, V.ry, sx, sy, sz) > 0.0f) {
        e.code = 3; e.ts = ts; e.sx = sx; e.sy = sy; e.sz = sz;
    }
    return e;
}

// vertex, normal, ray length and mask of a hit (TSDF.cu:552-566)
__device__ __noinline__ void ray_emit_hit(const RayVol& V, int x, int y, float ts, float sx, float sy, float sz, float dx, float dy, float dz,
                                          int32_t* hit_voxel, int w) {
    float g[3];
    trilinear_grad(V, sx, sy, sz, g);
    const float mx = fmul(dx, ts), my = fmul(dy, ts), mz = fmul(dz, ts);
    float* vp = (float*)((char*)V.vert + (size_t)y * V.vert_pitch) + 3 * x;
    float* np = (float*)((char*)V.norm + (size_t)y * V.norm_pitch) + 3 * x;
    // transpose(rot_CO) * (t* dir)
    vp[0] = dot_yxz(V.R[0], V.R[3], V.R[6], mx, my, mz);
    vp[1] = dot_yxz(V.R[1], V.R[4], V.R[7], mx, my, mz);
    vp[2] = dot_yxz(V.R[2], V.R[5], V.R[8], mx, my, mz);
    const float gn = norm3(g[0], g[1], g[2]);
    const float nx = fdiv(g[0], gn), ny = fdiv(g[1], gn), nz = fdiv(g[2], gn);
    np[0] = dot_yxz(V.R[0], V.R[3], V.R[6], nx, ny, nz);
    np[1] = dot_yxz(V.R[1], V.R[4], V.R[7], nx, ny, nz);
    np[2] = dot_yxz(V.R[2], V.R[5], V.R[8], nx, ny, nz);
    *((float*)((char*)V.ray + (size_t)y * V.ray_pitch) + x) = ts;
    V.mask[(size_t)y * V.mask_pitch + x] = 1;
    if (hit_voxel) {
        int32_t* hv = hit_voxel + 3 * ((size_t)y * w + x);
        hv[0] = __float2int_rz(sx); hv[1] = __float2int_rz(sy); hv[2] = __float2int_rz(sz);
    }
}

constexpr int kGrpLanes = 4;         // lanes per ray
constexpr int kGrpSamples = 2 * kGrpLanes;   // samples per ray and iteration
#ifndef EMF_CERT_MINB
#define EMF_CERT_MINB 6
#endif
template <bool STATS>
__global__ void __launch_bounds__(kRayThreads, EMF_CERT_MINB) k_raycast_cert(const __grid_constant__ CertRayParams P) {
    constexpr unsigned kFull = 0xffffffffu;
    unsigned long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned long long t_begin = (STATS && P.hist == 2) ? global_ns() : 0ull;
    const RayVol& V = P.v;
    // the CTA's tile (8 x 4 pixels), rim of the rectangle first
    const int n_tx = V.cert_txc, n_ty = 2 * V.tiles_y;
    int ty, tx;
    {
        int r = 0, rest = blockIdx.x, wr = n_tx, hr = n_ty;
        for (;;) {
            const int cnt = hr == 1 ? wr : (wr == 1 ? hr : 2 * wr + 2 * hr - 4);
            if (rest < cnt || wr <= 2 || hr <= 2) break;
            rest -= cnt; ++r; wr -= 2; hr -= 2;
        }
        if (wr <= 2 || hr <= 2) { ty = r + rest / wr; tx = r + rest - (rest / wr) * wr; }      // the core that is left: row-major
        else if (rest < wr) { ty = r; tx = r + rest; }
        else if (rest < 2 * wr) { ty = r + hr - 1; tx = r + rest - wr; }
        else if (rest < 2 * wr + hr - 2) { tx = r; ty = r + 1 + rest - 2 * wr; }
        else { tx = r + wr - 1; ty = r + 1 + rest - 2 * wr - (hr - 2); }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane & (kGrpLanes - 1), gl0 = lane & ~(kGrpLanes - 1);     // lane inside the group, first lane of the group
    const int xr = V.x0 + tx * kWarpW + (lane / kGrpLanes);
    const int yr = V.y0 + ty * kWarpH + warp;
    const bool valid = xr < V.x1 && yr < V.y1;
    __shared__ uint32_t s_cert[2 * kCertWords];   // [0, kCertWords): all +1 ; [kCertWords, 2 kCertWords): all 0
    if (warp == 0 && !V.cert) {          // (EMF_RAY_WIDE: the four-lane march without a certificate -- nothing is ever skipped)
        if (lane < 2 * kCertWords) s_cert[lane] = 0u;
    } else if (warp == 0) {
        // the tile's slab bits, one word per 32 slabs (lane = slab of the word, bit = the tile inside its group of 32)
        const int ti = ty * V.cert_txc + tx;
        const int g = ti >> 5, tb = ti & 31;
#pragma unroll
        for (int k = 0; k < 2 * kCertWords; ++k) {
            const int kk = k & (kCertWords - 1);
            const uint32_t* tab = V.cert + (k < kCertWords ? (size_t)0 : (size_t)kCertSlabs * V.cert_G);
            const uint32_t wd = kk < V.cert_nw ? __ldg(tab + (size_t)(32 * kk + lane) * V.cert_G + g) : 0u;
            const uint32_t col = __ballot_sync(kFull, (wd >> tb) & 1u);
            if (lane == 0) s_cert[k] = col;
        }
    }
    __syncthreads();
    const uint32_t* __restrict__ bits = s_cert;
    const int x = min(xr, V.x1 - 1), y = min(yr, V.y1 - 1);   // (lanes outside the rectangle idle in the loop)
    Ray r;
    RayConst c;
    bool done = !ray_begin(V, P.K, x, y, 0.0f, r, c) || !valid;
    bool hit = false;
    const ConstDiv div_s(c.s);
    const int rx = V.rx, plane = V.rx * V.ry;
    const int axis = V.cert_axis, L = V.cert_L;
    const float d_a = axis == 0 ? c.dx : (axis == 1 ? c.dy : c.dz);
    const float o_a = axis == 0 ? c.ox : (axis == 1 ? c.oy : c.oz);
    const float h_a = axis == 0 ? c.hxh : (axis == 1 ? c.hyh : c.hzh);
    const bool fwd = V.cert_sign > 0;
    const bool can = fwd ? d_a > 0.04f : d_a < -0.04f;
    const float inv_da = __fdividef(1.0f, d_a);
    float f = r.f, tcur = r.tcur, step = r.step;      // the ray's state: identical in the lanes of its group

    for (;;) {
        if (STATS) { if (lane == 0) ++st[4]; if (!done) ++st[5]; }
        // ---- this lane's two samples: 2 sub + 1 and 2 sub + 2 additions of the step; t1 = the ray's next sample
        const float t1 = fadd(tcur, step);
        float ta = t1, tb = t1;
        {
            float tk = t1;
#pragma unroll
            for (int i = 1; i < kGrpSamples; ++i) {
                tk = fadd(tk, step);
                if (i == 2 * sub) ta = tk;
                if (i == 2 * sub + 1) tb = tk;
            }
            if (sub == 0) ta = t1;
        }
        Samp qa, qb;
        qa.in = qb.in = false; qa.vx = qa.vy = qa.vz = qb.vx = qb.vy = qb.vz = 0.0f;
        bool ready = false;
        int j = 0;
        const uint32_t* tab = bits;
        bool have_a = false, have_b = false;
        if (!done) {
            if (!(t1 <= c.tmax)) done = true;
            else {
                have_a = ta <= c.tmax; have_b = tb <= c.tmax;
                if (have_a) samp_pos(qa, ta, c, div_s, V);
                if (have_b) samp_pos(qb, tb, c, div_s, V);
                // a ray that carries exactly +1 passes all-(+1) slabs unchanged; one that carries 0 at the half-voxel step
                // passes all-0 slabs unchanged (|0| < 0.8 keeps the step, no sign change, f stays 0)
                const bool one = f == 1.0f, zero = f == 0.0f && step == c.half_s;
                if (sub == 0 && can && qa.in && (one || zero)) {
                    const float va = axis == 0 ? qa.vx : (axis == 1 ? qa.vy : qa.vz);
                    j = __float2int_rz(va) >> L;                 // (va >= 0: the sample is inside the march bounds)
                    tab = one ? bits : bits + kCertWords;
                    ready = (j >> 5) < kCertWords && ((tab[j >> 5] >> (j & 31)) & 1u);
                }
            }
        }
        ready = __shfl_sync(kFull, (int)ready, gl0) != 0;
        if (__all_sync(kFull, done)) break;
        if (__all_sync(kFull, done || ready)) {
            // ---- skip phase: every ray still marching stands before a run of certified slabs
            j = __shfl_sync(kFull, j, gl0);
            if (!done) {
                tab = (f == 1.0f) ? bits : bits + kCertWords;
                int k = j >> 5;
                const uint32_t w = tab[k];
                float bound;
                if (fwd) {
                    uint32_t inv = ~w & (0xffffffffu << (j & 31));
                    while (inv == 0u && k + 1 < kCertWords) { ++k; inv = ~tab[k]; }
                    const int je = inv ? 32 * k + __ffs(inv) - 1 : kCertSlabs;      // first slab that is not certified
                    bound = (float)(je << L);                                        // samples with va < bound are no-ops
                } else {
                    uint32_t inv = ~w & (0xffffffffu >> (31 - (j & 31)));
                    while (inv == 0u && k > 0) { --k; inv = ~tab[k]; }
                    const int je = inv ? 32 * k + 31 - __clz(inv) : -1;             // last slab (below) that is not certified
                    bound = (float)((je + 1) << L);                                  // samples with va >= bound are no-ops
                }
                // ray parameter at which the ray leaves the run (plain arithmetic: an estimate, verified below)
                const float t_end = fminf(((bound - h_a) * c.s - o_a) * inv_da, c.tmax);
                const float nf = __fdividef(t_end - t1, step);
                const int n = (nf < 1.0e6f ? (int)nf : 1000000);                     // samples, t1 included, up to the end of the run (estimate)
                int made = 1;
                float t2 = t1;
                if (n >= 2) {
                    t2 = cert_hop(tcur, step, n, made);          // (one short of the estimate: the last sample before the bound is taken next)
                    Samp b;
                    samp_pos(b, t2, c, div_s, V);
                    const float vb = axis == 0 ? b.vx : (axis == 1 ? b.vy : b.vz);
                    if (!(t2 <= c.tmax && b.in && (fwd ? vb < bound : vb >= bound))) { t2 = t1; made = 1; }   // (the estimate is conservative: rare)
                }
                if (STATS && sub == 0) { ++st[2]; st[1] += made; }
                tcur = t2;
            }
            continue;
        }
        // ---- sample phase (rays standing before a certified run wait)
        const bool act = !done && !ready;
        if (!act) { qa.in = false; qb.in = false; }
        samp_load(qa, V, rx, plane);          // (predicated on .in, which is false for lanes that do not take part)
        samp_load(qb, V, rx, plane);
        float va = 0.0f, vb = 0.0f;
        if (qa.in) va = samp_value(qa);
        if (qb.in) vb = samp_value(qb);
        float nsa = step, nsb = step;
        if (fabsf(va) < 1.0f) nsa = c.s;
        if (fabsf(va) < 0.8f) nsa = c.half_s;
        if (fabsf(vb) < 1.0f) nsb = c.s;
        if (fabsf(vb) < 0.8f) nsb = c.half_s;
        float prev = __shfl_up_sync(kFull, vb, 1, kGrpLanes);
        if (sub == 0) prev = f;
        // a sign change against the predecessor is looked into right here, by its own lane, under the assumption that
        // everything before it is accepted: along the shadow volume of an object nearly every sample is such a candidate
        // and nearly none ends the ray (no weight there)
        bool stop_a = !have_a || !qa.in || nsa != step;
        if (!stop_a && ((prev < 0.0f && va > 0.0f) || (prev > 0.0f && va < 0.0f)))
            stop_a = ray_event(V, prev, va, ta, step, qa.vx, qa.vy, qa.vz, c.dx, c.dy, c.dz).code != 0;
        bool stop_b = stop_a || !have_b || !qb.in || nsb != step;
        if (!stop_b && ((va < 0.0f && vb > 0.0f) || (va > 0.0f && vb < 0.0f)))
            stop_b = ray_event(V, va, vb, tb, step, qb.vx, qb.vy, qb.vz, c.dx, c.dy, c.dz).code != 0;
        const unsigned ma = (__ballot_sync(kFull, stop_a) >> gl0) & ((1u << kGrpLanes) - 1u);
        const unsigned mb = (__ballot_sync(kFull, stop_b) >> gl0) & ((1u << kGrpLanes) - 1u);
        const int first = min(ma ? 2 * (__ffs(ma) - 1) : kGrpSamples, mb ? 2 * (__ffs(mb) - 1) + 1 : kGrpSamples);   // index of the stop sample
        // the samples before the stop: f = fn each, nothing else
        {
            const int a = max(first - 1, 0);
            const float t_acc = __shfl_sync(kFull, (a & 1) ? tb : ta, gl0 + (a >> 1));
            const float f_acc = __shfl_sync(kFull, (a & 1) ? vb : va, gl0 + (a >> 1));
            if (act && first > 0) { tcur = t_acc; f = f_acc; if (STATS && sub == 0) st[0] += first; }
        }
        // the stop sample, by the sequential rule
        {
            const int sidx = min(first, kGrpSamples - 1), sl = gl0 + (sidx >> 1);
            const bool sb = (sidx & 1) != 0;
            const float tS = __shfl_sync(kFull, sb ? tb : ta, sl), vS = __shfl_sync(kFull, sb ? vb : va, sl);
            const float sx = __shfl_sync(kFull, sb ? qb.vx : qa.vx, sl), sy = __shfl_sync(kFull, sb ? qb.vy : qa.vy, sl),
                        sz = __shfl_sync(kFull, sb ? qb.vz : qa.vz, sl);
            const int fl = __shfl_sync(kFull, (int)(sb ? have_b : have_a) | ((int)(sb ? qb.in : qa.in) << 1), sl);
            if (act && first < kGrpSamples) {
                if (!(fl & 1)) {
                    done = true;                         // t > t_max: the ray ends without a hit
                } else {
                    tcur = tS;
                    if (fl & 2) {
                        if (STATS && sub == 0) { ++st[0]; if (f == 1.0f && vS == 1.0f) ++st[6]; }
                        if (fabsf(vS) < 1.0f) step = c.s;
                        if (fabsf(vS) < 0.8f) step = c.half_s;
                        if ((f < 0.0f && vS > 0.0f) || (f > 0.0f && vS < 0.0f)) {
                            if (STATS && sub == 0 && f < 0.0f) ++st[3];
                            const RayEvent e = ray_event(V, f, vS, tcur, step, sx, sy, sz, c.dx, c.dy, c.dz);
                            if (e.code == 3) {
                                if (sub == 0) ray_emit_hit(V, x, y, e.ts, e.sx, e.sy, e.sz, c.dx, c.dy, c.dz, nullptr, P.w);
                                hit = true;
                            }
                            if (e.code == 1 || e.code == 3) done = true;
                            else if (e.code == 0) f = vS;
                        } else {
                            f = vS;
                        }
                    }
                }
            }
        }
    }
    if (valid && !hit && sub == 0) {
        *((float*)((char*)V.ray + (size_t)y * V.ray_pitch) + x) = 0.0f;
        V.mask[(size_t)y * V.mask_pitch + x] = 0;
    }
    if (STATS && P.stats) {
#pragma unroll
        for (int k = 0; k < 8; ++k) if (st[k]) atomicAdd(P.stats + k, st[k]);
        if (P.hist && lane == 0) {
            atomicAdd(P.stats + 8 + min(15ull, st[4] / 32ull), 1ull);
            atomicMax(P.stats + 24, st[4]);
        }
        if (P.hist == 2 && lane == 0) {   // timeline: 4 words per warp after the 32 counters
            unsigned smid;
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            unsigned long long* rec = P.stats + 32 + 4 * ((size_t)blockIdx.x * (kRayThreads / 32) + warp);
            rec[0] = t_begin; rec[1] = global_ns(); rec[2] = st[4]; rec[3] = ((unsigned long long)smid << 32);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// k_ray_certify: the ray-space certificate of k_raycast_cert.
// One thread per (8 x 4 pixel tile, slab); the lanes of a warp are 32 consecutive tiles of one slab (neighbouring tiles
// read neighbouring voxels of the same rows), the ballot of their verdicts is one word of the bit table.
// The box of a (tile, slab): the slab's planes along the dominant axis (a sample whose base voxel is in the slab has its
// corners in planes [j 2^L, (j + 1) 2^L]), and across it the bounding box of the tile's four corner rays between the two
// faces of the slab (the position on a ray at a given coordinate along the axis is a linear-fractional function of the
// pixel: over a pixel rectangle its extremes are at the corners; along the slab it is linear), widened by margins a
// hundred times the rounding error of the sampler's own arithmetic.  Every voxel of the box must hold exactly +1.
// ---------------------------------------------------------------------------------------------
struct CertParams {
    const float* tsdf;
    int rx, ry, rz;
    float R[9], t[3];
    float fx, fy, cx, cy, voxel;
    int x0, y0, x1, y1;
    int txc, n_tiles, G;
    int axis, sgn, L, n_slabs;
    uint32_t* table;
};
constexpr int kCertWarps = 8;        // warps per CTA
constexpr int kCertPerWarp = 4;      // consecutive slabs a warp certifies for its 32 tiles (the per-tile set-up is shared)

__global__ void __launch_bounds__(32 * kCertWarps) k_ray_certify(const __grid_constant__ CertParams P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x, j0 = (blockIdx.y * kCertWarps + warp) * kCertPerWarp;
    const int ti = g * 32 + lane;
    const int a = P.axis, b = a == 2 ? 0 : a + 1, c = a == 0 ? 2 : (a == 1 ? 0 : 1);   // (a, b, c) = (0,1,2), (1,2,0), (2,0,1)
    const int ra = a == 0 ? P.rx : (a == 1 ? P.ry : P.rz);
    const int rb = b == 0 ? P.rx : (b == 1 ? P.ry : P.rz);
    const int rc = c == 0 ? P.rx : (c == 1 ? P.ry : P.rz);
    // ---- per tile: position across the axis as a function of the coordinate along it, for the four corner rays:
    //      v_b = ib + v_a * mb, v_c = ic + v_a * mc (voxel coordinates)
    bool tile_ok = ti < P.n_tiles;
    float mb[4], mc[4], ib[4], ic[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) mb[q] = mc[q] = ib[q] = ic[q] = 0.0f;
    const float inv_s = 1.0f / P.voxel;
    const float cam_a = 0.5f * (float)(ra - 1) + P.t[a] * inv_s;       // the camera centre along the axis
    if (tile_ok) {
        const int wy = ti / P.txc, wx = ti - wy * P.txc;
        const int X0 = P.x0 + 8 * wx, Y0 = P.y0 + 4 * wy;
        tile_ok = X0 < P.x1 && Y0 < P.y1;
        if (tile_ok) {
            const int X1 = min(X0 + 7, P.x1 - 1), Y1 = min(Y0 + 3, P.y1 - 1);
            const float cam_b = 0.5f * (float)(rb - 1) + P.t[b] * inv_s, cam_c = 0.5f * (float)(rc - 1) + P.t[c] * inv_s;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float ux = ((float)((q & 1) ? X1 : X0) - P.cx) / P.fx, uy = ((float)((q & 2) ? Y1 : Y0) - P.cy) / P.fy;
                const float Da = P.R[3 * a] * ux + P.R[3 * a + 1] * uy + P.R[3 * a + 2];
                const float Db = P.R[3 * b] * ux + P.R[3 * b + 1] * uy + P.R[3 * b + 2];
                const float Dc = P.R[3 * c] * ux + P.R[3 * c + 1] * uy + P.R[3 * c + 2];
                if (!((float)P.sgn * Da > 0.05f * sqrtf(Da * Da + Db * Db + Dc * Dc))) tile_ok = false;
                mb[q] = Db / Da; mc[q] = Dc / Da;
                ib[q] = cam_b - cam_a * mb[q]; ic[q] = cam_c - cam_a * mc[q];
            }
        }
    }
    const uint32_t* __restrict__ vol = reinterpret_cast<const uint32_t*>(P.tsdf);
    const size_t sy = (size_t)P.rx, sz = (size_t)P.rx * P.ry;
#pragma unroll 1
    for (int s = 0; s < kCertPerWarp; ++s) {
        const int j = j0 + s;
        if (j >= P.n_slabs) break;                                   // (warp-uniform)
        const int a0 = j << P.L, a1 = min((j + 1) << P.L, ra - 1);   // planes along the axis that corners of the slab's samples lie in
        bool ok = tile_ok && a0 <= ra - 1;
        // the slab is cut off at the camera centre (no sample lies behind it)
        if (P.sgn > 0) ok = ok && (float)((j + 1) << P.L) + 0.02f > cam_a;
        else ok = ok && (float)a0 - 0.02f < cam_a;
        uint32_t acc = 0u, acz = 0u;         // some voxel is not +1 / not (+-)0
#pragma unroll 1
        for (int pl = a0; pl <= a1 && ok && (acc == 0u || acz == 0u); ++pl) {
            // samples with a corner in plane pl have their base in {pl - 1, pl}: coordinate along the axis in [pl - 1, pl + 1),
            // inside the slab, widened
            float c0 = (float)max(pl - 1, a0) - 0.02f, c1 = (float)min(pl + 1, (j + 1) << P.L) + 0.02f;
            if (P.sgn > 0) { c0 = fmaxf(c0, cam_a); c1 = fmaxf(c1, cam_a); }
            else { c0 = fminf(c0, cam_a); c1 = fminf(c1, cam_a); }
            float mnb = INFINITY, mxb = -INFINITY, mnc = INFINITY, mxc = -INFINITY;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float b0 = fmaf(c0, mb[q], ib[q]), b1 = fmaf(c1, mb[q], ib[q]);
                const float e0 = fmaf(c0, mc[q], ic[q]), e1 = fmaf(c1, mc[q], ic[q]);
                mnb = fminf(mnb, fminf(b0, b1)); mxb = fmaxf(mxb, fmaxf(b0, b1));
                mnc = fminf(mnc, fminf(e0, e1)); mxc = fmaxf(mxc, fmaxf(e0, e1));
            }
            if (!(fabsf(mnb) < 1.0e6f && fabsf(mxb) < 1.0e6f && fabsf(mnc) < 1.0e6f && fabsf(mxc) < 1.0e6f)) { ok = false; break; }   // (or NaN)
            const int blo = max((int)floorf(mnb - 0.05f), 0), bhi = min((int)floorf(mxb + 0.05f) + 1, rb - 1);
            const int clo = max((int)floorf(mnc - 0.05f), 0), chi = min((int)floorf(mxc + 0.05f) + 1, rc - 1);
            // (a rectangle wholly outside the volume belongs to samples outside the march bounds: never consulted)
            if (blo > bhi || clo > chi || (bhi - blo + 1) * (chi - clo + 1) > 400) { ok = false; break; }
            // rectangle in (x, y, z)
            const int xlo = a == 0 ? pl : (b == 0 ? blo : clo), xhi = a == 0 ? pl : (b == 0 ? bhi : chi);
            const int ylo = a == 1 ? pl : (b == 1 ? blo : clo), yhi = a == 1 ? pl : (b == 1 ? bhi : chi);
            const int zlo = a == 2 ? pl : (b == 2 ? blo : clo), zhi = a == 2 ? pl : (b == 2 ? bhi : chi);
            const int nx = xhi - xlo;
            for (int z = zlo; z <= zhi; ++z) {
                const uint32_t* row = vol + (size_t)z * sz + (size_t)ylo * sy + xlo;
                for (int y = ylo; y <= yhi; ++y, row += sy) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (i <= nx) { const uint32_t v = __ldg(row + i); acc |= v ^ 0x3f800000u; acz |= v << 1; }
                    for (int i = 8; i <= nx; ++i) { const uint32_t v = __ldg(row + i); acc |= v ^ 0x3f800000u; acz |= v << 1; }
                }
            }
        }
        const uint32_t word = __ballot_sync(0xffffffffu, ok && acc == 0u);
        const uint32_t wordz = __ballot_sync(0xffffffffu, ok && acz == 0u);
        if (lane == 0) {
            P.table[(size_t)j * P.G + g] = word;
            P.table[(size_t)(kCertSlabs + j) * P.G + g] = wordz;
        }
    }
}

static int cert_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("EMF_RAY_CERT"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
static size_t cert_table_bytes(int tiles_x, int tiles_y) {   // CTA tiles of the raycast (kTileW x kTileH pixels)
    const int n_tiles = 4 * tiles_x * tiles_y;
    return (size_t)2 * kCertSlabs * ((n_tiles + 31) / 32) * sizeof(uint32_t);      // all-(+1) table, all-0 table
}

// certify volume d (rays as fill_ray_vol prepared them) into the workspace and point the raycast at it; false = not applicable
static bool launch_certify(RayVol& d, const float* K, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!cert_enabled() || !workspace || d.fg_probs || d.bmap || d.tube) return false;
    if (d.x1 <= d.x0 || d.y1 <= d.y0) return false;
    if (!(K[0] != 0.0f && K[4] != 0.0f)) return false;
    const int tiles_y = (d.y1 - d.y0 + kTileH - 1) / kTileH;
    if (cert_table_bytes(d.tiles_x, tiles_y) > workspace_bytes) return false;
    uint32_t* table = (uint32_t*)workspace;
    CertParams C;
    C.tsdf = d.tsdf; C.rx = d.rx; C.ry = d.ry; C.rz = d.rz;
    for (int k = 0; k < 9; ++k) C.R[k] = d.R[k];
    for (int k = 0; k < 3; ++k) C.t[k] = d.t[k];
    C.fx = K[0]; C.fy = K[4]; C.cx = K[2]; C.cy = K[5]; C.voxel = d.voxel;
    C.x0 = d.x0; C.y0 = d.y0; C.x1 = d.x1; C.y1 = d.y1;
    C.txc = 2 * d.tiles_x; C.n_tiles = 4 * d.tiles_x * tiles_y; C.G = (C.n_tiles + 31) / 32;
    // the volume axis the optical axis (third column of rot_CO) is closest to
    int axis = 0;
    for (int k = 1; k < 3; ++k) if (fabsf(d.R[3 * k + 2]) > fabsf(d.R[3 * axis + 2])) axis = k;
    C.axis = axis; C.sgn = d.R[3 * axis + 2] > 0.0f ? 1 : -1;
    const int ra = axis == 0 ? d.rx : (axis == 1 ? d.ry : d.rz);
    int L = EMF_CERT_L;
    while ((kCertSlabs << L) < ra) ++L;
    C.L = L;
    const int n_slabs = min(kCertSlabs, (((ra + (1 << L) - 1) >> L) + 31) & ~31);     // whole words of slab bits
    C.n_slabs = n_slabs;
    C.table = table;
    const int per_cta = kCertWarps * kCertPerWarp;
    k_ray_certify<<<dim3((unsigned)C.G, (unsigned)((n_slabs + per_cta - 1) / per_cta)), 32 * kCertWarps, 0, stream>>>(C);
    d.tiles_y = tiles_y;
    d.cert = table; d.cert_G = C.G; d.cert_txc = C.txc; d.cert_axis = axis; d.cert_sign = C.sgn; d.cert_L = L;
    d.cert_nw = n_slabs / 32;
    return true;
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
    d.tiles_y = 0;
    d.cert = nullptr; d.cert_G = 0; d.cert_txc = 0; d.cert_axis = 2; d.cert_sign = 1; d.cert_L = 1; d.cert_nw = 0;
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
    P.hit_voxel = hit_voxel; P.write_all = 0; P.stats = nullptr; P.hist = 0; P.sched = nullptr; P.timeline = nullptr;
    const int blocks = P.v[0].tiles_x * ((h + kTileH - 1) / kTileH);
    k_raycast<false, 0><<<blocks, kRayThreads, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}

static size_t sched_offset(int width, int height) {
    return (cert_table_bytes((width + kTileW - 1) / kTileW, (height + kTileH - 1) / kTileH) + 255) & ~(size_t)255;
}
static int sched_front_pct() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("EMF_RAY_FRONT"); v = e ? max(0, min(100, atoi(e))) : 100; }
    return v;
}
extern "C" EMF_API size_t emf_raycast_workspace_bytes(int width, int height) {
    if (width <= 0 || height <= 0) return 0;
    const size_t tiles = (size_t)((width + kTileW - 1) / kTileW) * ((height + kTileH - 1) / kTileH);
    return sched_offset(width, height) + (4 + 2 * tiles) * sizeof(uint32_t);
}

extern "C" EMF_API int emf_raycast_volumes(int n_vol, const emf_volume* vols, const emf_pose* T_co, const float K[9],
                                   const int* rects, const emf_image* ray_out, const emf_image* vert_out,
                                   const emf_image* norm_out, const emf_image* mask_out, uint64_t* stats,
                                   emf_stream_t stream) {
    return emf_raycast_volumes_opt(n_vol, vols, T_co, K, rects, ray_out, vert_out, norm_out, mask_out, stats, nullptr, 0, 0, stream);
}
extern "C" EMF_API int emf_raycast_volumes_ws(int n_vol, const emf_volume* vols, const emf_pose* T_co, const float K[9],
                                      const int* rects, const emf_image* ray_out, const emf_image* vert_out,
                                      const emf_image* norm_out, const emf_image* mask_out, uint64_t* stats,
                                      void* workspace, size_t workspace_bytes, emf_stream_t stream) {
    return emf_raycast_volumes_opt(n_vol, vols, T_co, K, rects, ray_out, vert_out, norm_out, mask_out, stats, workspace, workspace_bytes,
                                   EMF_RAY_CERTIFICATE | EMF_RAY_SCHEDULE, stream);
}

extern "C" EMF_API int emf_raycast_volumes_opt(int n_vol, const emf_volume* vols, const emf_pose* T_co, const float K[9],
                                       const int* rects, const emf_image* ray_out, const emf_image* vert_out,
                                       const emf_image* norm_out, const emf_image* mask_out, uint64_t* stats,
                                       void* workspace, size_t workspace_bytes, unsigned options, emf_stream_t stream) {
    if (n_vol <= 0 || !vols || !T_co || !K || !ray_out || !vert_out || !norm_out || !mask_out) return EMF_ERR_INVALID;
    if (n_vol > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    static RayParams P;          // (tens of kilobytes: not on the stack; the ABI is single-threaded per process like the reference)
    const int w = ray_out[0].width, h = ray_out[0].height;
    for (int i = 0; i < n_vol; ++i) {
        const int rc = fill_ray_vol(P.v[i], vols[i], T_co[i], &ray_out[i], &vert_out[i], &norm_out[i], &mask_out[i],
                                    rects ? rects + 4 * i : nullptr, w, h);
        if (rc != EMF_OK) return rc;
    }
    P.w = w; P.h = h;
    for (int k = 0; k < 9; ++k) P.K[k] = K[k];
    P.hit_voxel = nullptr; P.write_all = 1; P.stats = (unsigned long long*)stats;
    P.timeline = nullptr;
    { const char* e = getenv("EMF_RAY_TIMELINE"); if (e && e[0] == '1' && stats) { P.timeline = P.stats; P.stats = nullptr; stats = nullptr; } }
    { const char* e = getenv("EMF_RAY_HIST"); P.hist = (e && e[0] == '1') ? 1 : ((e && e[0] == '2') ? 2 : 0); }
    bool jump = false;
    for (int i = 0; i < n_vol; ++i) jump = jump || P.v[i].bmap != nullptr;
    const cudaStream_t cs = (cudaStream_t)stream;
    int n_rest = n_vol;
    if (!jump && (((options & EMF_RAY_CERTIFICATE) && workspace && ((uintptr_t)workspace & 15) == 0) || (options & EMF_RAY_WIDE))) {
        // ray-space certificate for the first volume without a foreground mask (the background: long rays through free
        // space); that volume gets its own launch of the certified march, the others follow in the plain one
        for (int i = 0; i < n_vol; ++i) {
            if (P.v[i].fg_probs) continue;
            bool wide = false;
            if ((options & EMF_RAY_WIDE) && !(options & EMF_RAY_CERTIFICATE) && !P.v[i].bmap && !P.v[i].tube && P.v[i].x1 > P.v[i].x0 && P.v[i].y1 > P.v[i].y0) {
                RayVol& d = P.v[i];
                d.tiles_y = (d.y1 - d.y0 + kTileH - 1) / kTileH;
                d.cert = nullptr; d.cert_G = 0; d.cert_txc = 2 * d.tiles_x; d.cert_axis = 0; d.cert_sign = 1; d.cert_L = 0; d.cert_nw = 0;
                wide = true;
            }
            if (wide || ((options & EMF_RAY_CERTIFICATE) && launch_certify(P.v[i], K, workspace, workspace_bytes, cs))) {
                static CertRayParams Pc;
                Pc.v = P.v[i];
                Pc.v.first_block = 0;
                Pc.w = w; Pc.h = h;
                for (int k = 0; k < 9; ++k) Pc.K[k] = K[k];
                Pc.stats = P.stats; Pc.hist = P.hist;
                const unsigned cb = (unsigned)(4 * Pc.v.tiles_x * Pc.v.tiles_y);      // one CTA per 8 x 4 pixel tile
                if (stats) k_raycast_cert<true><<<cb, kRayThreads, 0, cs>>>(Pc);
                else k_raycast_cert<false><<<cb, kRayThreads, 0, cs>>>(Pc);
                for (int k = i; k + 1 < n_vol; ++k) P.v[k] = P.v[k + 1];
                --n_rest;
            }
            break;
        }
    }
    int64_t blocks = 0;
    for (int i = 0; i < n_rest; ++i) {
        P.v[i].first_block = (int)blocks;
        blocks += (int64_t)P.v[i].tiles_x * ((P.v[i].y1 - P.v[i].y0 + kTileH - 1) / kTileH);
    }
    P.n_vol = n_rest;
    P.sched = nullptr; P.sched_n0 = 0; P.sched_max = 0;
    if (!jump && n_rest == n_vol && blocks > 0 && workspace && ((uintptr_t)workspace & 15) == 0 && (options & EMF_RAY_SCHEDULE) &&
        workspace_bytes >= emf_raycast_workspace_bytes(w, h)) {
        const int64_t n0 = n_vol > 1 ? P.v[1].first_block : blocks;
        const int64_t n_max = (int64_t)((w + kTileW - 1) / kTileW) * ((h + kTileH - 1) / kTileH);
        if (n0 > 0 && n0 <= n_max) {
            P.sched = (uint32_t*)((char*)workspace + sched_offset(w, h));
            P.sched_n0 = (int)n0; P.sched_max = (int)n_max;
        }
    }
    if (blocks > 0) {
        if (jump) {
            if (stats) k_raycast<true, 2><<<(unsigned)blocks, kRayThreads, 0, cs>>>(P);
            else k_raycast<false, 2><<<(unsigned)blocks, kRayThreads, 0, cs>>>(P);
        } else {
            if (stats) k_raycast<true, 0><<<(unsigned)blocks, kRayThreads, 0, cs>>>(P);
            else k_raycast<false, 0><<<(unsigned)blocks, kRayThreads, 0, cs>>>(P);
            if (P.sched && !(options & EMF_RAY_SCHEDULE_DEFER)) k_ray_schedule<<<1, 1024, 0, cs>>>(P.sched, P.sched_n0, P.sched_max, sched_front_pct());
        }
    }
    return launch_status();
}

extern "C" EMF_API int emf_raycast_schedule_update(int width, int height, const int rect0[4], void* workspace, size_t workspace_bytes,
                                                   emf_stream_t stream) {
    if (width <= 0 || height <= 0 || !workspace || ((uintptr_t)workspace & 15) != 0 || workspace_bytes < emf_raycast_workspace_bytes(width, height))
        return EMF_ERR_INVALID;
    const int x0 = rect0 ? max(rect0[0], 0) : 0, y0 = rect0 ? max(rect0[1], 0) : 0;
    const int x1 = rect0 ? min(rect0[2], width) : width, y1 = rect0 ? min(rect0[3], height) : height;
    if (x1 <= x0 || y1 <= y0) return EMF_OK;
    const int64_t n0 = (int64_t)((x1 - x0 + kTileW - 1) / kTileW) * ((y1 - y0 + kTileH - 1) / kTileH);
    const int64_t n_max = (int64_t)((width + kTileW - 1) / kTileW) * ((height + kTileH - 1) / kTileH);
    if (n0 > n_max) return EMF_ERR_INVALID;
    k_ray_schedule<<<1, 1024, 0, (cudaStream_t)stream>>>((uint32_t*)((char*)workspace + sched_offset(width, height)), (int)n0, (int)n_max,
                                                        sched_front_pct());
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
