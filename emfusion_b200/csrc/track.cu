// track.cu -- one Levenberg-Marquardt linearisation of the SDF tracker for a whole batch of volumes in ONE launch.
//
// Replaces, per volume and per tracker iteration, the chain emf::EMFusion::performTracking drives
// (reference src/core/EMFusion.cpp:672-722):
//   TSDF::computeGradients     -> setTo(0) + kernel_computePoseGradients   src/core/TSDF.cpp:194-202, src/core/cuda/TSDF.cu:603-660
//   TSDF::computeTSDFVals      -> kernel_getVolumeVals(tsdfVol)            src/core/TSDF.cpp:204-212
//   TSDF::computeTSDFWeights   -> kernel_getVolumeVals(tsdfWeights)        src/core/TSDF.cpp:214-221
//   TSDF::computeHuberWeights  -> abs, divide(scalar, mat), min            src/core/TSDF.cpp:223-232
//   TSDF::normalizeTSDFWeights -> min, normalize(NORM_INF)                 src/core/TSDF.cpp:234-242
//   TSDF::combineWeights       -> multiply x 2                             src/core/TSDF.cpp:244-255
//   TSDF::computeHessians      -> kernel_computeAb (36 + 6 floats / pixel) src/core/TSDF.cpp:257-265, src/core/cuda/TSDF.cu:729-766
//   TSDF::reduceAb             -> multSingletonCol x 2, cv::cuda::reduce x 2   src/core/TSDF.cpp:375-388
//   TSDF::computeError         -> sqr, multiply, sum                       src/core/TSDF.cpp:390-394
// i.e. ~20 launches, a 44 MB `As` buffer written once and read twice, and three blocking downloads per volume and
// iteration.  Here one thread owns one pixel: it gathers the TSDF value, the integration weight and the 8 corner
// gradients (forward differences on the fly, or the materialised float3 volume when the caller has one), forms the
// 6-vector J = [g, p x g] and accumulates  sum w J J^T (21 unique entries),  sum w f J,  sum w f^2  and  max w_int  in
// registers; warps reduce with shuffles, the CTA in shared memory (double), CTAs through a per-volume slot array that the
// last CTA to finish folds in a fixed order -- a deterministic result with no atomics on the data.
//
// Arithmetic: every per-pixel quantity (p, v, f, w_int, g, p x g, Huber weight) is computed with the reference build's
// own instruction sequence (read from its PTX: "xyz" contraction for R p, lerp as fma(1-a, lo, a*hi), IEEE divisions), so
// the optional per-pixel outputs are bit-identical to the reference's images.  Two things are not bit-reproducible in
// the reference itself and are matched to tolerance only: the order of the 307 200-term float sums (cv::cuda::reduce), and
// the NORM_INF normalisation, which is applied here once to the sums instead of once per pixel.
#include "common.cuh"
#include <float.h>

namespace emfb {

void box_screen_rect(const double b[3], const emf_pose* T_co, const float K[9], int width, int height, int rect_out[4]);

constexpr int kTrackThreads = 256;
constexpr int kTrackAcc = 29;        // 21 (A upper triangle) + 6 (b) + err + max
constexpr int kTrackCtasPerVol = 296; // most CTAs (= reduction slots) one volume can use: 2 per SM
constexpr size_t kTrackTicketBytes = 512;   // EMF_MAX_VOLUMES tickets at the head of the workspace

struct TrackVol {
    const float* tsdf;
    const float* weights;
    const float* grads;                      // nullable: float3 per voxel
    const float* assoc; size_t assoc_pitch;  // association weights of this volume
    float* vals; size_t vals_pitch;          // nullable: tsdfVals
    float* wimg; size_t wimg_pitch;          // combined weights BEFORE the NORM_INF scale (mode 1: written, mode 2: read)
    float* huber; size_t huber_pitch;        // nullable: trackWeights
    float* g6;                               // nullable: grads, W*H x 6 continuous
    float R[9], t[3];
    int rx, ry, rz;
    float voxel;
    int mode;                                // 0 skip, 1 linearise, 2 error only
    int tx0, ty0, tx1, ty1;                  // tiles (32 x 8 px) with a pixel whose point can lie inside the volume (exclusive upper)
    int lx0, ly0, ltx, ltiles;               // the tiles this launch loops over: origin, width, count (all tiles if it has
    float inv_ltx;                           //   to write complete optional images, else the rectangle above)
    int n_cta;                               // CTAs working on this volume (sized by the tiles inside the rectangle)
};

struct TrackParams {
    TrackVol v[EMF_MAX_VOLUMES];
    int n_vol, w, h;
    const float* points; size_t points_pitch;
    float huber_thresh, max_weight;
    float* out;            // n_vol x EMF_TRACK_RECORD floats
    double* slots;         // n_vol x kTrackCtasPerVol x kTrackAcc
    unsigned* tickets;     // n_vol
};

// gradient at an integer voxel: the materialised volume, or forward differences with a zero last plane per axis
// (reference src/core/cuda/TSDF.cu:436-447 after the setTo(0) of src/core/TSDF.cpp:121)
__device__ __forceinline__ void track_grad_at(const TrackVol& V, int x, int y, int z, float g[3]) {
    const int64_t i = ((int64_t)z * V.ry + y) * V.rx + x;
    if (V.grads) {
        const float* p = V.grads + 3 * i;
        g[0] = __ldg(p); g[1] = __ldg(p + 1); g[2] = __ldg(p + 2);
        return;
    }
    if (x >= V.rx - 1 || y >= V.ry - 1 || z >= V.rz - 1) { g[0] = g[1] = g[2] = 0.f; return; }
    const float* p = V.tsdf + i;
    const float f = __ldg(p);
    g[0] = fsub(__ldg(p + 1), f);
    g[1] = fsub(__ldg(p + V.rx), f);
    g[2] = fsub(__ldg(p + (int64_t)V.ry * V.rx), f);
}

__global__ void __launch_bounds__(kTrackThreads, 2) k_track(const __grid_constant__ TrackParams P) {
    const int vi = blockIdx.y;
    const TrackVol& V = P.v[vi];
    if (V.mode == 0 || (int)blockIdx.x >= V.n_cta) return;
    const bool lin = V.mode == 1;
    const unsigned n_cta = (unsigned)V.n_cta;
    float acc[kTrackAcc];
#pragma unroll
    for (int k = 0; k < kTrackAcc; ++k) acc[k] = 0.0f;

    const float frx = (float)V.rx, fry = (float)V.ry, frz = (float)V.rz;
    const float hx = fmul((float)(V.rx - 1), 0.5f), hy = fmul((float)(V.ry - 1), 0.5f), hz = fmul((float)(V.rz - 1), 0.5f);
    float* const out = P.out + (size_t)vi * EMF_TRACK_RECORD;
    // error-only launches weigh with the image of the last linearisation, which is defined on ITS tile rectangle only
    int wx0 = 0, wx1 = 0, wy0 = 0, wy1 = 0;
    if (!lin) {
        const unsigned a = __float_as_uint(out[45]), b = __float_as_uint(out[46]);
        wx0 = a & 0xffff; wx1 = a >> 16; wy0 = b & 0xffff; wy1 = b >> 16;
    }
    for (int tile = blockIdx.x; tile < V.ltiles; tile += n_cta) {
        const int tyl = __float2int_rz(((float)tile + 0.5f) * V.inv_ltx);   // tile / ltx (exact: tile < 2^20)
        const int tx = V.lx0 + tile - tyl * V.ltx, ty = V.ly0 + tyl;
        const int x = tx * 32 + (threadIdx.x & 31), y = ty * 8 + (threadIdx.x >> 5);
        if (x >= P.w || y >= P.h) continue;
        if (tx < V.tx0 || tx >= V.tx1 || ty < V.ty0 || ty >= V.ty1) {
            // (only when complete optional images were asked for) no point of this tile can gather from the volume:
            // every per-pixel quantity is 0 there
            if (V.vals) *((float*)((char*)V.vals + (size_t)y * V.vals_pitch) + x) = 0.0f;
            if (lin) {
                *((float*)((char*)V.wimg + (size_t)y * V.wimg_pitch) + x) = 0.0f;
                if (V.huber) *((float*)((char*)V.huber + (size_t)y * V.huber_pitch) + x) = 0.0f;
                if (V.g6) {
                    float* gp = V.g6 + 6 * ((size_t)y * P.w + x);
#pragma unroll
                    for (int k = 0; k < 6; ++k) gp[k] = 0.0f;
                }
            }
            continue;
        }
        const float* pp = (const float*)((const char*)P.points + (size_t)y * P.points_pitch) + 3 * x;
        const float px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
        float f = 0.0f, wint = 0.0f;
        float J[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (!(pz <= 0.0f)) {
            const float qx = fadd(V.t[0], dot_xyz(V.R[0], V.R[1], V.R[2], px, py, pz));
            const float qy = fadd(V.t[1], dot_xyz(V.R[3], V.R[4], V.R[5], px, py, pz));
            const float qz = fadd(V.t[2], dot_xyz(V.R[6], V.R[7], V.R[8], px, py, pz));
            const float vx = fadd(hx, fdiv(qx, V.voxel)), vy = fadd(hy, fdiv(qy, V.voxel)), vz = fadd(hz, fdiv(qz, V.voxel));
            if (!out_of(vx, vy, vz, 1.0f, frx, fry, frz)) {           // getVolumeVals (TSDF.cu:680-684)
                f = trilinear(V.tsdf, V.rx, V.ry, vx, vy, vz);
                if (lin) wint = trilinear(V.weights, V.rx, V.ry, vx, vy, vz);
            }
            if (lin && !out_of(vx, vy, vz, 2.0f, frx, fry, frz)) {   // computePoseGradients (TSDF.cu:622-626)
                const TriSetup s(vx, vy, vz, V.rx, V.ry);
                const int lx = __float2int_rz(vx), ly = __float2int_rz(vy), lz = __float2int_rz(vz);
                float g[8][3];
#pragma unroll
                for (int c = 0; c < 8; ++c) track_grad_at(V, lx + (c & 1), ly + ((c >> 1) & 1), lz + (c >> 2), g[c]);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    J[k] = fdiv(s.combine(g[0][k], g[1][k], g[2][k], g[3][k], g[4][k], g[5][k], g[6][k], g[7][k]), V.voxel);
                // [p]x g as the reference build evaluates it (SASS of kernel_computePoseGradients for sm_100a: ptxas contracts
                // the second product of every component; the products with the literal zeros only decide zero signs)
                J[3] = ffma(qy, J[2], -fmul(qz, J[1]));
                J[4] = ffma(-qx, J[2], fmul(qz, J[0]));
                J[5] = ffma(qx, J[1], -fmul(qy, J[0]));
            }
        }
        if (V.vals) *((float*)((char*)V.vals + (size_t)y * V.vals_pitch) + x) = f;
        float* wp = (float*)((char*)V.wimg + (size_t)y * V.wimg_pitch) + x;
        float wgt;
        if (lin) {
            // Huber: cv::cuda::divide(scalar, mat) is 0 where the divisor is 0; then min(., 1)
            const float af = fabsf(f);
            const float hub = af != 0.0f ? fminf(fdiv(P.huber_thresh, af), 1.0f) : 0.0f;
            const float wc = fminf(wint, P.max_weight);
            acc[28] = fmaxf(acc[28], fabsf(wc));
            const float a = __ldg((const float*)((const char*)V.assoc + (size_t)y * V.assoc_pitch) + x);
            wgt = fmul(fmul(hub, wc), a);
            *wp = wgt;
            if (V.huber) *((float*)((char*)V.huber + (size_t)y * V.huber_pitch) + x) = hub;
            if (V.g6) {
                float* gp = V.g6 + 6 * ((size_t)y * P.w + x);
#pragma unroll
                for (int k = 0; k < 6; ++k) gp[k] = J[k];
            }
            int n = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = i; j < 6; ++j) { acc[n] = ffma(fmul(J[i], J[j]), wgt, acc[n]); ++n; }
#pragma unroll
            for (int i = 0; i < 6; ++i) acc[21 + i] = ffma(fmul(f, J[i]), wgt, acc[21 + i]);
        } else {
            wgt = (tx >= wx0 && tx < wx1 && ty >= wy0 && ty < wy1) ? *wp : 0.0f;
        }
        acc[27] = ffma(fmul(f, f), wgt, acc[27]);
    }

    // ---- warp -> CTA -> volume
    __shared__ double s_part[kTrackThreads / 32][kTrackAcc];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // warp level: the 28 sums by recursive halving (at offset o a lane keeps the half of its values selected by its bit o
    // and adds the partner's copy of that half: 16 + 8 + 4 + 2 + 1 shuffles, lane k ends up with the total of sum k), the
    // maximum by a butterfly
    {
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = k < 28 ? acc[k] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int k = 0; k < o; ++k) {
                const float send = up ? v[k] : v[k + o];
                const float keep = up ? v[k + o] : v[k];
                v[k] = fadd(keep, __shfl_xor_sync(0xffffffffu, send, o));
            }
        }
        float m = acc[28];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane < 28) s_part[warp][lane] = (double)v[0];
        if (lane == 28) s_part[warp][28] = (double)m;
    }
    __syncthreads();
    double* slot = P.slots + ((size_t)vi * kTrackCtasPerVol + blockIdx.x) * kTrackAcc;
    if (threadIdx.x < kTrackAcc) {
        double a = s_part[0][threadIdx.x];
        for (int wv = 1; wv < kTrackThreads / 32; ++wv)
            a = (threadIdx.x == 28) ? fmax(a, s_part[wv][threadIdx.x]) : a + s_part[wv][threadIdx.x];
        slot[threadIdx.x] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(P.tickets + vi, 1u) == n_cta - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // fold the CTA slots in a fixed order: warp p sums the slots c = p (mod 8) in ascending order, then the 8 partial
    // sums are combined in ascending order
    __shared__ double s_tot[kTrackAcc];
    if (lane < kTrackAcc) {
        const double* s0 = P.slots + (size_t)vi * kTrackCtasPerVol * kTrackAcc + lane;
        double a = 0.0;   // (every accumulated quantity, the maximum included, is >= 0 or a sum)
#pragma unroll 4
        for (unsigned c = warp; c < n_cta; c += kTrackThreads / 32) {
            const double b = __ldcg(s0 + (size_t)c * kTrackAcc);
            a = (lane == 28) ? fmax(a, b) : a + b;
        }
        s_part[warp][lane] = a;
    }
    __syncthreads();
    if (threadIdx.x < kTrackAcc) {
        double a = s_part[0][threadIdx.x];
        for (int wv = 1; wv < kTrackThreads / 32; ++wv)
            a = (threadIdx.x == 28) ? fmax(a, s_part[wv][threadIdx.x]) : a + s_part[wv][threadIdx.x];
        s_tot[threadIdx.x] = a;
    }
    __syncthreads();
    if (lin) {
        // cv::cuda::normalize(NORM_INF, alpha = 1): scale = 1 / max|w| (0 when the norm vanishes), applied as a float
        const double wmax = s_tot[28];
        const float scale = wmax > DBL_EPSILON ? (float)(1.0 / wmax) : 0.0f;
        if (threadIdx.x < 36) {
            const int i = threadIdx.x / 6, j = threadIdx.x % 6;
            const int a = i < j ? i : j, b = i < j ? j : i;
            const int n = a * 6 - a * (a - 1) / 2 + (b - a);     // index into the packed upper triangle
            out[threadIdx.x] = (float)(s_tot[n] * (double)scale);
        } else if (threadIdx.x < 42) {
            out[threadIdx.x] = (float)(s_tot[21 + threadIdx.x - 36] * (double)scale);
        } else if (threadIdx.x == 42) {
            out[42] = (float)(s_tot[27] * (double)scale);
            out[43] = (float)wmax;
            out[44] = scale;
            out[45] = __uint_as_float((unsigned)V.tx0 | ((unsigned)V.tx1 << 16));   // where int_weights is defined
            out[46] = __uint_as_float((unsigned)V.ty0 | ((unsigned)V.ty1 << 16));
        }
    } else if (threadIdx.x == 0) {
        out[42] = (float)(s_tot[27] * (double)out[44]);
    }
    if (threadIdx.x == 0) P.tickets[vi] = 0;   // ready for the next launch on this stream
}

// intWeights as the reference holds it (normalised): img <- img * scale[vol]
__global__ void __launch_bounds__(256) k_track_scale(Img<const float> src, Img<float> dst, const float* __restrict__ rec) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= src.w || y >= src.h) return;
    const unsigned a = __float_as_uint(__ldg(rec + 45)), b = __float_as_uint(__ldg(rec + 46));
    const bool in = blockIdx.x >= (a & 0xffff) && blockIdx.x < (a >> 16) && blockIdx.y >= (b & 0xffff) && blockIdx.y < (b >> 16);
    dst.at(y, x) = in ? fmul(src.at(y, x), __ldg(rec + 44)) : 0.0f;
}

}  // namespace emfb

using namespace emfb;

extern "C" EMF_API size_t emf_track_workspace_bytes(int n_vol) {
    if (n_vol <= 0 || n_vol > EMF_MAX_VOLUMES) return 0;
    return kTrackTicketBytes + (size_t)n_vol * kTrackCtasPerVol * kTrackAcc * sizeof(double);
}

extern "C" EMF_API int emf_track_linearise(int n_vol, const emf_volume* vols, const emf_pose* T_co, const int* modes,
                                           const emf_image* points, const float K[9], const emf_image* assoc, float huber_thresh,
                                           float max_tsdf_weight, const emf_image* int_weights, const emf_image* tsdf_vals,
                                           const emf_image* track_weights, float* const* pose_grads, float* records,
                                           void* workspace, size_t workspace_bytes, emf_stream_t stream) {
    if (n_vol <= 0 || !vols || !T_co || !modes || !int_weights || !records || !workspace || !image_ok(points, 12))
        return EMF_ERR_INVALID;
    if (n_vol > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    if (workspace_bytes < emf_track_workspace_bytes(n_vol) || !aligned16(workspace)) return EMF_ERR_INVALID;
    TrackParams P;
    const int w = points->width, h = points->height;
    bool any = false;
    int max_cta = 1;
    for (int i = 0; i < n_vol; ++i) {
        TrackVol& d = P.v[i];
        d.mode = modes[i];
        if (d.mode < 0 || d.mode > 2) return EMF_ERR_INVALID;
        if (d.mode == 0) continue;
        any = true;
        const emf_volume& v = vols[i];
        if (!v.tsdf || !v.weights || !res_ok(v.res)) return EMF_ERR_INVALID;
        if (!image_ok(&int_weights[i], 4) || int_weights[i].width != w || int_weights[i].height != h) return EMF_ERR_INVALID;
        if (d.mode == 1 && (!assoc || !image_ok(&assoc[i], 4) || assoc[i].width != w || assoc[i].height != h))
            return EMF_ERR_INVALID;
        d.tsdf = v.tsdf; d.weights = v.weights; d.grads = v.grads;
        d.assoc = d.mode == 1 ? (const float*)assoc[i].ptr : nullptr; d.assoc_pitch = d.mode == 1 ? assoc[i].pitch : 0;
        d.wimg = (float*)int_weights[i].ptr; d.wimg_pitch = int_weights[i].pitch;
        d.vals = nullptr; d.vals_pitch = 0; d.huber = nullptr; d.huber_pitch = 0; d.g6 = nullptr;
        if (tsdf_vals && tsdf_vals[i].ptr) {
            if (!image_ok(&tsdf_vals[i], 4) || tsdf_vals[i].width != w || tsdf_vals[i].height != h) return EMF_ERR_INVALID;
            d.vals = (float*)tsdf_vals[i].ptr; d.vals_pitch = tsdf_vals[i].pitch;
        }
        if (track_weights && track_weights[i].ptr) {
            if (!image_ok(&track_weights[i], 4) || track_weights[i].width != w || track_weights[i].height != h) return EMF_ERR_INVALID;
            d.huber = (float*)track_weights[i].ptr; d.huber_pitch = track_weights[i].pitch;
        }
        if (pose_grads) d.g6 = pose_grads[i];
        for (int k = 0; k < 9; ++k) d.R[k] = T_co[i].R[k];
        for (int k = 0; k < 3; ++k) d.t[k] = T_co[i].t[k];
        d.rx = v.res[0]; d.ry = v.res[1]; d.rz = v.res[2];
        d.voxel = v.voxel_size;
        const int tiles_x = (w + 31) / 32, tiles_y = (h + 7) / 8;
        if (tiles_x > 0xffff || tiles_y > 0xffff) return EMF_ERR_UNSUPPORTED;
        d.tx0 = 0; d.ty0 = 0; d.tx1 = tiles_x; d.ty1 = tiles_y;
        if (K) {
            // a point gathers only if it lies inside the voxel grid, |p| < (R - 1) / 2 * voxel per axis, and it lies on its
            // pixel's ray: outside the screen rectangle of that box (taken one voxel larger, padded by 2 px) nothing gathers
            const double b[3] = {(0.5 * v.res[0] + 1.0) * v.voxel_size, (0.5 * v.res[1] + 1.0) * v.voxel_size,
                                 (0.5 * v.res[2] + 1.0) * v.voxel_size};
            int r[4];
            box_screen_rect(b, &T_co[i], K, w, h, r);
            d.tx0 = r[0] / 32; d.ty0 = r[1] / 8; d.tx1 = (r[2] + 31) / 32; d.ty1 = (r[3] + 7) / 8;
            if (r[2] <= r[0] || r[3] <= r[1]) { d.tx1 = d.tx0; d.ty1 = d.ty0; }
        }
        // complete optional images: loop over every tile (zeros outside the rectangle); otherwise over the rectangle only
        const bool full = d.vals || d.huber || d.g6;
        d.lx0 = full ? 0 : d.tx0; d.ly0 = full ? 0 : d.ty0;
        d.ltx = full ? tiles_x : d.tx1 - d.tx0;
        d.ltiles = full ? tiles_x * tiles_y : (d.tx1 - d.tx0) * (d.ty1 - d.ty0);
        d.inv_ltx = d.ltx > 0 ? 1.0f / (float)d.ltx : 0.0f;
        if (d.ltiles >= (1 << 20)) return EMF_ERR_UNSUPPORTED;
        // four tiles of real work per CTA
        const int64_t rect_tiles = (int64_t)(d.tx1 - d.tx0) * (d.ty1 - d.ty0);
        d.n_cta = (int)((rect_tiles + 3) / 4);
        if (d.n_cta < 1) d.n_cta = 1;
        if (d.n_cta > kTrackCtasPerVol) d.n_cta = kTrackCtasPerVol;
        if (d.n_cta > max_cta) max_cta = d.n_cta;
    }
    if (!any) return EMF_OK;
    P.n_vol = n_vol; P.w = w; P.h = h;
    P.points = (const float*)points->ptr; P.points_pitch = points->pitch;
    P.huber_thresh = huber_thresh; P.max_weight = max_tsdf_weight;
    P.out = records;
    P.tickets = (unsigned*)workspace;                       // fixed place: they return to 0 after every launch
    P.slots = (double*)((char*)workspace + kTrackTicketBytes);
    const dim3 grid(max_cta, n_vol);
    k_track<<<grid, kTrackThreads, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}

extern "C" EMF_API int emf_track_workspace_init(void* workspace, size_t workspace_bytes, emf_stream_t stream) {
    if (!workspace || workspace_bytes == 0) return EMF_ERR_INVALID;
    return cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream) == cudaSuccess ? EMF_OK : EMF_ERR_CUDA;
}

extern "C" EMF_API int emf_track_normalised_weights(const emf_image* int_weights, const float* record, const emf_image* out,
                                                    emf_stream_t stream) {
    if (!record || !image_ok(int_weights, 4) || !image_ok(out, 4) || !same_size(int_weights, out)) return EMF_ERR_INVALID;
    const dim3 grid((out->width + 31) / 32, (out->height + 7) / 8);
    k_track_scale<<<grid, 256, 0, (cudaStream_t)stream>>>(view<const float>(int_weights), view<float>(out), record);
    return launch_status();
}
