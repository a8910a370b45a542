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
    // device-resident Levenberg-Marquardt loop (emf_track_iterate): poses, modes and tile rectangles come from `states`, and
    // the last CTA of a volume runs the host part of the iteration (TSDF::reduceHessians / computePoseUpdate)
    emf_track_state* states;   // nullptr: poses / modes from the table above (emf_track_linearise)
    int phase;                 // 0: linearise (or error only) at the current pose, then the step; 1: the trial pose, then accept / reject
    float K[9];
    float tau, eps1, eps2, nu_init;
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


// ---------------------------------------------------------------------------------------------------------------------
// The host part of one tracker iteration, on the device (one thread).  Statement by statement emf::TSDF::reduceHessians
// (reference src/core/TSDF.cpp:267-283) and computePoseUpdate (:285-338); the pose algebra is Sophus' SE3 exp / log in double.
// ---------------------------------------------------------------------------------------------------------------------
__device__ void lm_so3_exp(const double w[3], double R[9]) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
    double a, b;
    if (th < 1e-10) { a = 1.0; b = 0.5; } else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double oo = 0;
            for (int k = 0; k < 3; ++k) oo += O[3 * i + k] * O[3 * k + j];
            R[3 * i + j] = (i == j ? 1.0 : 0.0) + a * O[3 * i + j] + b * oo;
        }
}
__device__ void lm_se3_exp(const double x[6], double R[9], double t[3]) {
    const double* u = x; const double* w = x + 3;
    lm_so3_exp(w, R);
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
    double b, c;
    if (th < 1e-10) { b = 0.5; c = 0.0; } else { b = (1.0 - cos(th)) / th2; c = (th - sin(th)) / (th2 * th); }
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    for (int i = 0; i < 3; ++i) {
        double acc = 0;
        for (int j = 0; j < 3; ++j) {
            double oo = 0;
            for (int k = 0; k < 3; ++k) oo += O[3 * i + k] * O[3 * k + j];
            acc += ((i == j ? 1.0 : 0.0) + b * O[3 * i + j] + c * oo) * u[j];
        }
        t[i] = acc;
    }
}
// |log(T)| of a rigid transform (only the norm of the twist is needed, :299)
__device__ double lm_se3_log_norm(const double R[9], const double t[3]) {
    double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    const double th = acos(c);
    double v[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]}, w[3];
    if (th < 1e-10) { for (int k = 0; k < 3; ++k) w[k] = 0.5 * v[k]; }
    else if (3.14159265358979323846 - th < 1e-6) {
        double ax[3] = {sqrt(fmax((R[0] + 1.0) * 0.5, 0.0)), sqrt(fmax((R[4] + 1.0) * 0.5, 0.0)), sqrt(fmax((R[8] + 1.0) * 0.5, 0.0))};
        int k = ax[0] >= ax[1] ? (ax[0] >= ax[2] ? 0 : 2) : (ax[1] >= ax[2] ? 1 : 2);
        double col[3] = {(R[k] + (k == 0)) * 0.5, (R[3 + k] + (k == 1)) * 0.5, (R[6 + k] + (k == 2)) * 0.5};
        double n = 0; for (int i = 0; i < 3; ++i) { col[i] /= ax[k]; n += col[i] * col[i]; }
        n = sqrt(n);
        const double sgn = (col[0] * v[0] + col[1] * v[1] + col[2] * v[2]) < 0 ? -1.0 : 1.0;
        for (int i = 0; i < 3; ++i) w[i] = sgn * th * col[i] / n;
    } else { const double f = th / (2.0 * sin(th)); for (int k = 0; k < 3; ++k) w[k] = f * v[k]; }
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double g;
    if (th < 1e-10) g = 1.0 / 12.0; else { const double hf = 0.5 * th; g = (1.0 - th * cos(hf) / (2.0 * sin(hf))) / (th * th); }
    double n2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    for (int i = 0; i < 3; ++i) {
        double acc = 0;
        for (int j = 0; j < 3; ++j) {
            double oo = 0;
            for (int k = 0; k < 3; ++k) oo += O[3 * i + k] * O[3 * k + j];
            acc += ((i == j ? 1.0 : 0.0) - 0.5 * O[3 * i + j] + g * oo) * t[j];
        }
        n2 += acc * acc;
    }
    return sqrt(n2);
}
// x = M^-1 b by LU with partial pivoting in float (cv::solve, DECOMP_LU, on a 6 x 6 float system); false if singular
__device__ bool lm_solve6(const float A[36], float mu, const float b[6], float x[6]) {
    float M[6][7];
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) M[i][j] = A[6 * i + j] + (i == j ? mu : 0.0f); M[i][6] = b[i]; }
    for (int c = 0; c < 6; ++c) {
        int p = c;
        for (int r = c + 1; r < 6; ++r) if (fabsf(M[r][c]) > fabsf(M[p][c])) p = r;
        if (fabsf(M[p][c]) < 1.1920929e-6f) return false;        // FLT_EPSILON * 10, as cv::LU for float
        if (p != c) for (int j = 0; j < 7; ++j) { const float tmp = M[c][j]; M[c][j] = M[p][j]; M[p][j] = tmp; }
        const float d = -1.0f / M[c][c];
        for (int r = c + 1; r < 6; ++r) {
            const float f = M[r][c] * d;
            for (int j = c + 1; j < 7; ++j) M[r][j] += f * M[c][j];
        }
    }
    for (int i = 5; i >= 0; --i) {
        float v = M[i][6];
        for (int j = i + 1; j < 6; ++j) v -= M[i][j] * x[j];
        x[i] = v / M[i][i];
    }
    return true;
}
// tile rectangle of the pixels whose point can lie inside the volume's grid, for pose (R, t) = T_co (as box_screen_rect)
__device__ void lm_tile_rect(const TrackVol& V, const double R[9], const double t[3], const float K[9], int w, int h, int rect[4]) {
    const int tiles_x = (w + 31) / 32, tiles_y = (h + 7) / 8;
    const double b[3] = {(0.5 * V.rx + 1.0) * V.voxel, (0.5 * V.ry + 1.0) * V.voxel, (0.5 * V.rz + 1.0) * V.voxel};
    double x0 = 1e30, y0 = 1e30, x1 = -1e30, y1 = -1e30;
    bool full = false;
    for (int c = 0; c < 8 && !full; ++c) {
        const double p[3] = {(c & 1 ? b[0] : -b[0]) - t[0], (c & 2 ? b[1] : -b[1]) - t[1], (c & 4 ? b[2] : -b[2]) - t[2]};
        double q[3];
        for (int k = 0; k < 3; ++k) q[k] = R[k] * p[0] + R[3 + k] * p[1] + R[6 + k] * p[2];
        if (q[2] < 1e-3) { full = true; break; }
        const double u = K[0] * q[0] + K[1] * q[1] + K[2] * q[2], v = K[3] * q[0] + K[4] * q[1] + K[5] * q[2];
        const double wq = K[6] * q[0] + K[7] * q[1] + K[8] * q[2];
        if (wq < 1e-6) { full = true; break; }
        x0 = fmin(x0, u / wq); x1 = fmax(x1, u / wq); y0 = fmin(y0, v / wq); y1 = fmax(y1, v / wq);
    }
    if (full) { rect[0] = 0; rect[1] = 0; rect[2] = tiles_x; rect[3] = tiles_y; return; }
    int ix0 = (int)floor(x0) - 2, iy0 = (int)floor(y0) - 2, ix1 = (int)ceil(x1) + 3, iy1 = (int)ceil(y1) + 3;
    ix0 = max(ix0, 0); iy0 = max(iy0, 0); ix1 = min(ix1, w); iy1 = min(iy1, h);
    if (ix1 <= ix0 || iy1 <= iy0) { rect[0] = rect[1] = rect[2] = rect[3] = 0; return; }
    rect[0] = ix0 / 32; rect[1] = iy0 / 8; rect[2] = (ix1 + 31) / 32; rect[3] = (iy1 + 7) / 8;
}

// after the sums of a launch are known (thread 0 of the volume's last CTA).  rec = the volume's record.
__device__ void lm_step(const TrackParams& P, const TrackVol& V, emf_track_state& S, const float* rec, int mode) {
    if (P.phase == 0) {
        S.iterations += 1;
        if (mode == 1) {
            S.linearisations += 1;
            float bmax = 0.0f;
            for (int k = 0; k < 36; ++k) S.A[k] = rec[k];
            for (int k = 0; k < 6; ++k) { S.b[k] = rec[36 + k]; bmax = fmaxf(bmax, fabsf(rec[36 + k])); }
            if (bmax < P.eps1) { S.converged = 1; return; }                       // reduceHessians (:278-282)
        }
        S.err = rec[42];
        if (S.first_iteration) {                                                   // (:290-295)
            float dmax = S.A[0];
            for (int k = 1; k < 6; ++k) dmax = fmaxf(dmax, S.A[7 * k]);
            S.mu = (double)P.tau * (double)dmax;
            S.first_iteration = 0;
        }
        float x[6];
        for (int k = 0; k < 6; ++k) x[k] = S.x[k];
        const bool solved = lm_solve6(S.A, (float)S.mu, S.b, x);                   // cv::solve zeroes dst when the LU fails: the step test below then ends the run
        for (int k = 0; k < 6; ++k) S.x[k] = solved ? x[k] : 0.0f;
        float xn = 0.0f;
        for (int k = 0; k < 6; ++k) xn += S.x[k] * S.x[k];
        xn = sqrtf(xn);
        if ((double)xn < (double)P.eps2 * (lm_se3_log_norm(S.R, S.t) + (double)P.eps2)) { S.converged = 1; return; }   // (:298-302)
        double mx[6], Ri[9], ti[3];
        for (int k = 0; k < 6; ++k) { mx[k] = -(double)S.x[k]; }
        lm_se3_exp(mx, Ri, ti);
        for (int k = 0; k < 9; ++k) S.R_old[k] = S.R[k];
        for (int k = 0; k < 3; ++k) S.t_old[k] = S.t[k];
        double Rn[9], tn[3];                                                       // pose_incr * rel_pose_CO (:308-310)
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j)
                Rn[3 * i + j] = (Ri[3 * i] * S.R_old[j] + Ri[3 * i + 1] * S.R_old[3 + j]) + Ri[3 * i + 2] * S.R_old[6 + j];
            tn[i] = ((Ri[3 * i] * S.t_old[0] + Ri[3 * i + 1] * S.t_old[1]) + Ri[3 * i + 2] * S.t_old[2]) + ti[i];
        }
        for (int k = 0; k < 9; ++k) S.R[k] = Rn[k];
        for (int k = 0; k < 3; ++k) S.t[k] = tn[k];
        S.trial_pending = 1;
    } else {
        S.err_new = rec[42];                                                       // (:312-315)
        float g = 0.0f;
        for (int k = 0; k < 6; ++k) g += -S.x[k] * ((float)S.mu * -S.x[k] - S.b[k]);
        g *= 0.5f;
        S.rho = g != 0.0f ? ((double)S.err - (double)S.err_new) / (double)g : -1.0;
        if (S.rho > 0.0) {                                                         // (:320-326)
            const double c = 2.0 * S.rho - 1.0;
            S.mu *= fmax(1.0 / 3.0, 1.0 - c * c * c);
            S.nu = (double)P.nu_init;
            S.evaluate_gradient = 1;
        } else {                                                                   // (:327-336)
            for (int k = 0; k < 9; ++k) S.R[k] = S.R_old[k];
            for (int k = 0; k < 3; ++k) S.t[k] = S.t_old[k];
            S.mu *= S.nu;
            S.nu *= (double)P.nu_init;
            S.evaluate_gradient = 0;
            }
        S.trial_pending = 0;
    }
}

__global__ void __launch_bounds__(kTrackThreads, 2) k_track(const __grid_constant__ TrackParams P) {
    const int vi = blockIdx.y;
    const TrackVol& V = P.v[vi];
    int mode = V.mode;
    float Rm[9], tm[3];
    int tx0 = V.tx0, ty0 = V.ty0, tx1 = V.tx1, ty1 = V.ty1, lx0 = V.lx0, ly0 = V.ly0, ltx = V.ltx, ltiles = V.ltiles;
    float inv_ltx = V.inv_ltx;
    if (P.states) {
        // (the last CTA of this volume rewrites the state only after every CTA of the volume has passed its ticket, i.e.
        //  after all of them have read it here)
        const emf_track_state& S = P.states[vi];
        if (mode != 0) mode = P.phase == 0 ? (S.converged ? 0 : (S.evaluate_gradient ? 1 : 2)) : (S.trial_pending ? 2 : 0);
#pragma unroll
        for (int k = 0; k < 9; ++k) Rm[k] = (float)S.R[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) tm[k] = (float)S.t[k];
        __shared__ int s_rect[4];
        if (threadIdx.x == 0 && mode != 0) lm_tile_rect(V, S.R, S.t, P.K, P.w, P.h, s_rect);
        __syncthreads();
        tx0 = s_rect[0]; ty0 = s_rect[1]; tx1 = s_rect[2]; ty1 = s_rect[3];
        lx0 = tx0; ly0 = ty0; ltx = tx1 - tx0; ltiles = ltx * (ty1 - ty0);
        inv_ltx = ltx > 0 ? 1.0f / (float)ltx : 0.0f;
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) Rm[k] = V.R[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) tm[k] = V.t[k];
    }
    if (mode == 0 || (int)blockIdx.x >= V.n_cta) return;
    const bool lin = mode == 1;
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
    for (int tile = blockIdx.x; tile < ltiles; tile += n_cta) {
        const int tyl = __float2int_rz(((float)tile + 0.5f) * inv_ltx);   // tile / ltx (exact: tile < 2^20)
        const int tx = lx0 + tile - tyl * ltx, ty = ly0 + tyl;
        const int x = tx * 32 + (threadIdx.x & 31), y = ty * 8 + (threadIdx.x >> 5);
        if (x >= P.w || y >= P.h) continue;
        if (tx < tx0 || tx >= tx1 || ty < ty0 || ty >= ty1) {
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
            const float qx = fadd(tm[0], dot_xyz(Rm[0], Rm[1], Rm[2], px, py, pz));
            const float qy = fadd(tm[1], dot_xyz(Rm[3], Rm[4], Rm[5], px, py, pz));
            const float qz = fadd(tm[2], dot_xyz(Rm[6], Rm[7], Rm[8], px, py, pz));
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
            out[45] = __uint_as_float((unsigned)tx0 | ((unsigned)tx1 << 16));   // where int_weights is defined
            out[46] = __uint_as_float((unsigned)ty0 | ((unsigned)ty1 << 16));
        }
    } else if (threadIdx.x == 0) {
        out[42] = (float)(s_tot[27] * (double)out[44]);
    }
    if (P.states) {      // the host part of the iteration, here
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); lm_step(P, V, P.states[vi], out, mode); }
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
    P.states = nullptr; P.phase = 0;
    for (int k = 0; k < 9; ++k) P.K[k] = K ? K[k] : 0.0f;
    P.tau = P.eps1 = P.eps2 = P.nu_init = 0.0f;
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

extern "C" EMF_API int emf_track_iterate(int n_vol, const emf_volume* vols, emf_track_state* states, const emf_pose* T_co_hint,
                                         const emf_image* points, const float K[9], const emf_image* assoc,
                                         const emf_track_lm_params* lm, const emf_image* int_weights, float* records,
                                         void* workspace, size_t workspace_bytes, int n_iterations, emf_stream_t stream) {
    if (n_vol <= 0 || !vols || !states || !T_co_hint || !K || !assoc || !lm || !int_weights || !records || !workspace ||
        !image_ok(points, 12) || n_iterations < 0)
        return EMF_ERR_INVALID;
    if (n_vol > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    if (workspace_bytes < emf_track_workspace_bytes(n_vol) || !aligned16(workspace)) return EMF_ERR_INVALID;
    TrackParams P;
    const int w = points->width, h = points->height;
    const int tiles_x = (w + 31) / 32, tiles_y = (h + 7) / 8;
    if (tiles_x > 0xffff || tiles_y > 0xffff || tiles_x * tiles_y >= (1 << 20)) return EMF_ERR_UNSUPPORTED;
    int max_cta = 1;
    for (int i = 0; i < n_vol; ++i) {
        TrackVol& d = P.v[i];
        const emf_volume& v = vols[i];
        if (!v.tsdf || !v.weights || !res_ok(v.res)) return EMF_ERR_INVALID;
        if (!image_ok(&int_weights[i], 4) || int_weights[i].width != w || int_weights[i].height != h) return EMF_ERR_INVALID;
        if (!image_ok(&assoc[i], 4) || assoc[i].width != w || assoc[i].height != h) return EMF_ERR_INVALID;
        d.mode = 1;       // the real mode of a launch is derived from the volume's state on the device
        d.tsdf = v.tsdf; d.weights = v.weights; d.grads = v.grads;
        d.assoc = (const float*)assoc[i].ptr; d.assoc_pitch = assoc[i].pitch;
        d.wimg = (float*)int_weights[i].ptr; d.wimg_pitch = int_weights[i].pitch;
        d.vals = nullptr; d.vals_pitch = 0; d.huber = nullptr; d.huber_pitch = 0; d.g6 = nullptr;
        for (int k = 0; k < 9; ++k) d.R[k] = T_co_hint[i].R[k];
        for (int k = 0; k < 3; ++k) d.t[k] = T_co_hint[i].t[k];
        d.rx = v.res[0]; d.ry = v.res[1]; d.rz = v.res[2];
        d.voxel = v.voxel_size;
        // the grid of a volume is sized by the rectangle of its starting pose (the rectangle itself follows the pose on the device)
        const double b[3] = {(0.5 * v.res[0] + 1.0) * v.voxel_size, (0.5 * v.res[1] + 1.0) * v.voxel_size, (0.5 * v.res[2] + 1.0) * v.voxel_size};
        int r[4];
        box_screen_rect(b, &T_co_hint[i], K, w, h, r);
        d.tx0 = r[0] / 32; d.ty0 = r[1] / 8; d.tx1 = (r[2] + 31) / 32; d.ty1 = (r[3] + 7) / 8;
        d.lx0 = d.tx0; d.ly0 = d.ty0; d.ltx = d.tx1 - d.tx0; d.ltiles = d.ltx * (d.ty1 - d.ty0); d.inv_ltx = 0.0f;
        d.n_cta = (d.ltiles + 3) / 4;
        if (d.n_cta < 1) d.n_cta = 1;
        if (d.n_cta > kTrackCtasPerVol) d.n_cta = kTrackCtasPerVol;
        if (d.n_cta > max_cta) max_cta = d.n_cta;
    }
    P.n_vol = n_vol; P.w = w; P.h = h;
    P.points = (const float*)points->ptr; P.points_pitch = points->pitch;
    P.huber_thresh = lm->huber_thresh; P.max_weight = lm->max_tsdf_weight;
    P.out = records;
    P.tickets = (unsigned*)workspace;
    P.slots = (double*)((char*)workspace + kTrackTicketBytes);
    P.states = states;
    for (int k = 0; k < 9; ++k) P.K[k] = K[k];
    P.tau = lm->tau; P.eps1 = lm->eps1; P.eps2 = lm->eps2; P.nu_init = lm->nu_init;
    const dim3 grid(max_cta, n_vol);
    for (int it = 0; it < n_iterations; ++it) {
        P.phase = 0;
        k_track<<<grid, kTrackThreads, 0, (cudaStream_t)stream>>>(P);
        P.phase = 1;
        k_track<<<grid, kTrackThreads, 0, (cudaStream_t)stream>>>(P);
    }
    return launch_status();
}
