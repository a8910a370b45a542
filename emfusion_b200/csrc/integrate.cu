// integrate.cu -- depth -> TSDF integration for one or many volumes in ONE launch.
//
// Replaces emf::cuda::TSDF::updateTSDF (reference src/core/cuda/TSDF.cu:327-427),
// called once per volume on its own stream by emf::EMFusion::integrateDepth
// (src/core/EMFusion.cpp:865-889).  Here every volume of the frame is a row in a
// kernel-parameter descriptor table and a CTA finds its volume with a short
// search over the table's block prefix sums.
//
// Traversal: a thread owns 4 consecutive x-voxels (one float4 of tsdf, one of
// weights), a warp a contiguous 512-byte run, a CTA 512 voxels.  Voxels whose
// projection misses the image are never read or written (as in the reference);
// the others move 16 B (update), 8 B (occluded/unseen) or 4 B (occluded/seen).
// All arithmetic is the canonical sequence of emf_math.cuh, so the result is
// bit-identical to the reference build for any input.
#include "common.cuh"

namespace emfb {

struct IntVol {
    float* tsdf;
    float* weights;
    const float* assoc;   // this volume's association image
    size_t assoc_pitch;
    float R[9];           // T_OC
    float t[3];
    int rx, ry, rz;
    float voxel, trunc;
    int first_block;      // prefix sum of CTAs
};

struct IntParams {
    IntVol v[EMF_MAX_VOLUMES];
    int n_vol;
    int total_blocks;
    const float* depth;
    size_t depth_pitch;
    int w, h;
    float K[9];
    float max_weight;
};

constexpr int kIntThreads = 128;
constexpr int kVec = 4;

// class of a voxel after projection
enum : int { kSkip = 0, kCheckOnly = 1, kValid = 2 };

template <bool PINHOLE, int VEC>
__global__ void __launch_bounds__(kIntThreads) k_integrate(const __grid_constant__ IntParams P) {
    // ---- which volume? (uniform per CTA; table lives in the constant bank)
    int lo = 0, hi = P.n_vol - 1;
    const int b = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (P.v[mid].first_block <= b) lo = mid; else hi = mid - 1;
    }
    const IntVol& V = P.v[lo];
    const int rx = V.rx, ry = V.ry;
    const int64_t n_vox = (int64_t)rx * ry * V.rz;
    const int64_t i0 = ((int64_t)(b - V.first_block) * kIntThreads + threadIdx.x) * VEC;
    if (i0 >= n_vox) return;

    const int64_t row = i0 / rx;            // z*Ry + y
    const int x0 = (int)(i0 - row * rx);
    const int z = (int)(row / ry);
    const int y = (int)(row - (int64_t)z * ry);

    const float s = V.voxel;
    // (i - (R-1)/2.f) * voxelSize ; (R-1)*0.5 is exact
    const float cy = fmul(fsub((float)y, fmul((float)(ry - 1), 0.5f)), s);
    const float cz = fmul(fsub((float)z, fmul((float)(V.rz - 1), 0.5f)), s);
    const float hx = fmul((float)(rx - 1), 0.5f);
    const float my0 = fmul(V.R[1], cy), my1 = fmul(V.R[4], cy), my2 = fmul(V.R[7], cy);

    int cls[VEC];
    int pix_x[VEC], pix_y[VEC];
    float dep[VEC], pcx[VEC], pcy[VEC], pcz[VEC];
    int any = 0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const float cx = fmul(fsub((float)(x0 + j), hx), s);
        pcx[j] = fadd(V.t[0], ffma(V.R[2], cz, ffma(V.R[0], cx, my0)));
        pcy[j] = fadd(V.t[1], ffma(V.R[5], cz, ffma(V.R[3], cx, my1)));
        pcz[j] = fadd(V.t[2], ffma(V.R[8], cz, ffma(V.R[6], cx, my2)));
        cls[j] = kCheckOnly;
        dep[j] = 0.f; pix_x[j] = 0; pix_y[j] = 0;
        if (pcz[j] > 0.0f) {
            float qx, qy, qz;
            if (PINHOLE) {
                qx = ffma(P.K[2], pcz[j], fmul(P.K[0], pcx[j]));
                qy = ffma(P.K[5], pcz[j], fmul(P.K[4], pcy[j]));
                qz = pcz[j];
            } else {
                qx = dot_yxz(P.K[0], P.K[1], P.K[2], pcx[j], pcy[j], pcz[j]);
                qy = dot_yxz(P.K[3], P.K[4], P.K[5], pcx[j], pcy[j], pcz[j]);
                qz = dot_yxz(P.K[6], P.K[7], P.K[8], pcx[j], pcy[j], pcz[j]);
            }
            const int px = __float2int_rn(fdiv(qx, qz));
            const int py = __float2int_rn(fdiv(qy, qz));
            if (px < 0 || px >= P.w || py < 0 || py >= P.h) {
                cls[j] = kSkip;
            } else {
                pix_x[j] = px; pix_y[j] = py;
                const float d = __ldg((const float*)((const char*)P.depth + (size_t)py * P.depth_pitch) + px);
                dep[j] = d;
                if (d > 0.0f) cls[j] = kValid;
            }
        }
        any |= cls[j];
    }
    if (!any) return;

    float* wp = V.weights + i0;
    float* tp = V.tsdf + i0;
    float w[VEC], tv[VEC];
    if (VEC == 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wp);
        w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) w[j] = wp[j];
    }

    float sdf[VEC];
    int need_t = 0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        sdf[j] = 0.f;
        if (cls[j] == kValid) {
            const float lx = fdiv(fsub((float)pix_x[j], P.K[2]), P.K[0]);
            const float ly = fdiv(fsub((float)pix_y[j], P.K[5]), P.K[4]);
            const float lambda = fsqrt(fadd(ffma(lx, lx, fmul(ly, ly)), 1.0f));
            const float inv_lambda = frcp(lambda);
            const float nrm = norm3(pcx[j], pcy[j], pcz[j]);
            sdf[j] = ffma(-nrm, inv_lambda, dep[j]);   // depth - (1/lambda)*|pc| as one FFMA (reference SASS)
            if (sdf[j] >= -V.trunc) need_t |= 1 << j;
        }
    }
    bool have_t = false;
    if (need_t) {
        have_t = true;
        if (VEC == 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(tp);
            tv[0] = t4.x; tv[1] = t4.y; tv[2] = t4.z; tv[3] = t4.w;
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) tv[j] = tp[j];
        }
    }

    int wrote_t = 0, wrote_w = 0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        if (cls[j] == kCheckOnly) {
            // behind the camera or no depth: un-mark never-seen voxels (TSDF.cu:349-353, 369-374)
            if (w[j] == 0.0f) { tv[j] = 0.0f; wrote_t |= 1 << j; }
        } else if (cls[j] == kValid) {
            if (need_t & (1 << j)) {
                const float q = fdiv(sdf[j], V.trunc);
                const float val = copysignf(fminf(1.0f, fabsf(q)), sdf[j]);
                float a = 1.0f;
                if (sdf[j] < V.trunc)
                    a = __ldg((const float*)((const char*)V.assoc + (size_t)pix_y[j] * V.assoc_pitch) + pix_x[j]);
                const float ws = fadd(w[j], a);
                if (ws > 0.0f) {
                    tv[j] = fdiv(ffma(w[j], tv[j], fmul(val, a)), ws);
                    w[j] = fminf(ws, P.max_weight);
                    wrote_t |= 1 << j; wrote_w |= 1 << j;
                }
            } else if (w[j] == 0.0f) {
                tv[j] = -1.0f; wrote_t |= 1 << j;   // occluded and never seen
            }
        }
    }

    if (wrote_w) {
        if (VEC == 4) *reinterpret_cast<float4*>(wp) = make_float4(w[0], w[1], w[2], w[3]);
        else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) if (wrote_w & (1 << j)) wp[j] = w[j];
        }
    }
    if (wrote_t) {
        if (VEC == 4 && (have_t || wrote_t == 0xF)) {
            *reinterpret_cast<float4*>(tp) = make_float4(tv[0], tv[1], tv[2], tv[3]);
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) if (wrote_t & (1 << j)) tp[j] = tv[j];
        }
    }
}

static int launch_integrate(int n_vol, const emf_volume* vols, const emf_pose* T_oc, const float* K,
                            const emf_image* depth, const emf_image* assoc, float max_weight,
                            cudaStream_t stream) {
    if (n_vol <= 0 || !vols || !T_oc || !K || !assoc) return EMF_ERR_INVALID;
    if (n_vol > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    if (!image_ok(depth, 4)) return EMF_ERR_INVALID;
    IntParams P;  // ~10 KB descriptor table, passed by value as the kernel parameter
    bool vec_ok = true;
    int64_t blocks = 0;
    for (int i = 0; i < n_vol; ++i) {
        const emf_volume& v = vols[i];
        if (!v.tsdf || !v.weights || !res_ok(v.res) || !image_ok(&assoc[i], 4) || !same_size(&assoc[i], depth))
            return EMF_ERR_INVALID;
        vec_ok = vec_ok && (v.res[0] % kVec == 0) && aligned16(v.tsdf) && aligned16(v.weights);
    }
    const int vec = vec_ok ? kVec : 1;
    for (int i = 0; i < n_vol; ++i) {
        const emf_volume& v = vols[i];
        IntVol& d = P.v[i];
        d.tsdf = v.tsdf; d.weights = v.weights;
        d.assoc = (const float*)assoc[i].ptr; d.assoc_pitch = assoc[i].pitch;
        for (int k = 0; k < 9; ++k) d.R[k] = T_oc[i].R[k];
        for (int k = 0; k < 3; ++k) d.t[k] = T_oc[i].t[k];
        d.rx = v.res[0]; d.ry = v.res[1]; d.rz = v.res[2];
        d.voxel = v.voxel_size; d.trunc = v.truncdist;
        d.first_block = (int)blocks;
        const int64_t n_vox = (int64_t)v.res[0] * v.res[1] * v.res[2];
        blocks += (n_vox + (int64_t)kIntThreads * vec - 1) / ((int64_t)kIntThreads * vec);
        if (blocks > 0x7fffffff) return EMF_ERR_UNSUPPORTED;
    }
    P.n_vol = n_vol; P.total_blocks = (int)blocks;
    P.depth = (const float*)depth->ptr; P.depth_pitch = depth->pitch; P.w = depth->width; P.h = depth->height;
    for (int k = 0; k < 9; ++k) P.K[k] = K[k];
    P.max_weight = max_weight;
    const bool pin = is_pinhole(K);
    const dim3 grid((unsigned)blocks), block(kIntThreads);
    if (vec == 4) {
        if (pin) k_integrate<true, 4><<<grid, block, 0, stream>>>(P);
        else k_integrate<false, 4><<<grid, block, 0, stream>>>(P);
    } else {
        if (pin) k_integrate<true, 1><<<grid, block, 0, stream>>>(P);
        else k_integrate<false, 1><<<grid, block, 0, stream>>>(P);
    }
    return launch_status();
}

}  // namespace emfb

extern "C" EMF_API int emf_integrate_volumes(int n_vol, const emf_volume* vols, const emf_pose* T_oc, const float K[9],
                                     const emf_image* depth, const emf_image* assoc, float max_weight,
                                     emf_stream_t stream) {
    return emfb::launch_integrate(n_vol, vols, T_oc, K, depth, assoc, max_weight, (cudaStream_t)stream);
}

extern "C" EMF_API int emf_update_tsdf(const emf_image* depth, const emf_image* assoc_weights, float* tsdf, float* weights,
                               const emf_pose* T_oc, const float K[9], const int res[3], float voxel_size,
                               float truncdist, float max_weight, emf_stream_t stream) {
    if (!res || !assoc_weights || !T_oc) return EMF_ERR_INVALID;
    emf_volume v;
    v.tsdf = tsdf; v.weights = weights; v.grads = nullptr; v.fg_probs = nullptr;
    v.res[0] = res[0]; v.res[1] = res[1]; v.res[2] = res[2];
    v.voxel_size = voxel_size; v.truncdist = truncdist; v.id = 0;
    return emfb::launch_integrate(1, &v, T_oc, K, depth, assoc_weights, max_weight, (cudaStream_t)stream);
}
