// fgprob.cu -- foreground/background probability volumes of object TSDFs.
//
// Replaces emf::cuda::ObjTSDF::updateFgBgProbs (reference src/core/cuda/ObjTSDF.cu:29-107)
// and the six OpenCV launches of emf::ObjTSDF::computeFgProbs (src/core/ObjTSDF.cpp:218-226).
// Their product, fgProbs, is an input of the association and raycast kernels.
#include "common.cuh"

namespace emfb {

__global__ void __launch_bounds__(256) k_update_fgbg(Img<const uint8_t> mask, Img<const uint8_t> occluded,
                                                     const float* __restrict__ tsdf, const float* __restrict__ weights,
                                                     float2* __restrict__ fgbg, const __grid_constant__ Pose T,
                                                     const __grid_constant__ Intr I, int rx, int ry, int rz, float s) {
    const int64_t n = (int64_t)rx * ry * rz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float tv = __ldg(tsdf + i), wv = __ldg(weights + i);
        // only seen voxels within the truncation band (ObjTSDF.cu:49)
        if (fabsf(tv) >= 1.0f || wv == 0.0f) continue;
        const int64_t row = i / rx;
        const int x = (int)(i - row * rx);
        const int z = (int)(row / ry);
        const int y = (int)(row - (int64_t)z * ry);
        const float cx = fmul(fsub((float)x, fmul((float)(rx - 1), 0.5f)), s);
        const float cy = fmul(fsub((float)y, fmul((float)(ry - 1), 0.5f)), s);
        const float cz = fmul(fsub((float)z, fmul((float)(rz - 1), 0.5f)), s);
        const float pcx = fadd(T.t[0], dot_yxz(T.R[0], T.R[1], T.R[2], cx, cy, cz));
        const float pcy = fadd(T.t[1], dot_yxz(T.R[3], T.R[4], T.R[5], cx, cy, cz));
        const float pcz = fadd(T.t[2], dot_yxz(T.R[6], T.R[7], T.R[8], cx, cy, cz));
        if (pcz <= 0.0f) continue;
        const float qx = dot_yxz(I.K[0], I.K[1], I.K[2], pcx, pcy, pcz);
        const float qy = dot_yxz(I.K[3], I.K[4], I.K[5], pcx, pcy, pcz);
        const float qz = dot_yxz(I.K[6], I.K[7], I.K[8], pcx, pcy, pcz);
        const int px = __float2int_rn(fdiv(qx, qz)), py = __float2int_rn(fdiv(qy, qz));
        if (px < 0 || px >= mask.w || py < 0 || py >= mask.h) continue;
        if (!occluded.at(py, px)) {
            const int m = mask.at(py, px) ? 1 : 0;
            float2 p = fgbg[i];
            p.x = fadd(p.x, (float)m);
            p.y = fadd(p.y, (float)(1 - m));
            fgbg[i] = p;
        }
    }
}

__global__ void __launch_bounds__(256) k_fg_probs(const float2* __restrict__ fgbg, int64_t n, float* __restrict__ fg,
                                                  uint8_t* __restrict__ vol_mask) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 p = fgbg[i];
        const float s = fadd(p.x, p.y);
        float v = (s != 0.0f) ? fdiv(p.x, s) : 0.0f;   // guarded cv::cuda::divide
        if (v != v) v = 0.0f;                            // NaN patch (ObjTSDF.cpp:223-224)
        fg[i] = v;
        if (vol_mask) vol_mask[i] = (v > 0.5f) ? 255 : 0;
    }
}

// computeFgProbs + inclusive voxel bounds of {fgProb > 0.5} (emf_volume::fg_box).  box must hold
// (INT_MAX, INT_MAX, INT_MAX, -1, -1, -1) on entry (k_fg_box_init on the same stream).
__global__ void k_fg_box_init(int32_t* box) {
    if (threadIdx.x < 3) box[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) box[threadIdx.x] = -1;
}
__global__ void __launch_bounds__(256) k_fg_probs_box(const float2* __restrict__ fgbg, int rx, int ry, int rz, float* __restrict__ fg,
                                                      uint8_t* __restrict__ vol_mask, int32_t* __restrict__ box) {
    const int64_t n = (int64_t)rx * ry * rz;
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-1, -1, -1};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 p = fgbg[i];
        const float s = fadd(p.x, p.y);
        float v = (s != 0.0f) ? fdiv(p.x, s) : 0.0f;   // guarded cv::cuda::divide
        if (v != v) v = 0.0f;                            // NaN patch (ObjTSDF.cpp:223-224)
        fg[i] = v;
        if (vol_mask) vol_mask[i] = (v > 0.5f) ? 255 : 0;
        if (v > 0.5f) {
            const int64_t row = i / rx;
            const int x = (int)(i - row * rx), z = (int)(row / ry), y = (int)(row - (int64_t)z * ry);
            lo[0] = min(lo[0], x); lo[1] = min(lo[1], y); lo[2] = min(lo[2], z);
            hi[0] = max(hi[0], x); hi[1] = max(hi[1], y); hi[2] = max(hi[2], z);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
        hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
    }
    if ((threadIdx.x & 31) == 0 && hi[0] >= 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(box + k, lo[k]); atomicMax(box + 3 + k, hi[k]); }
    }
}

}  // namespace emfb

using namespace emfb;

static unsigned grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)(b < 1 ? 1 : b);
}

extern "C" EMF_API int emf_update_fgbg_probs(const emf_image* mask, const emf_image* occluded, const float* tsdf,
                                     const float* weights, float* fgbg, const emf_pose* T_oc, const float K[9],
                                     const int res[3], float voxel_size, emf_stream_t stream) {
    if (!image_ok(mask, 1) || !image_ok(occluded, 1) || !same_size(mask, occluded) || !tsdf || !weights || !fgbg ||
        !T_oc || !K || !res_ok(res))
        return EMF_ERR_INVALID;
    const int64_t n = (int64_t)res[0] * res[1] * res[2];
    k_update_fgbg<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(view<const uint8_t>(mask), view<const uint8_t>(occluded),
                                                                tsdf, weights, (float2*)fgbg, to_pose(T_oc), to_intr(K),
                                                                res[0], res[1], res[2], voxel_size);
    return launch_status();
}

extern "C" EMF_API int emf_compute_fg_probs(const float* fgbg, int64_t n_voxels, float* fg_probs, uint8_t* fg_vol_mask,
                                    emf_stream_t stream) {
    if (!fgbg || !fg_probs || n_voxels <= 0) return EMF_ERR_INVALID;
    k_fg_probs<<<grid_for(n_voxels), 256, 0, (cudaStream_t)stream>>>((const float2*)fgbg, n_voxels, fg_probs, fg_vol_mask);
    return launch_status();
}

extern "C" EMF_API int emf_compute_fg_probs_box(const float* fgbg, const int res[3], float* fg_probs, uint8_t* fg_vol_mask,
                                        int32_t* fg_box, emf_stream_t stream) {
    if (!fgbg || !fg_probs || !fg_box || !res_ok(res)) return EMF_ERR_INVALID;
    const int64_t n = (int64_t)res[0] * res[1] * res[2];
    k_fg_box_init<<<1, 32, 0, (cudaStream_t)stream>>>(fg_box);
    k_fg_probs_box<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float2*)fgbg, res[0], res[1], res[2], fg_probs,
                                                                fg_vol_mask, fg_box);
    return launch_status();
}
