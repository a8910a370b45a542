// assoc.cu -- EM data association (E-step) for all volumes of a frame in ONE launch,
// plus the per-pixel producers/gathers next to it.
//
// Replaces, per frame and with K objects, the 14+17K OpenCV-CUDA launches of
// emf::EMFusion::computeAssociationWeights (reference src/core/EMFusion.cpp:635-670):
//   TSDF::computeAssociation / computeLaplace   src/core/TSDF.cpp:125-156
//   ObjTSDF::computeAssociation                  src/core/ObjTSDF.cpp:181-201
//   getVolumeVals                                src/core/cuda/TSDF.cu:662-726
//   the cross-volume normaliser                  src/core/EMFusion.cpp:653-665
//
// One thread owns one pixel and walks the volume table (kernel parameter, constant
// bank, warp-uniform).  Lanes of a warp are neighbouring pixels, so the 8-corner
// gathers of a warp fall into a handful of 32-byte sectors of the same volume.  The
// normaliser is accumulated in the reference's order (background, then objects in
// table order) in a register, and only the non-zero weights -- tracked in a bit
// set -- are divided in a second sweep over the thread's own outputs.
#include "common.cuh"

namespace emfb {

struct AssocVol {
    const float* tsdf;
    const float* fg_probs;   // nullable
    float* out; size_t out_pitch;
    uint8_t* mask_out; size_t mask_pitch;   // nullable (single-volume API)
    float R[9];              // T_CO
    float t[3];
    int rx, ry, rz;
    float voxel;
    float k1;                // -trunc / sigma   (src/core/TSDF.cpp:151)
};

struct AssocParams {
    AssocVol v[EMF_MAX_VOLUMES];
    int n_vol;
    int w, h;
    const float* points; size_t points_pitch;
    float k2;       // 1 / (2 sigma)            (src/core/TSDF.cpp:154)
    float alpha;    //                          (src/core/TSDF.cpp:131)
    float k3;       // (1 - alpha) * uniPrior   (src/core/TSDF.cpp:133)
    int mode;       // 0 normalise in place; 1 write partial normaliser; 2 raw (single volume); 3 as 1 without volume 0
    float* norm; size_t norm_pitch;
};

__device__ __forceinline__ float assoc_one(const AssocVol& V, const AssocParams& P, float px, float py, float pz,
                                           bool& invalid) {
    float f = 0.0f, fg = 1.0f;
    if (!(pz <= 0.0f)) {
        const float qx = fadd(V.t[0], dot_xyz(V.R[0], V.R[1], V.R[2], px, py, pz));
        const float qy = fadd(V.t[1], dot_xyz(V.R[3], V.R[4], V.R[5], px, py, pz));
        const float qz = fadd(V.t[2], dot_xyz(V.R[6], V.R[7], V.R[8], px, py, pz));
        const float vx = fadd(fmul((float)(V.rx - 1), 0.5f), fdiv(qx, V.voxel));
        const float vy = fadd(fmul((float)(V.ry - 1), 0.5f), fdiv(qy, V.voxel));
        const float vz = fadd(fmul((float)(V.rz - 1), 0.5f), fdiv(qz, V.voxel));
        if (!out_of(vx, vy, vz, 1.0f, (float)V.rx, (float)V.ry, (float)V.rz)) {
            f = trilinear(V.tsdf, V.rx, V.ry, vx, vy, vz);
            if (V.fg_probs && f != 0.0f) fg = trilinear(V.fg_probs, V.rx, V.ry, vx, vy, vz);
        }
    }
    invalid = (f == 0.0f);   // associationMask (src/core/TSDF.cpp:148)
    if (invalid) return 0.0f;
    float L = fmul(expf(fmul(fabsf(f), V.k1)), P.k2);
    if (V.fg_probs) L = fmul(L, fg);
    return fadd(fmul(L, P.alpha), P.k3);
}

__global__ void __launch_bounds__(256) k_assoc(const __grid_constant__ AssocParams P) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= P.w || y >= P.h) return;
    const float* pp = (const float*)((const char*)P.points + (size_t)y * P.points_pitch) + 3 * x;
    const float px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
    float n = 0.0f;
    uint32_t nz[(EMF_MAX_VOLUMES + 31) / 32] = {0, 0, 0};
    for (int i = 0; i < P.n_vol; ++i) {
        const AssocVol& V = P.v[i];
        bool invalid;
        const float wgt = assoc_one(V, P, px, py, pz, invalid);
        *((float*)((char*)V.out + (size_t)y * V.out_pitch) + x) = wgt;
        if (V.mask_out) V.mask_out[(size_t)y * V.mask_pitch + x] = invalid ? 255 : 0;
        if (P.mode == 3 && i == 0) { n = 0.0f; }      // a replica of the background: its weight is summed by its owner
        else n = (i == 0) ? wgt : fadd(n, wgt);      // copyTo, then add in table order (EMFusion.cpp:654-657)
        if (wgt != 0.0f) nz[i >> 5] |= 1u << (i & 31);
    }
    if (P.mode == 1 || P.mode == 3) { *((float*)((char*)P.norm + (size_t)y * P.norm_pitch) + x) = n; return; }
    if (P.mode != 0) return;
    if (P.norm) *((float*)((char*)P.norm + (size_t)y * P.norm_pitch) + x) = n;
#pragma unroll
    for (int wd = 0; wd < (EMF_MAX_VOLUMES + 31) / 32; ++wd) {
        uint32_t bits = nz[wd];
        while (bits) {
            const int i = wd * 32 + __ffs(bits) - 1;
            bits &= bits - 1;
            float* o = (float*)((char*)P.v[i].out + (size_t)y * P.v[i].out_pitch) + x;
            // cv::cuda::divide: x / 0 -> 0 (EMFusion.cpp:659-665)
            *o = (n != 0.0f) ? fdiv(*o, n) : 0.0f;
        }
    }
}

struct NormParams {
    float* img[EMF_MAX_VOLUMES]; size_t pitch[EMF_MAX_VOLUMES];
    int n_img, w, h;
    const float* norm; size_t norm_pitch;
};
__global__ void __launch_bounds__(256) k_assoc_normalise(const __grid_constant__ NormParams P) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= P.w || y >= P.h) return;
    const float n = *((const float*)((const char*)P.norm + (size_t)y * P.norm_pitch) + x);
    for (int i = 0; i < P.n_img; ++i) {
        float* o = (float*)((char*)P.img[i] + (size_t)y * P.pitch[i]) + x;
        const float v = *o;
        if (v != 0.0f) *o = (n != 0.0f) ? fdiv(v, n) : 0.0f;
    }
}

// Multi-GPU: the normaliser is the sum of every rank's partial sum.  The consumer kernel itself waits for the producers'
// flags (local memory; every CTA polls them, bounded by a time-out), then reads the partial images -- this rank's own and the
// peers', over NVLink -- adds them in rank order (the same bits on every rank), keeps the total and divides its images:
// the all-reduce and the normalisation (reference src/core/EMFusion.cpp:653-665) in one kernel.
struct NormPartsParams {
    float* img[EMF_MAX_VOLUMES]; size_t pitch[EMF_MAX_VOLUMES];
    int n_img, w, h;
    const float* part[16]; int n_parts;
    float* norm; size_t norm_pitch;
    const uint32_t* flags; uint32_t value; uint32_t* err; unsigned long long timeout_ns;
};
__global__ void __launch_bounds__(256) k_assoc_normalise_parts(const __grid_constant__ NormPartsParams P) {
    if (threadIdx.x < P.n_parts) {
        const volatile uint32_t* f = P.flags + threadIdx.x;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int32_t)(*f - P.value) < 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > P.timeout_ns) { atomicExch(P.err, 1u + threadIdx.x); break; }
            __nanosleep(100);
        }
        __threadfence_system();
    }
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= P.w || y >= P.h) return;
    const size_t i = (size_t)y * P.w + x;
    // (plain loads: the parts were written by other GPUs / earlier kernels, never through the read-only path)
    float n = *(const volatile float*)(P.part[0] + i);
    for (int r = 1; r < P.n_parts; ++r) n = fadd(n, *(const volatile float*)(P.part[r] + i));
    *((float*)((char*)P.norm + (size_t)y * P.norm_pitch) + x) = n;
    for (int k = 0; k < P.n_img; ++k) {
        float* o = (float*)((char*)P.img[k] + (size_t)y * P.pitch[k]) + x;
        const float v = *o;
        if (v != 0.0f) *o = (n != 0.0f) ? fdiv(v, n) : 0.0f;
    }
}

// getVolumeVals<float>
__global__ void __launch_bounds__(256) k_gather(const float* __restrict__ vol, Img<const float> points, Img<float> vals,
                                                const __grid_constant__ Pose T, int rx, int ry, int rz, float voxel) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= points.w || y >= points.h) return;
    const float* pp = points.row(y) + 3 * x;
    const float px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
    float f = 0.0f;
    if (!(pz <= 0.0f)) {
        const float qx = fadd(T.t[0], dot_xyz(T.R[0], T.R[1], T.R[2], px, py, pz));
        const float qy = fadd(T.t[1], dot_xyz(T.R[3], T.R[4], T.R[5], px, py, pz));
        const float qz = fadd(T.t[2], dot_xyz(T.R[6], T.R[7], T.R[8], px, py, pz));
        const float vx = fadd(fmul((float)(rx - 1), 0.5f), fdiv(qx, voxel));
        const float vy = fadd(fmul((float)(ry - 1), 0.5f), fdiv(qy, voxel));
        const float vz = fadd(fmul((float)(rz - 1), 0.5f), fdiv(qz, voxel));
        if (!out_of(vx, vy, vz, 1.0f, (float)rx, (float)ry, (float)rz)) f = trilinear(vol, rx, ry, vx, vy, vz);
    }
    vals.at(y, x) = f;
}

// computePoints (reference src/core/cuda/EMFusion.cu:29-47)
__global__ void __launch_bounds__(256) k_points(Img<const float> depth, Img<float> points, float fx, float fy, float cx,
                                                float cy) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= depth.w || y >= depth.h) return;
    const float d = __ldg(depth.row(y) + x);
    float* p = points.row(y) + 3 * x;
    p[0] = fdiv(fmul(fsub((float)x, cx), d), fx);
    p[1] = fdiv(fmul(fsub((float)y, cy), d), fy);
    p[2] = d;
}

__global__ void __launch_bounds__(256) k_fill_f32(Img<float> im, float v) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < im.w && y < im.h) im.at(y, x) = v;
}

static int fill_assoc_vol(AssocVol& d, const emf_volume& v, const emf_pose& T, const emf_image* out,
                          const emf_image* mask_out, const emf_tsdf_params& prm, int w, int h) {
    if (!v.tsdf || !res_ok(v.res) || !image_ok(out, 4) || out->width != w || out->height != h) return EMF_ERR_INVALID;
    if (mask_out && (!image_ok(mask_out, 1) || mask_out->width != w || mask_out->height != h)) return EMF_ERR_INVALID;
    d.tsdf = v.tsdf; d.fg_probs = v.fg_probs;
    d.out = (float*)out->ptr; d.out_pitch = out->pitch;
    d.mask_out = mask_out ? (uint8_t*)mask_out->ptr : nullptr; d.mask_pitch = mask_out ? mask_out->pitch : 0;
    for (int k = 0; k < 9; ++k) d.R[k] = T.R[k];
    for (int k = 0; k < 3; ++k) d.t[k] = T.t[k];
    d.rx = v.res[0]; d.ry = v.res[1]; d.rz = v.res[2];
    d.voxel = v.voxel_size;
    d.k1 = -v.truncdist / prm.assoc_sigma;
    return EMF_OK;
}

static void fill_assoc_consts(AssocParams& P, const emf_tsdf_params& prm) {
    P.k2 = 1.f / (2.f * prm.assoc_sigma);
    P.alpha = prm.alpha;
    P.k3 = (1 - prm.alpha) * prm.uni_prior;
}

}  // namespace emfb

using namespace emfb;

extern "C" EMF_API int emf_assoc_weights(int n_vol, const emf_volume* vols, const emf_pose* T_co, const emf_image* points,
                                 const emf_tsdf_params* params, const emf_image* assoc_out, int mode,
                                 const emf_image* norm_partial, emf_stream_t stream) {
    if (n_vol <= 0 || !vols || !T_co || !params || !assoc_out || !image_ok(points, 12)) return EMF_ERR_INVALID;
    if (n_vol > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    if (mode != 0 && mode != 1 && mode != 3) return EMF_ERR_INVALID;
    if (mode != 0 && !image_ok(norm_partial, 4)) return EMF_ERR_INVALID;
    AssocParams P;
    const int w = points->width, h = points->height;
    for (int i = 0; i < n_vol; ++i) {
        const int rc = fill_assoc_vol(P.v[i], vols[i], T_co[i], &assoc_out[i], nullptr, *params, w, h);
        if (rc != EMF_OK) return rc;
    }
    P.n_vol = n_vol; P.w = w; P.h = h;
    P.points = (const float*)points->ptr; P.points_pitch = points->pitch;
    fill_assoc_consts(P, *params);
    P.mode = mode;
    P.norm = norm_partial ? (float*)norm_partial->ptr : nullptr;
    P.norm_pitch = norm_partial ? norm_partial->pitch : 0;
    if (norm_partial && (norm_partial->width != w || norm_partial->height != h)) return EMF_ERR_INVALID;
    const dim3 grid((w + 31) / 32, (h + 7) / 8);
    k_assoc<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}

extern "C" EMF_API int emf_compute_association(const emf_volume* vol, const emf_image* points, const emf_pose* T_co,
                                       const emf_tsdf_params* params, const emf_image* assoc_out,
                                       const emf_image* assoc_mask_out, emf_stream_t stream) {
    if (!vol || !T_co || !params || !image_ok(points, 12)) return EMF_ERR_INVALID;
    AssocParams P;
    const int w = points->width, h = points->height;
    const int rc = fill_assoc_vol(P.v[0], *vol, *T_co, assoc_out, assoc_mask_out, *params, w, h);
    if (rc != EMF_OK) return rc;
    P.n_vol = 1; P.w = w; P.h = h;
    P.points = (const float*)points->ptr; P.points_pitch = points->pitch;
    fill_assoc_consts(P, *params);
    P.mode = 2; P.norm = nullptr; P.norm_pitch = 0;
    const dim3 grid((w + 31) / 32, (h + 7) / 8);
    k_assoc<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}

extern "C" EMF_API int emf_assoc_normalise(int n_img, const emf_image* assoc_io, const emf_image* norm, emf_stream_t stream) {
    if (n_img <= 0 || !assoc_io || !image_ok(norm, 4)) return EMF_ERR_INVALID;
    if (n_img > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    NormParams P;
    for (int i = 0; i < n_img; ++i) {
        if (!image_ok(&assoc_io[i], 4) || !same_size(&assoc_io[i], norm)) return EMF_ERR_INVALID;
        P.img[i] = (float*)assoc_io[i].ptr; P.pitch[i] = assoc_io[i].pitch;
    }
    P.n_img = n_img; P.w = norm->width; P.h = norm->height;
    P.norm = (const float*)norm->ptr; P.norm_pitch = norm->pitch;
    const dim3 grid((P.w + 31) / 32, (P.h + 7) / 8);
    k_assoc_normalise<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}

extern "C" EMF_API int emf_assoc_normalise_parts(int n_img, const emf_image* assoc_io, int n_parts, const float* const* parts,
                                                 const emf_image* norm_out, const uint32_t* flags, uint32_t value, uint32_t* err,
                                                 double timeout_s, emf_stream_t stream) {
    if (n_img < 0 || (n_img > 0 && !assoc_io) || n_parts <= 0 || n_parts > 16 || !parts || !image_ok(norm_out, 4) || !flags || !err ||
        !(timeout_s > 0.0))
        return EMF_ERR_INVALID;
    if (n_img > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    NormPartsParams P;
    for (int i = 0; i < n_img; ++i) {
        if (!image_ok(&assoc_io[i], 4) || !same_size(&assoc_io[i], norm_out)) return EMF_ERR_INVALID;
        P.img[i] = (float*)assoc_io[i].ptr; P.pitch[i] = assoc_io[i].pitch;
    }
    for (int r = 0; r < n_parts; ++r) { if (!parts[r]) return EMF_ERR_INVALID; P.part[r] = parts[r]; }
    P.n_img = n_img; P.n_parts = n_parts; P.w = norm_out->width; P.h = norm_out->height;
    P.norm = (float*)norm_out->ptr; P.norm_pitch = norm_out->pitch;
    P.flags = flags; P.value = value; P.err = err; P.timeout_ns = (unsigned long long)(timeout_s * 1e9);
    const dim3 grid((P.w + 31) / 32, (P.h + 7) / 8);
    k_assoc_normalise_parts<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}

extern "C" EMF_API int emf_get_volume_vals(const float* vol, const emf_image* points, const emf_pose* T_co, const int res[3],
                                   float voxel_size, const emf_image* vals, emf_stream_t stream) {
    if (!vol || !T_co || !res_ok(res) || !image_ok(points, 12) || !image_ok(vals, 4) || !same_size(points, vals))
        return EMF_ERR_INVALID;
    const dim3 grid((points->width + 31) / 32, (points->height + 7) / 8);
    k_gather<<<grid, 256, 0, (cudaStream_t)stream>>>(vol, view<const float>(points), view<float>(vals), to_pose(T_co),
                                                     res[0], res[1], res[2], voxel_size);
    return launch_status();
}

extern "C" EMF_API int emf_compute_points(const emf_image* depth, const emf_image* points, const float K[9],
                                  emf_stream_t stream) {
    if (!K || !image_ok(depth, 4) || !image_ok(points, 12) || !same_size(depth, points)) return EMF_ERR_INVALID;
    const dim3 grid((depth->width + 31) / 32, (depth->height + 7) / 8);
    k_points<<<grid, 256, 0, (cudaStream_t)stream>>>(view<const float>(depth), view<float>(points), K[0], K[4], K[2],
                                                     K[5]);
    return launch_status();
}

extern "C" EMF_API int emf_fill_image_f32(const emf_image* img, float value, emf_stream_t stream) {
    if (!image_ok(img, 4)) return EMF_ERR_INVALID;
    const dim3 grid((img->width + 31) / 32, (img->height + 7) / 8);
    k_fill_f32<<<grid, 256, 0, (cudaStream_t)stream>>>(view<float>(img), value);
    return launch_status();
}
