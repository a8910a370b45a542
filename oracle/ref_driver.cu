// ref_driver.cu -- C entry points around the UNMODIFIED reference kernels.
//
// TEST INFRASTRUCTURE ONLY (oracle/).  This file is compiled together with
// /root/reference/src/core/cuda/{TSDF,ObjTSDF}.cu (in place, never copied; see
// oracle/Makefile) against the type shim in oracle/shim/, into
// oracle/_ref/libemf_ref.so.  It is the authoritative oracle for "matches the
// reference's own CUDA path" and the timed reference arm of bench.py.
//
// Three layers:
//  1. emfref_<op>: thin wrappers that build GpuMat headers over caller memory and
//     call the reference's level-1 operators (emf::cuda::TSDF::*, ::ObjTSDF::*).
//  2. cvk_*: one-kernel-per-call restatements of the OpenCV-CUDA element-wise ops
//     the reference host code chains around those kernels (cudaarithm is an
//     un-vendored dependency; SURVEY.md section 8c lists the ops and semantics).
//  3. emfref_frame_*: restatement of the three hot methods of emf::EMFusion
//     (src/core/EMFusion.cpp:635-670, 726-795, 865-889) and of the TSDF/ObjTSDF
//     methods they call (src/core/TSDF.cpp:108-168, src/core/ObjTSDF.cpp:181-226)
//     with the reference's launch structure: one launch per op per volume, one
//     stream per volume, host barriers where the reference calls waitForCompletion,
//     countNonZero as a blocking device->host read.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <set>
#include <cvshim.hpp>
#include "EMFusion/core/cuda/TSDF.cuh"
#include "EMFusion/core/cuda/ObjTSDF.cuh"

using cv::cuda::GpuMat;
using cv::cuda::Stream;

#define REF_API extern "C" __attribute__((visibility("default")))

static GpuMat mat(void* p, int rows, int cols, int type) {
    static const int sz[] = {1, 1, 2, 2, 4, 4, 8};
    const size_t es = (size_t)sz[CV_MAT_DEPTH(type)] * CV_MAT_CN(type);
    return GpuMat(rows, cols, type, p, (size_t)cols * es);
}
static cv::Matx33f m33(const float* v) { cv::Matx33f m; for (int i = 0; i < 9; ++i) m.val[i] = v[i]; return m; }
static cv::Vec3f v3(const float* v) { cv::Vec3f m; for (int i = 0; i < 3; ++i) m.val[i] = v[i]; return m; }
static cv::Vec3i v3i(const int* v) { cv::Vec3i m; for (int i = 0; i < 3; ++i) m.val[i] = v[i]; return m; }
static int status() { return cudaPeekAtLastError() == cudaSuccess ? 0 : -2; }

// ------------------------------------------------------------------------------------------------
// 1. level-1 wrappers (continuous buffers)
// ------------------------------------------------------------------------------------------------
REF_API int emfref_update_tsdf(const float* depth, const float* assoc, int w, int h, float* tsdf, float* weights,
                               const float* R, const float* t, const float* K, const int* res, float voxel,
                               float trunc, float maxw, cudaStream_t s) {
    Stream st(s);
    GpuMat d = mat((void*)depth, h, w, CV_32FC1), a = mat((void*)assoc, h, w, CV_32FC1);
    GpuMat tv = mat(tsdf, res[1] * res[2], res[0], CV_32FC1), wv = mat(weights, res[1] * res[2], res[0], CV_32FC1);
    emf::cuda::TSDF::updateTSDF(d, a, tv, wv, m33(R), v3(t), m33(K), v3i(res), voxel, trunc, maxw, st);
    return status();
}

// TSDF::updateGradients (src/core/TSDF.cpp:120-123): setTo(0) + computeTSDFGrads
REF_API int emfref_update_gradients(const float* tsdf, float* grads, const int* res, cudaStream_t s) {
    Stream st(s);
    GpuMat tv = mat((void*)tsdf, res[1] * res[2], res[0], CV_32FC1);
    GpuMat g = mat(grads, res[1] * res[2], res[0], CV_32FC3);
    g.setTo(cv::Scalar::all(0.f), st);
    emf::cuda::TSDF::computeTSDFGrads(tv, g, v3i(res), st);
    return status();
}

REF_API int emfref_raycast(const float* tsdf, const float* grads, const float* weights, float* ray, float* vert,
                           float* norm, uint8_t* mask, int w, int h, const float* R, const float* t, const float* K,
                           const int* res, float voxel, float trunc, cudaStream_t s) {
    Stream st(s);
    GpuMat tv = mat((void*)tsdf, res[1] * res[2], res[0], CV_32FC1);
    GpuMat gv = mat((void*)grads, res[1] * res[2], res[0], CV_32FC3);
    GpuMat wv = mat((void*)weights, res[1] * res[2], res[0], CV_32FC1);
    GpuMat r = mat(ray, h, w, CV_32FC1), v = mat(vert, h, w, CV_32FC3), n = mat(norm, h, w, CV_32FC3),
           m = mat(mask, h, w, CV_8UC1);
    emf::cuda::TSDF::raycastTSDF(tv, gv, wv, r, v, n, m, m33(R), v3(t), m33(K), v3i(res), voxel, trunc, st);
    return status();
}

REF_API int emfref_get_volume_vals(const float* vol, const float* points, int w, int h, const float* R,
                                   const float* t, const int* res, float voxel, float* vals, cudaStream_t s) {
    Stream st(s);
    GpuMat vv = mat((void*)vol, res[1] * res[2], res[0], CV_32FC1);
    GpuMat p = mat((void*)points, h, w, CV_32FC3), o = mat(vals, h, w, CV_32FC1);
    emf::cuda::TSDF::getVolumeVals(vv, p, m33(R), v3(t), v3i(res), voxel, o, st);
    return status();
}

REF_API int emfref_update_fgbg(const uint8_t* mask, const uint8_t* occluded, int w, int h, const float* tsdf,
                               const float* weights, float* fgbg, const float* R, const float* t, const float* K,
                               const int* res, float voxel, cudaStream_t s) {
    Stream st(s);
    GpuMat m = mat((void*)mask, h, w, CV_8UC1), o = mat((void*)occluded, h, w, CV_8UC1);
    GpuMat tv = mat((void*)tsdf, res[1] * res[2], res[0], CV_32FC1);
    GpuMat wv = mat((void*)weights, res[1] * res[2], res[0], CV_32FC1);
    GpuMat f = mat(fgbg, res[1] * res[2], res[0], CV_32FC2);
    emf::cuda::ObjTSDF::updateFgBgProbs(m, o, tv, wv, f, m33(R), v3(t), m33(K), v3i(res), voxel, st);
    return status();
}

// ------------------------------------------------------------------------------------------------
// 2. OpenCV-CUDA element-wise ops, one launch each (cudaarithm semantics, fp32)
// ------------------------------------------------------------------------------------------------
namespace cvk {
constexpr int T = 256;
static inline unsigned nb(size_t n) { return (unsigned)((n + T - 1) / T); }
#define IDX size_t i = (size_t)blockIdx.x * T + threadIdx.x; if (i >= n) return;

__global__ void k_set_f(float* d, size_t n, float v) { IDX d[i] = v; }
__global__ void k_set_u8(uint8_t* d, size_t n, uint8_t v) { IDX d[i] = v; }
__global__ void k_set_f_masked(float* d, size_t n, float v, const uint8_t* m) { IDX if (m[i]) d[i] = v; }
__global__ void k_set_u8_masked(uint8_t* d, size_t n, uint8_t v, const uint8_t* m) { IDX if (m[i]) d[i] = v; }
__global__ void k_cmp_eq_s(const float* a, size_t n, float v, uint8_t* d) { IDX d[i] = a[i] == v ? 255 : 0; }
__global__ void k_cmp_le_s(const float* a, size_t n, float v, uint8_t* d) { IDX d[i] = a[i] <= v ? 255 : 0; }
__global__ void k_cmp_gt_s(const float* a, size_t n, float v, uint8_t* d) { IDX d[i] = a[i] > v ? 255 : 0; }
__global__ void k_cmp_lt(const float* a, const float* b, size_t n, uint8_t* d) { IDX d[i] = a[i] < b[i] ? 255 : 0; }
__global__ void k_cmp_eq_u8s(const uint8_t* a, size_t n, int v, uint8_t* d) { IDX d[i] = (int)a[i] == v ? 255 : 0; }
__global__ void k_abs(float* a, size_t n) { IDX a[i] = fabsf(a[i]); }
__global__ void k_mul_s(const float* a, size_t n, float v, float* d) { IDX d[i] = __fmul_rn(a[i], v); }
__global__ void k_mul(const float* a, const float* b, size_t n, float* d) { IDX d[i] = __fmul_rn(a[i], b[i]); }
__global__ void k_exp(float* a, size_t n) { IDX a[i] = expf(a[i]); }
__global__ void k_add_s(float* a, size_t n, float v) { IDX a[i] = __fadd_rn(a[i], v); }
__global__ void k_add(const float* a, const float* b, size_t n, float* d) { IDX d[i] = __fadd_rn(a[i], b[i]); }
__global__ void k_div(const float* a, const float* b, size_t n, float* d) { IDX d[i] = b[i] != 0.f ? __fdiv_rn(a[i], b[i]) : 0.f; }
__global__ void k_sub_masked(const float* a, const float* b, size_t n, float* d, const uint8_t* m) { IDX if (m[i]) d[i] = __fsub_rn(a[i], b[i]); }
__global__ void k_or(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* d) { IDX d[i] = a[i] | b[i]; }
__global__ void k_and(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* d) { IDX d[i] = a[i] & b[i]; }
__global__ void k_copy_f(const float* a, size_t n, float* d) { IDX d[i] = a[i]; }
__global__ void k_copy_f_masked(const float* a, size_t n, float* d, const uint8_t* m) { IDX if (m[i]) d[i] = a[i]; }
__global__ void k_copy_f3_masked(const float* a, size_t n, float* d, const uint8_t* m) {
    IDX if (m[i]) { d[3 * i] = a[3 * i]; d[3 * i + 1] = a[3 * i + 1]; d[3 * i + 2] = a[3 * i + 2]; }
}
__global__ void k_count_rect(const uint8_t* a, int w, int x0, int y0, int rw, int rh, int* out) {
    size_t n = (size_t)rw * rh;
    IDX const int x = x0 + (int)(i % rw), y = y0 + (int)(i / rw);
    if (a[(size_t)y * w + x]) atomicAdd(out, 1);
}
#undef IDX
}  // namespace cvk

// ------------------------------------------------------------------------------------------------
// 3. frame-level restatement of the reference host code
// ------------------------------------------------------------------------------------------------
struct RefVolume {           // emf::TSDF / emf::ObjTSDF device state
    float *tsdf, *weights, *grads, *fg_probs;   // caller memory (fg_probs == nullptr: background)
    int res[3];
    float voxel, trunc;
    int id;
    // owned scratch / outputs
    float *tmpAssoc = nullptr, *fgProbVals = nullptr, *raycastWeights = nullptr;
    uint8_t *assocMask = nullptr, *fgVolMask = nullptr;
    float *assoc = nullptr;                     // associationWeights[id] / bg_associationWeights
    float *ray = nullptr, *vert = nullptr, *norm = nullptr; uint8_t* seg = nullptr;   // obj_raylengths ... obj_modelSegmentation
    cudaStream_t stream = nullptr;
    size_t nvox() const { return (size_t)res[0] * res[1] * res[2]; }
};

struct RefFrame {
    int w, h;
    size_t n;                 // pixels
    std::vector<RefVolume> vols;   // [0] = background, then objects in list (= id) order
    // EMFusion members
    float *raylengths, *vertices, *normals, *associationNorm, *diffRaylengths;
    uint8_t *modelSegmentation, *mask, *obj_mask, *takeBgMask, *noObjMask;
    int* d_count;
    std::set<int> vis_objs;
    std::vector<int> vis_flags;
};

template <typename T> static T* dalloc(size_t n) { void* p = nullptr; cudaMalloc(&p, n * sizeof(T)); cudaMemset(p, 0, n * sizeof(T)); return (T*)p; }

struct emfref_volume_desc {
    float* tsdf; float* weights; float* grads; float* fg_probs;
    int res[3]; float voxel; float trunc; int id;
};

REF_API void* emfref_frame_create(int w, int h, int n_vol, const emfref_volume_desc* d) {
    RefFrame* F = new RefFrame();
    F->w = w; F->h = h; F->n = (size_t)w * h;
    for (int i = 0; i < n_vol; ++i) {
        RefVolume v;
        v.tsdf = d[i].tsdf; v.weights = d[i].weights; v.grads = d[i].grads; v.fg_probs = d[i].fg_probs;
        for (int k = 0; k < 3; ++k) v.res[k] = d[i].res[k];
        v.voxel = d[i].voxel; v.trunc = d[i].trunc; v.id = d[i].id;
        v.tmpAssoc = dalloc<float>(F->n); v.assocMask = dalloc<uint8_t>(F->n); v.assoc = dalloc<float>(F->n);
        v.ray = dalloc<float>(F->n); v.vert = dalloc<float>(3 * F->n); v.norm = dalloc<float>(3 * F->n);
        v.seg = dalloc<uint8_t>(F->n);
        if (v.fg_probs) {
            v.fgProbVals = dalloc<float>(F->n);
            v.raycastWeights = dalloc<float>(v.nvox());
            v.fgVolMask = dalloc<uint8_t>(v.nvox());
        }
        cudaStreamCreate(&v.stream);   // cv::cuda::Stream(): blocking w.r.t. the legacy default stream
        F->vols.push_back(v);
    }
    F->raylengths = dalloc<float>(F->n); F->vertices = dalloc<float>(3 * F->n); F->normals = dalloc<float>(3 * F->n);
    F->associationNorm = dalloc<float>(F->n); F->diffRaylengths = dalloc<float>(F->n);
    F->modelSegmentation = dalloc<uint8_t>(F->n); F->mask = dalloc<uint8_t>(F->n); F->obj_mask = dalloc<uint8_t>(F->n);
    F->takeBgMask = dalloc<uint8_t>(F->n); F->noObjMask = dalloc<uint8_t>(F->n);
    F->d_count = dalloc<int>(1);
    F->vis_flags.assign(n_vol, 1);
    cudaDeviceSynchronize();
    return F;
}

REF_API void emfref_frame_destroy(void* h) {
    RefFrame* F = (RefFrame*)h;
    for (auto& v : F->vols) {
        cudaFree(v.tmpAssoc); cudaFree(v.assocMask); cudaFree(v.assoc); cudaFree(v.ray); cudaFree(v.vert);
        cudaFree(v.norm); cudaFree(v.seg); cudaFree(v.fgProbVals); cudaFree(v.raycastWeights); cudaFree(v.fgVolMask);
        cudaStreamDestroy(v.stream);
    }
    cudaFree(F->raylengths); cudaFree(F->vertices); cudaFree(F->normals); cudaFree(F->associationNorm);
    cudaFree(F->diffRaylengths); cudaFree(F->modelSegmentation); cudaFree(F->mask); cudaFree(F->obj_mask);
    cudaFree(F->takeBgMask); cudaFree(F->noObjMask); cudaFree(F->d_count);
    delete F;
}

// fgVolMask = fgProbs > 0.5 (last line of ObjTSDF::computeFgProbs, src/core/ObjTSDF.cpp:225); call after fg_probs change
REF_API int emfref_frame_update_fg_masks(void* h) {
    RefFrame* F = (RefFrame*)h;
    for (auto& v : F->vols)
        if (v.fg_probs) cvk::k_cmp_gt_s<<<cvk::nb(v.nvox()), cvk::T, 0, v.stream>>>(v.fg_probs, v.nvox(), 0.5f, v.fgVolMask);
    cudaDeviceSynchronize();
    return status();
}

static void wait_all(RefFrame* F) { for (auto& v : F->vols) cudaStreamSynchronize(v.stream); }

REF_API int emfref_frame_assoc(void* h, const float* points, const float* R_co /*n_vol*9*/, const float* t_co /*n_vol*3*/,
                               float sigma, float alpha, float uni) {
    using namespace cvk;
    RefFrame* F = (RefFrame*)h;
    const size_t n = F->n;
    const int nv = (int)F->vols.size();
    // EMFusion::computeAssociationWeights, src/core/EMFusion.cpp:635-670
    for (auto& v : F->vols) cudaMemsetAsync(v.assoc, 0, n * 4, v.stream);       // :636-639
    for (int i = 0; i < nv; ++i) {                                              // :642-647
        RefVolume& v = F->vols[i];
        cudaStream_t s = v.stream; Stream st(s);
        const float* R = R_co + 9 * i; const float* t = t_co + 3 * i;
        // --- computeLaplace, src/core/TSDF.cpp:138-156
        cudaMemsetAsync(v.tmpAssoc, 0, n * 4, s);
        GpuMat vv = mat(v.tsdf, v.res[1] * v.res[2], v.res[0], CV_32FC1);
        GpuMat p = mat((void*)points, F->h, F->w, CV_32FC3), o = mat(v.tmpAssoc, F->h, F->w, CV_32FC1);
        emf::cuda::TSDF::getVolumeVals(vv, p, m33(R), v3(t), v3i(v.res), v.voxel, o, st);
        k_cmp_eq_s<<<nb(n), T, 0, s>>>(v.tmpAssoc, n, 0.f, v.assocMask);
        k_abs<<<nb(n), T, 0, s>>>(v.tmpAssoc, n);
        k_mul_s<<<nb(n), T, 0, s>>>(v.tmpAssoc, n, -v.trunc / sigma, v.tmpAssoc);
        k_exp<<<nb(n), T, 0, s>>>(v.tmpAssoc, n);
        k_mul_s<<<nb(n), T, 0, s>>>(v.tmpAssoc, n, 1.f / (2.f * sigma), v.tmpAssoc);
        if (v.fg_probs) {                                                       // src/core/ObjTSDF.cpp:189-194
            GpuMat fv = mat(v.fg_probs, v.res[1] * v.res[2], v.res[0], CV_32FC1), fo = mat(v.fgProbVals, F->h, F->w, CV_32FC1);
            emf::cuda::TSDF::getVolumeVals(fv, p, m33(R), v3(t), v3i(v.res), v.voxel, fo, st);
            k_mul<<<nb(n), T, 0, s>>>(v.tmpAssoc, v.fgProbVals, n, v.tmpAssoc);
        }
        // --- computeAssociation tail, src/core/TSDF.cpp:131-135
        k_mul_s<<<nb(n), T, 0, s>>>(v.tmpAssoc, n, alpha, v.assoc);
        k_add_s<<<nb(n), T, 0, s>>>(v.assoc, n, (1 - alpha) * uni);
        k_set_f_masked<<<nb(n), T, 0, s>>>(v.assoc, n, 0.f, v.assocMask);
    }
    wait_all(F);                                                                // :649-651
    k_copy_f<<<nb(n), T>>>(F->vols[0].assoc, n, F->associationNorm);            // :654 (default stream)
    for (int i = 1; i < nv; ++i) k_add<<<nb(n), T>>>(F->associationNorm, F->vols[i].assoc, n, F->associationNorm);
    for (auto& v : F->vols) k_div<<<nb(n), T, 0, v.stream>>>(v.assoc, F->associationNorm, n, v.assoc);   // :659-665
    wait_all(F);                                                                // :667-669
    return status();
}

REF_API int emfref_frame_raycast(void* h, const float* R_co, const float* t_co, const float* K, int boundary,
                                 int visibility_thresh) {
    using namespace cvk;
    RefFrame* F = (RefFrame*)h;
    const size_t n = F->n;
    const int nv = (int)F->vols.size();
    cudaStream_t s0 = F->vols[0].stream;
    // EMFusion::raycast, src/core/EMFusion.cpp:726-795.  The reference clears the composite
    // outputs on streams[objects.size()] (:727-733); here that is the last volume's stream.
    cudaStream_t sl = F->vols[nv - 1].stream;
    cudaMemsetAsync(F->raylengths, 0, n * 4, sl);
    cudaMemsetAsync(F->vols[0].ray, 0, n * 4, s0);
    cudaMemsetAsync(F->vertices, 0, n * 12, sl);
    cudaMemsetAsync(F->vols[0].vert, 0, n * 12, s0);
    cudaMemsetAsync(F->normals, 0, n * 12, sl);
    cudaMemsetAsync(F->vols[0].norm, 0, n * 12, s0);
    cudaMemsetAsync(F->modelSegmentation, 0, n, sl);
    cudaMemsetAsync(F->vols[0].seg, 0, n, s0);
    for (int i = 1; i < nv; ++i) {
        RefVolume& v = F->vols[i];
        cudaMemsetAsync(v.ray, 0, n * 4, v.stream); cudaMemsetAsync(v.vert, 0, n * 12, v.stream);
        cudaMemsetAsync(v.norm, 0, n * 12, v.stream); cudaMemsetAsync(v.seg, 0, n, v.stream);
    }
    F->vis_objs.clear();
    for (int i = 0; i < nv; ++i) {                                              // :747-754
        RefVolume& v = F->vols[i];
        Stream st(v.stream);
        const float* weights = v.weights;
        if (v.fg_probs) {                                                       // ObjTSDF::raycast, src/core/ObjTSDF.cpp:209-210
            cudaMemsetAsync(v.raycastWeights, 0, v.nvox() * 4, v.stream);
            k_copy_f_masked<<<nb(v.nvox()), T, 0, v.stream>>>(v.weights, v.nvox(), v.raycastWeights, v.fgVolMask);
            weights = v.raycastWeights;
        }
        GpuMat tv = mat(v.tsdf, v.res[1] * v.res[2], v.res[0], CV_32FC1), gv = mat(v.grads, v.res[1] * v.res[2], v.res[0], CV_32FC3),
               wv = mat((void*)weights, v.res[1] * v.res[2], v.res[0], CV_32FC1);
        GpuMat r = mat(v.ray, F->h, F->w, CV_32FC1), ve = mat(v.vert, F->h, F->w, CV_32FC3), no = mat(v.norm, F->h, F->w, CV_32FC3),
               m = mat(v.seg, F->h, F->w, CV_8UC1);
        emf::cuda::TSDF::raycastTSDF(tv, gv, wv, r, ve, no, m, m33(R_co + 9 * i), v3(t_co + 3 * i), m33(K), v3i(v.res), v.voxel, v.trunc, st);
    }
    wait_all(F);                                                                // :756-758
    for (int i = 1; i < nv; ++i) {                                              // :760-771 (default stream)
        RefVolume& v = F->vols[i];
        k_cmp_le_s<<<nb(n), T>>>(F->raylengths, n, 0.f, F->mask);
        k_cmp_lt<<<nb(n), T>>>(v.ray, F->raylengths, n, F->obj_mask);
        k_or<<<nb(n), T>>>(F->obj_mask, F->mask, n, F->mask);
        k_and<<<nb(n), T>>>(v.seg, F->mask, n, F->mask);
        k_copy_f_masked<<<nb(n), T>>>(v.ray, n, F->raylengths, F->mask);
        k_copy_f3_masked<<<nb(n), T>>>(v.vert, n, F->vertices, F->mask);
        k_copy_f3_masked<<<nb(n), T>>>(v.norm, n, F->normals, F->mask);
        k_set_u8_masked<<<nb(n), T>>>(F->modelSegmentation, n, (uint8_t)(v.id > 255 ? 255 : v.id), F->mask);
    }
    // :773-776.  diffRaylengths is stale where bg_mask == 0 in the reference; zeroed here so those
    // pixels never take the background (the documented deterministic reading).
    cudaMemsetAsync(F->diffRaylengths, 0, n * 4, nullptr);
    k_sub_masked<<<nb(n), T>>>(F->raylengths, F->vols[0].ray, n, F->diffRaylengths, F->vols[0].seg);
    k_cmp_gt_s<<<nb(n), T>>>(F->diffRaylengths, n, 0.05f, F->takeBgMask);
    k_set_u8_masked<<<nb(n), T>>>(F->modelSegmentation, n, 0, F->takeBgMask);
    k_cmp_eq_u8s<<<nb(n), T>>>(F->modelSegmentation, n, 0, F->noObjMask);
    for (int i = 1; i < nv; ++i) {                                              // :778-791
        RefVolume& v = F->vols[i];
        k_cmp_eq_u8s<<<nb(n), T>>>(F->modelSegmentation, n, v.id, F->obj_mask);
        const int rw = F->w - 2 * boundary, rh = F->h - 2 * boundary;
        int cnt = 0;
        cudaMemsetAsync(F->d_count, 0, 4, nullptr);
        if (rw > 0 && rh > 0)
            k_count_rect<<<nb((size_t)rw * rh), T>>>(F->obj_mask, F->w, boundary, boundary, rw, rh, F->d_count);
        cudaMemcpy(&cnt, F->d_count, 4, cudaMemcpyDeviceToHost);                // countNonZero: blocking read
        F->vis_flags[i] = cnt > visibility_thresh;
        if (cnt > visibility_thresh) F->vis_objs.insert(v.id);
    }
    k_copy_f3_masked<<<nb(n), T>>>(F->vols[0].vert, n, F->vertices, F->noObjMask);   // :793-794
    k_copy_f3_masked<<<nb(n), T>>>(F->vols[0].norm, n, F->normals, F->noObjMask);
    cudaStreamSynchronize(nullptr);
    return status();
}

// EMFusion::integrateDepth, src/core/EMFusion.cpp:865-889.  use_vis != 0: only objects in vis_objs.
REF_API int emfref_frame_integrate(void* h, const float* depth, const float* R_oc, const float* t_oc, const float* K,
                                   float maxw, int use_vis) {
    RefFrame* F = (RefFrame*)h;
    const int nv = (int)F->vols.size();
    GpuMat d = mat((void*)depth, F->h, F->w, CV_32FC1);
    for (int i = 0; i < nv; ++i) {
        RefVolume& v = F->vols[i];
        if (i > 0 && use_vis && !F->vis_flags[i]) continue;
        Stream st(v.stream);
        GpuMat a = mat(v.assoc, F->h, F->w, CV_32FC1);
        GpuMat tv = mat(v.tsdf, v.res[1] * v.res[2], v.res[0], CV_32FC1), wv = mat(v.weights, v.res[1] * v.res[2], v.res[0], CV_32FC1);
        emf::cuda::TSDF::updateTSDF(d, a, tv, wv, m33(R_oc + 9 * i), v3(t_oc + 3 * i), m33(K), v3i(v.res), v.voxel, v.trunc, maxw, st);
    }
    for (int i = 0; i < nv; ++i) {
        RefVolume& v = F->vols[i];
        if (i > 0 && use_vis && !F->vis_flags[i]) continue;
        Stream st(v.stream);
        GpuMat tv = mat(v.tsdf, v.res[1] * v.res[2], v.res[0], CV_32FC1), g = mat(v.grads, v.res[1] * v.res[2], v.res[0], CV_32FC3);
        g.setTo(cv::Scalar::all(0.f), st);                                       // TSDF::updateGradients, src/core/TSDF.cpp:120-123
        emf::cuda::TSDF::computeTSDFGrads(tv, g, v3i(v.res), st);
    }
    wait_all(F);
    return status();
}

// accessors (device pointers owned by the frame)
REF_API float* emfref_frame_assoc_ptr(void* h, int i) { return ((RefFrame*)h)->vols[i].assoc; }
REF_API float* emfref_frame_vol_ray(void* h, int i) { return ((RefFrame*)h)->vols[i].ray; }
REF_API float* emfref_frame_vol_vert(void* h, int i) { return ((RefFrame*)h)->vols[i].vert; }
REF_API float* emfref_frame_vol_norm(void* h, int i) { return ((RefFrame*)h)->vols[i].norm; }
REF_API uint8_t* emfref_frame_vol_mask(void* h, int i) { return ((RefFrame*)h)->vols[i].seg; }
REF_API float* emfref_frame_ray(void* h) { return ((RefFrame*)h)->raylengths; }
REF_API float* emfref_frame_vert(void* h) { return ((RefFrame*)h)->vertices; }
REF_API float* emfref_frame_norm(void* h) { return ((RefFrame*)h)->normals; }
REF_API uint8_t* emfref_frame_seg(void* h) { return ((RefFrame*)h)->modelSegmentation; }
REF_API int emfref_frame_visible(void* h, int i) { return ((RefFrame*)h)->vis_flags[i]; }
REF_API void emfref_frame_set_visible(void* h, int i, int v) { ((RefFrame*)h)->vis_flags[i] = v; }
// set the association image of volume i to a constant (bg_associationWeights.setTo(1), EMFusion.cpp:55; createObj :920)
REF_API int emfref_frame_fill_assoc(void* h, int i, float v) {
    RefFrame* F = (RefFrame*)h;
    cvk::k_set_f<<<cvk::nb(F->n), cvk::T>>>(F->vols[i].assoc, F->n, v);
    cudaDeviceSynchronize();
    return status();
}
// device-to-device copy out of frame-owned buffers (used by the Python test binding)
REF_API int emfref_memcpy_d2d(void* dst, const void* src, size_t n) {
    cudaDeviceSynchronize();
    const cudaError_t e = cudaMemcpy(dst, src, n, cudaMemcpyDefault);   // any direction (UVA)
    cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : -2;
}

// emf::cuda::EMFusion::computePoints (src/core/cuda/EMFusion.cu:29-61).  That file does not compile
// against CCCL 2.8 (thrust::sort at :87), so its ten lines of arithmetic are restated here, including the
// launcher's setTo(0) and cudaDeviceSynchronize() (:57-60).
__global__ void k_ref_points(const float* depth, float* points, int w, int h, float fx, float fy, float cx, float cy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float d = depth[(size_t)y * w + x];
    float* p = points + 3 * ((size_t)y * w + x);
    p[0] = __fdiv_rn(__fmul_rn(__fsub_rn((float)x, cx), d), fx);
    p[1] = __fdiv_rn(__fmul_rn(__fsub_rn((float)y, cy), d), fy);
    p[2] = d;
}
REF_API int emfref_compute_points(const float* depth, float* points, int w, int h, const float* K) {
    cudaMemsetAsync(points, 0, (size_t)w * h * 12, nullptr);
    dim3 threads(32, 32), blocks((w + 31) / 32, (h + 31) / 32);
    k_ref_points<<<blocks, threads>>>(depth, points, w, h, K[0], K[4], K[2], K[5]);
    cudaDeviceSynchronize();
    return status();
}

#ifndef EMFREF_NO_TRACKER   // (the 'reference launch structure + this repo's operators' build has no tracker operators)
// ------------------------------------------------------------------------------------------------
// 4. tracker: the device part of one iteration of emf::TSDF's Levenberg-Marquardt loop, as
//    emf::EMFusion::performTracking drives it (src/core/EMFusion.cpp:672-722): the reference kernels
//    (computePoseGradients, getVolumeVals x 2, computeAb, multSingletonCol x 2) + the OpenCV-CUDA ops between them,
//    one launch per op, on the four streams and events of src/core/TSDF.cpp:194-265,375-394.
// ------------------------------------------------------------------------------------------------
namespace cvk {
#define IDX size_t i = (size_t)blockIdx.x * T + threadIdx.x; if (i >= n) return;
__global__ void k_abs_to(const float* a, size_t n, float* d) { IDX d[i] = fabsf(a[i]); }
// cv::cuda::divide(scalar, mat): scalar / a, 0 where a == 0
__global__ void k_sdiv(float s, const float* a, size_t n, float* d) { IDX d[i] = a[i] != 0.f ? __fdiv_rn(s, a[i]) : 0.f; }
__global__ void k_min_s(const float* a, size_t n, float v, float* d) { IDX d[i] = fminf(a[i], v); }
__global__ void k_sqr(const float* a, size_t n, float* d) { IDX d[i] = __fmul_rn(a[i], a[i]); }
__global__ void k_absmax(const float* a, size_t n, unsigned* out) {   // NORM_INF of a float image (bit pattern of |x| is monotone)
    IDX atomicMax(out, __float_as_uint(fabsf(a[i])));
}
#undef IDX
// cv::cuda::reduce(src, dst, 0, REDUCE_SUM): column sums of a rows x cols float matrix, float accumulation
__global__ void k_colsum(const float* a, size_t rows, int cols, float* out) {
    __shared__ float s[256];
    const int c = blockIdx.x;
    float acc = 0.f;
    for (size_t r = threadIdx.x; r < rows; r += 256) acc = __fadd_rn(acc, a[r * cols + c]);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) s[threadIdx.x] = __fadd_rn(s[threadIdx.x], s[threadIdx.x + o]); __syncthreads(); }
    if (threadIdx.x == 0) out[c] = s[0];
}
// cv::cuda::sum: double accumulation
__global__ void k_sum_d(const float* a, size_t n, double* out) {
    __shared__ double s[256];
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) acc += (double)a[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) atomicAdd(out, s[0]);
}
}  // namespace cvk

struct RefTracker {   // the tracking members of emf::TSDF (include/EMFusion/core/TSDF.h:302-326)
    int w, h; size_t n;
    float *grads, *tsdfVals, *intWeights, *trackWeights, *As, *bs, *A_gpu, *b_gpu, *errors;
    unsigned* d_max; double* d_sum;
    cudaStream_t streams[4];
    cudaEvent_t events[5];
};

REF_API void* emfref_tracker_create(int w, int h) {
    RefTracker* K = new RefTracker();
    K->w = w; K->h = h; K->n = (size_t)w * h;
    K->grads = dalloc<float>(6 * K->n); K->tsdfVals = dalloc<float>(K->n); K->intWeights = dalloc<float>(K->n);
    K->trackWeights = dalloc<float>(K->n); K->As = dalloc<float>(36 * K->n); K->bs = dalloc<float>(6 * K->n);
    K->A_gpu = dalloc<float>(36); K->b_gpu = dalloc<float>(6); K->errors = dalloc<float>(K->n);
    K->d_max = dalloc<unsigned>(1); K->d_sum = dalloc<double>(1);
    for (auto& s : K->streams) cudaStreamCreate(&s);
    for (auto& e : K->events) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaDeviceSynchronize();
    return K;
}
REF_API void emfref_tracker_destroy(void* h) {
    RefTracker* K = (RefTracker*)h;
    cudaFree(K->grads); cudaFree(K->tsdfVals); cudaFree(K->intWeights); cudaFree(K->trackWeights); cudaFree(K->As);
    cudaFree(K->bs); cudaFree(K->A_gpu); cudaFree(K->b_gpu); cudaFree(K->errors); cudaFree(K->d_max); cudaFree(K->d_sum);
    for (auto& s : K->streams) cudaStreamDestroy(s);
    for (auto& e : K->events) cudaEventDestroy(e);
    delete K;
}
REF_API float* emfref_tracker_ptr(void* h, int what) {
    RefTracker* K = (RefTracker*)h;
    float* p[] = {K->grads, K->tsdfVals, K->intWeights, K->trackWeights, K->As, K->bs, K->A_gpu, K->b_gpu};
    return what >= 0 && what < 8 ? p[what] : nullptr;
}

// TSDF::computeError (src/core/TSDF.cpp:390-394): blocking
static float ref_compute_error(RefTracker* K) {
    using namespace cvk;
    k_sqr<<<nb(K->n), T>>>(K->tsdfVals, K->n, K->errors);
    k_mul<<<nb(K->n), T>>>(K->errors, K->intWeights, K->n, K->errors);
    cudaMemsetAsync(K->d_sum, 0, 8, nullptr);
    k_sum_d<<<296, 256>>>(K->errors, K->n, K->d_sum);
    double s = 0;
    cudaMemcpy(&s, K->d_sum, 8, cudaMemcpyDeviceToHost);
    return (float)s;
}

// One gradient-evaluating iteration up to and including reduceHessians' downloads (A, b on the host) and the
// err = computeError() of computePoseUpdate.  A_out[36], b_out[6], err_out on the host.
REF_API int emfref_tracker_linearise(void* h, const float* tsdf, const float* grads_vol, const float* weights,
                                     const float* points, const float* assoc, const float* R, const float* t,
                                     const int* res, float voxel, float huber, float maxw, float* A_out, float* b_out,
                                     float* err_out) {
    using namespace cvk;
    RefTracker* K = (RefTracker*)h;
    const size_t n = K->n;
    cudaStream_t* s = K->streams;
    Stream st0(s[0]), st1(s[1]), st2(s[2]);
    GpuMat tv = mat((void*)tsdf, res[1] * res[2], res[0], CV_32FC1), gv = mat((void*)grads_vol, res[1] * res[2], res[0], CV_32FC3),
           wv = mat((void*)weights, res[1] * res[2], res[0], CV_32FC1);
    GpuMat p = mat((void*)points, K->h, K->w, CV_32FC3);
    GpuMat g = mat(K->grads, (int)n, 6, CV_32FC1), tvals = mat(K->tsdfVals, K->h, K->w, CV_32FC1),
           iw = mat(K->intWeights, K->h, K->w, CV_32FC1);
    GpuMat As = mat(K->As, (int)n, 36, CV_32FC1), bs = mat(K->bs, (int)n, 6, CV_32FC1);
    // computeGradients / computeTSDFVals / computeTSDFWeights (src/core/TSDF.cpp:194-221)
    emf::cuda::TSDF::computePoseGradients(gv, p, m33(R), v3(t), v3i(res), voxel, g, st0);
    emf::cuda::TSDF::getVolumeVals(tv, p, m33(R), v3(t), v3i(res), voxel, tvals, st1);
    cudaEventRecord(K->events[1], s[1]);
    emf::cuda::TSDF::getVolumeVals(wv, p, m33(R), v3(t), v3i(res), voxel, iw, st2);
    // computeHuberWeights (:223-232)
    cudaStreamWaitEvent(s[3], K->events[1], 0);
    k_abs_to<<<nb(n), T, 0, s[3]>>>(K->tsdfVals, n, K->trackWeights);
    k_sdiv<<<nb(n), T, 0, s[3]>>>(huber, K->trackWeights, n, K->trackWeights);
    k_min_s<<<nb(n), T, 0, s[3]>>>(K->trackWeights, n, 1.0f, K->trackWeights);
    // normalizeTSDFWeights (:234-242): min, normalize(alpha = 1, NORM_INF) = norm (blocking read) + convertTo(scale)
    k_min_s<<<nb(n), T, 0, s[2]>>>(K->intWeights, n, maxw, K->intWeights);
    cudaMemsetAsync(K->d_max, 0, 4, s[2]);
    k_absmax<<<nb(n), T, 0, s[2]>>>(K->intWeights, n, K->d_max);
    unsigned mx = 0;
    cudaMemcpyAsync(&mx, K->d_max, 4, cudaMemcpyDeviceToHost, s[2]);
    cudaStreamSynchronize(s[2]);
    float fmx; memcpy(&fmx, &mx, 4);
    const double nrm = (double)fmx;
    const float scale = nrm > 2.220446049250313e-16 ? (float)(1.0 / nrm) : 0.f;
    k_mul_s<<<nb(n), T, 0, s[2]>>>(K->intWeights, n, scale, K->intWeights);
    cudaEventRecord(K->events[2], s[2]);
    // combineWeights (:244-255)
    cudaStreamWaitEvent(s[3], K->events[2], 0);
    k_mul<<<nb(n), T, 0, s[3]>>>(K->trackWeights, K->intWeights, n, K->intWeights);
    k_mul<<<nb(n), T, 0, s[3]>>>(K->intWeights, assoc, n, K->intWeights);
    cudaEventRecord(K->events[3], s[3]);
    // computeHessians (:257-265); events[4] is never recorded in the reference, so the wait is a no-op
    emf::cuda::TSDF::computeAb(g, tvals, As, bs, st0);
    cudaEventRecord(K->events[0], s[0]);
    // reduceAb (:375-388)
    cudaStreamWaitEvent(s[0], K->events[3], 0);
    cudaStreamWaitEvent(s[1], K->events[3], 0);
    cudaStreamWaitEvent(s[1], K->events[0], 0);
    GpuMat iw_col = mat(K->intWeights, (int)n, 1, CV_32FC1);
    emf::cuda::TSDF::multSingletonCol(iw_col, As, As, st0);
    emf::cuda::TSDF::multSingletonCol(iw_col, bs, bs, st1);
    k_colsum<<<36, 256, 0, s[0]>>>(K->As, n, 36, K->A_gpu);
    k_colsum<<<6, 256, 0, s[1]>>>(K->bs, n, 6, K->b_gpu);
    // reduceHessians (:267-283): two downloads + waitForCompletion
    cudaMemcpyAsync(A_out, K->A_gpu, 36 * 4, cudaMemcpyDeviceToHost, s[0]);
    cudaMemcpyAsync(b_out, K->b_gpu, 6 * 4, cudaMemcpyDeviceToHost, s[1]);
    cudaStreamSynchronize(s[0]);
    cudaStreamSynchronize(s[1]);
    if (err_out) *err_out = ref_compute_error(K);
    return status();
}

// the trial pose of computePoseUpdate (src/core/TSDF.cpp:312-315): computeTSDFVals + waitForCompletion + computeError
REF_API int emfref_tracker_error(void* h, const float* tsdf, const float* points, const float* R, const float* t,
                                 const int* res, float voxel, float* err_out) {
    RefTracker* K = (RefTracker*)h;
    Stream st1(K->streams[1]);
    GpuMat tv = mat((void*)tsdf, res[1] * res[2], res[0], CV_32FC1);
    GpuMat p = mat((void*)points, K->h, K->w, CV_32FC3), tvals = mat(K->tsdfVals, K->h, K->w, CV_32FC1);
    emf::cuda::TSDF::getVolumeVals(tv, p, m33(R), v3(t), v3i(res), voxel, tvals, st1);
    cudaStreamSynchronize(K->streams[1]);
    *err_out = ref_compute_error(K);
    return status();
}

#endif   // EMFREF_NO_TRACKER
// emf::cuda::TSDF::copyValues (src/core/cuda/TSDF.cu:768-819); channels 1 / 2 / 3
REF_API int emfref_copy_values(const float* src, float* dst, int channels, const int* offset, const int* src_res,
                               const int* dst_res) {
    const int type = channels == 1 ? CV_32FC1 : (channels == 2 ? CV_32FC2 : CV_32FC3);
    GpuMat s = mat((void*)src, src_res[1] * src_res[2], src_res[0], type);
    GpuMat d = mat(dst, dst_res[1] * dst_res[2], dst_res[0], type);
    emf::cuda::TSDF::copyValues(s, d, v3i(offset), v3i(src_res), v3i(dst_res));
    cudaDeviceSynchronize();
    return status();
}


#ifndef EMFREF_NO_TRACKER   // (reference-only operators: not in the build over this repo's operators)
// emf::cuda::TSDF::marchingCubes (src/core/cuda/TSDF.cu:1107-1152) with the mask of emf::TSDF::getMesh (src/core/TSDF.cpp:357:
// weights > 0; objects: & fgVolMask, src/core/ObjTSDF.cpp:251-252) built by the caller.  Two calls like the product's: the first
// runs the reference launcher and keeps its outputs, the second copies them out.
struct RefMesh { GpuMat vertices, normals, triangles; };
REF_API void* emfref_marching_cubes(const float* tsdf, const float* grads, const uint8_t* mask, const int* res, float voxel, int* counts) {
    const int rows = res[1] * res[2];
    GpuMat tv = mat((void*)tsdf, rows, res[0], CV_32FC1), gv = mat((void*)grads, rows, res[0], CV_32FC3), mv = mat((void*)mask, rows, res[0], CV_8UC1);
    GpuMat cls = cv::cuda::createContinuous((res[1] - 1) * (res[2] - 1), res[0] - 1, CV_8UC1);
    GpuMat vib = cv::cuda::createContinuous((res[1] - 1) * (res[2] - 1), res[0] - 1, CV_32SC1);
    GpuMat tib = cv::cuda::createContinuous((res[1] - 1) * (res[2] - 1), res[0] - 1, CV_32SC1);
    cls.setTo(cv::Scalar(0)); vib.setTo(cv::Scalar(0)); tib.setTo(cv::Scalar(0));
    cudaDeviceSynchronize();
    RefMesh* m = new RefMesh();
    emf::cuda::TSDF::marchingCubes(tv, gv, mv, v3i(res), voxel, cls, vib, tib, m->vertices, m->normals, m->triangles);
    cudaDeviceSynchronize();
    counts[0] = m->vertices.cols; counts[1] = m->triangles.cols;
    return m;
}
REF_API int emfref_mesh_fetch(void* h, float* vertices, float* normals, int* triangles) {
    RefMesh* m = (RefMesh*)h;
    if (m->vertices.cols > 0) {
        cudaMemcpy(vertices, m->vertices.data, (size_t)m->vertices.cols * 12, cudaMemcpyDeviceToDevice);
        cudaMemcpy(normals, m->normals.data, (size_t)m->normals.cols * 12, cudaMemcpyDeviceToDevice);
        cudaMemcpy(triangles, m->triangles.data, (size_t)m->triangles.cols * 4, cudaMemcpyDeviceToDevice);
    }
    delete m;
    return status();
}
#endif
