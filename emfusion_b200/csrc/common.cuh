// common.cuh -- shared device/host plumbing for the C-ABI kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/emf_b200.h"
#include "emf_math.cuh"

namespace emfb {

// status codes of an 8-voxel x-segment (emf_volume::seg_status) / an 8^3 brick (emf_volume::brick_flags)
enum : int { kMixed = 0, kAllOne = 1, kAllZero = 2, kAllMinusOne = 3 };

// Pitched 2-D image view (device side).
template <typename T>
struct Img {
    T* ptr;
    size_t pitch;  // bytes
    int w, h;
    __device__ __forceinline__ T* row(int y) const { return (T*)((char*)ptr + (size_t)y * pitch); }
    __device__ __forceinline__ T& at(int y, int x) const { return row(y)[x]; }
};

template <typename T>
static inline Img<T> view(const emf_image* im) {
    Img<T> v;
    v.ptr = (T*)im->ptr; v.pitch = im->pitch; v.w = im->width; v.h = im->height;
    return v;
}
template <typename T>
static inline Img<T> null_view() { Img<T> v; v.ptr = nullptr; v.pitch = 0; v.w = 0; v.h = 0; return v; }

static inline bool image_ok(const emf_image* im, size_t elem) {
    return im && im->ptr && im->width > 0 && im->height > 0 && im->pitch >= (size_t)im->width * elem &&
           (im->pitch % 4 == 0 || elem == 1);
}
static inline bool same_size(const emf_image* a, const emf_image* b) {
    return a->width == b->width && a->height == b->height;
}
static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
static inline bool res_ok(const int* res) {
    if (!res) return false;
    if (res[0] < 2 || res[1] < 2 || res[2] < 2) return false;
    // row index (z*Ry + y) and in-volume int math stay below 2^31
    return (int64_t)res[1] * res[2] < (int64_t)1 << 31 && (int64_t)res[0] * res[1] < (int64_t)1 << 31;
}

static inline int launch_status() {
    return cudaPeekAtLastError() == cudaSuccess ? EMF_OK : EMF_ERR_CUDA;
}

struct Pose { float R[9]; float t[3]; };
static inline Pose to_pose(const emf_pose* p) {
    Pose q;
    for (int i = 0; i < 9; ++i) q.R[i] = p->R[i];
    for (int i = 0; i < 3; ++i) q.t[i] = p->t[i];
    return q;
}
struct Intr { float K[9]; };
static inline Intr to_intr(const float* K) { Intr k; for (int i = 0; i < 9; ++i) k.K[i] = K[i]; return k; }
// Standard pinhole matrix [[fx,0,cx],[0,fy,cy],[0,0,1]]: lets K*p drop its exact-zero terms
// without changing a bit of the result.
static inline bool is_pinhole(const float* K) {
    return K[1] == 0.f && K[3] == 0.f && K[6] == 0.f && K[7] == 0.f && K[8] == 1.f;
}

}  // namespace emfb
