// resize.cu -- moving a volume's contents into a grid of another resolution.
//
// Replaces emf::cuda::TSDF::copyValues (reference src/core/cuda/TSDF.cu:768-819; include/EMFusion/core/cuda/TSDF.cuh:228)
// and the device part of emf::ObjTSDF::resize (src/core/ObjTSDF.cpp:116-147): four `setTo(0)` over the new arrays plus
// four copyValues launches (tsdf, weights, gradients, fg/bg counts), each a full pass with one 4/8/12-byte access per
// thread.  emf_resize_volume writes every array of the new grid exactly once in ONE launch -- the copied value where the
// old grid covers the voxel, zero elsewhere -- with 16-byte stores along x; the gradient volume is not copied at all
// (it is a pure function of tsdf: emf_compute_tsdf_grads rebuilds it for whoever wants it, bit-identically away from the
// old grid's border planes -- DESIGN.md section 4.4).
#include "common.cuh"

namespace emfb {

// dst(x - ox, y - oy, z - oz) = src(x, y, z) wherever the target is inside dst; CH floats per voxel.
// One thread per DESTINATION voxel; FILL: destinations without a source are zeroed (the callee's setTo(0)).
template <int CH, bool FILL>
__global__ void __launch_bounds__(256) k_copy_values(const float* __restrict__ src, float* __restrict__ dst, int ox, int oy, int oz,
                                                     int sx, int sy, int sz, int dx, int dy, int dz) {
    const int64_t n = (int64_t)dx * dy * dz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / dx;
        const int x = (int)(i - row * dx);
        const int z = (int)(row / dy);
        const int y = (int)(row - (int64_t)z * dy);
        const int xs = x + ox, ys = y + oy, zs = z + oz;
        const bool in = xs >= 0 && xs < sx && ys >= 0 && ys < sy && zs >= 0 && zs < sz;
        if (in) {
            const float* s = src + CH * (((int64_t)zs * sy + ys) * sx + xs);
#pragma unroll
            for (int c = 0; c < CH; ++c) dst[CH * i + c] = __ldg(s + c);
        } else if (FILL) {
#pragma unroll
            for (int c = 0; c < CH; ++c) dst[CH * i + c] = 0.0f;
        }
    }
}

// ObjTSDF::resize: tsdf, weights (1 float) and fg/bg counts (2 floats) of the new grid in one pass; a thread owns four
// consecutive destination voxels of a row (dx % 4 == 0, 16-byte aligned arrays) and writes them with 128-bit stores.
struct ResizeParams {
    const float* s_tsdf; const float* s_w; const float* s_fgbg;   // s_fgbg nullable
    float* d_tsdf; float* d_w; float* d_fgbg;
    int ox, oy, oz, sx, sy, sz, dx, dy, dz;
};
__global__ void __launch_bounds__(256) k_resize(const __grid_constant__ ResizeParams P) {
    const int qx = P.dx >> 2;
    const int64_t n = (int64_t)qx * P.dy * P.dz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / qx;
        const int x0 = (int)(i - row * qx) << 2;
        const int z = (int)(row / P.dy);
        const int y = (int)(row - (int64_t)z * P.dy);
        const int ys = y + P.oy, zs = z + P.oz;
        const bool row_in = ys >= 0 && ys < P.sy && zs >= 0 && zs < P.sz;
        float t[4] = {0.f, 0.f, 0.f, 0.f}, w[4] = {0.f, 0.f, 0.f, 0.f}, f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (row_in) {
            const int64_t sb = ((int64_t)zs * P.sy + ys) * P.sx;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xs = x0 + j + P.ox;
                if (xs >= 0 && xs < P.sx) {
                    t[j] = __ldg(P.s_tsdf + sb + xs);
                    w[j] = __ldg(P.s_w + sb + xs);
                    if (P.s_fgbg) { f[2 * j] = __ldg(P.s_fgbg + 2 * (sb + xs)); f[2 * j + 1] = __ldg(P.s_fgbg + 2 * (sb + xs) + 1); }
                }
            }
        }
        const int64_t db = row * P.dx + x0;
        *reinterpret_cast<float4*>(P.d_tsdf + db) = make_float4(t[0], t[1], t[2], t[3]);
        *reinterpret_cast<float4*>(P.d_w + db) = make_float4(w[0], w[1], w[2], w[3]);
        if (P.d_fgbg) {
            *reinterpret_cast<float4*>(P.d_fgbg + 2 * db) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(P.d_fgbg + 2 * db + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
    }
}

static unsigned grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace emfb

using namespace emfb;

extern "C" EMF_API int emf_copy_values(const float* src, float* dst, int channels, const int offset[3], const int src_res[3],
                                       const int dst_res[3], emf_stream_t stream) {
    if (!src || !dst || !offset || !res_ok(src_res) || !res_ok(dst_res) || src == dst) return EMF_ERR_INVALID;
    if (channels < 1 || channels > 3) return EMF_ERR_UNSUPPORTED;
    const int64_t n = (int64_t)dst_res[0] * dst_res[1] * dst_res[2];
    const cudaStream_t s = (cudaStream_t)stream;
#define EMF_CV(CH) k_copy_values<CH, false><<<grid_for(n), 256, 0, s>>>(src, dst, offset[0], offset[1], offset[2], src_res[0], \
                                                                      src_res[1], src_res[2], dst_res[0], dst_res[1], dst_res[2])
    if (channels == 1) EMF_CV(1); else if (channels == 2) EMF_CV(2); else EMF_CV(3);
#undef EMF_CV
    return launch_status();
}

extern "C" EMF_API int emf_resize_volume(const float* src_tsdf, const float* src_weights, const float* src_fgbg,
                                         const int src_res[3], float* dst_tsdf, float* dst_weights, float* dst_fgbg,
                                         const int dst_res[3], const int offset[3], emf_stream_t stream) {
    if (!src_tsdf || !src_weights || !dst_tsdf || !dst_weights || !offset || !res_ok(src_res) || !res_ok(dst_res))
        return EMF_ERR_INVALID;
    if ((src_fgbg == nullptr) != (dst_fgbg == nullptr)) return EMF_ERR_INVALID;
    const cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = (int64_t)dst_res[0] * dst_res[1] * dst_res[2];
    if (dst_res[0] % 4 == 0 && aligned16(dst_tsdf) && aligned16(dst_weights) && (!dst_fgbg || aligned16(dst_fgbg))) {
        ResizeParams P;
        P.s_tsdf = src_tsdf; P.s_w = src_weights; P.s_fgbg = src_fgbg;
        P.d_tsdf = dst_tsdf; P.d_w = dst_weights; P.d_fgbg = dst_fgbg;
        P.ox = offset[0]; P.oy = offset[1]; P.oz = offset[2];
        P.sx = src_res[0]; P.sy = src_res[1]; P.sz = src_res[2];
        P.dx = dst_res[0]; P.dy = dst_res[1]; P.dz = dst_res[2];
        k_resize<<<grid_for(n / 4), 256, 0, s>>>(P);
        return launch_status();
    }
    // any resolution / alignment: one filling pass per array
#define EMF_RS(CH, S, D) k_copy_values<CH, true><<<grid_for(n), 256, 0, s>>>(S, D, offset[0], offset[1], offset[2], src_res[0], \
                                                                           src_res[1], src_res[2], dst_res[0], dst_res[1], dst_res[2])
    EMF_RS(1, src_tsdf, dst_tsdf);
    EMF_RS(1, src_weights, dst_weights);
    if (dst_fgbg) EMF_RS(2, src_fgbg, dst_fgbg);
#undef EMF_RS
    return launch_status();
}
