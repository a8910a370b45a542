// preprocess.cu -- depth pre-filter of a frame, fused with its NaN / zero patch and (optionally) the un-projection.
//
// Replaces emf::EMFusion::preprocessDepth (reference src/core/EMFusion.cpp:294-305): cv::cuda::bilateralFilter + compare +
// setTo + compare + setTo (five launches), and -- when a points image is given -- emf::cuda::EMFusion::computePoints
// (src/core/cuda/EMFusion.cu:29-61) right after it: one launch, the depth tile and its halo staged once in shared memory.
//
// PARITY UNPINNED.  The filter's arithmetic lives in OpenCV-CUDA (opencv_contrib cudaimgproc, bilateral_filter.cu), an
// un-vendored, version-unpinned dependency of the reference that is absent from this image ("tested with 4.3.0",
// README.md:42-43).  Its published kernel is restated: window = the disc of radius ksize / 2 inside the ksize x ksize
// square, weight = exp(space2 * (-0.5 / sigma_spatial^2) + (v - centre)^2 * (-0.5 / sigma_depth^2)), border
// BORDER_REFLECT_101, result = sum(w v) / sum(w); rows outermost, columns innermost, float accumulation.  Checked against
// an independent scalar C restatement under the test tree -- not against OpenCV itself.
#include "common.cuh"

namespace emfb {

constexpr int kPreW = 32, kPreH = 8, kPreMaxR = 7;

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

__global__ void __launch_bounds__(kPreW * kPreH) k_preprocess(Img<const float> raw, Img<float> out, Img<float> points, int r,
                                                              float ss, float sc, float fx, float fy, float cx, float cy) {
    __shared__ float tile[kPreH + 2 * kPreMaxR][kPreW + 2 * kPreMaxR + 1];
    const int x0 = blockIdx.x * kPreW, y0 = blockIdx.y * kPreH;
    const int tw = kPreW + 2 * r, th = kPreH + 2 * r;
    for (int i = threadIdx.x; i < tw * th; i += kPreW * kPreH) {
        const int ty = i / tw, tx = i - ty * tw;
        tile[ty][tx] = __ldg(raw.row(reflect101(y0 + ty - r, raw.h)) + reflect101(x0 + tx - r, raw.w));
    }
    __syncthreads();
    const int lx = threadIdx.x & (kPreW - 1), ly = threadIdx.x / kPreW;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= raw.w || y >= raw.h) return;
    const float centre = tile[ly + r][lx + r];
    const float r2 = (float)(r * r);
    float sum1 = 0.0f, sum2 = 0.0f;
    for (int dy = -r; dy <= r; ++dy)
        for (int dx = -r; dx <= r; ++dx) {
            const float space2 = (float)(dx * dx + dy * dy);
            if (space2 > r2) continue;
            const float v = tile[ly + r + dy][lx + r + dx];
            const float d = fsub(v, centre);
            const float w = expf(ffma(space2, ss, fmul(fmul(d, d), sc)));
            sum1 = ffma(w, v, sum1);
            sum2 = fadd(sum2, w);
        }
    float res = fdiv(sum1, sum2);
    if (res != res) res = 0.0f;              // compare(depth, depth, CMP_NE) + setTo(0)   (EMFusion.cpp:301-302)
    if (centre == 0.0f) res = 0.0f;          // compare(depth_raw, 0, CMP_EQ) + setTo(0)   (:303-304)
    out.at(y, x) = res;
    if (points.ptr) {
        float* p = points.row(y) + 3 * x;
        p[0] = fdiv(fmul(fsub((float)x, cx), res), fx);
        p[1] = fdiv(fmul(fsub((float)y, cy), res), fy);
        p[2] = res;
    }
}

}  // namespace emfb

using namespace emfb;

extern "C" EMF_API int emf_preprocess_depth(const emf_image* depth_raw, const emf_image* depth_out, const emf_image* points_out,
                                            const float K[9], int kernel_size, float sigma_depth, float sigma_spatial,
                                            emf_stream_t stream) {
    if (!image_ok(depth_raw, 4) || !image_ok(depth_out, 4) || !same_size(depth_raw, depth_out) || depth_raw->ptr == depth_out->ptr)
        return EMF_ERR_INVALID;
    if (points_out && (!K || !image_ok(points_out, 12) || !same_size(depth_raw, points_out))) return EMF_ERR_INVALID;
    // cv::cuda::bilateralFilter's parameter handling
    if (sigma_depth <= 0.0f) sigma_depth = 1.0f;
    if (sigma_spatial <= 0.0f) sigma_spatial = 1.0f;
    int radius = kernel_size <= 0 ? (int)lrintf(sigma_spatial * 1.5f) : kernel_size / 2;
    if (radius < 1) radius = 1;
    if (radius > kPreMaxR) return EMF_ERR_UNSUPPORTED;
    const float ss = -0.5f / (sigma_spatial * sigma_spatial), sc = -0.5f / (sigma_depth * sigma_depth);
    const dim3 grid((depth_raw->width + kPreW - 1) / kPreW, (depth_raw->height + kPreH - 1) / kPreH);
    k_preprocess<<<grid, kPreW * kPreH, 0, (cudaStream_t)stream>>>(
        view<const float>(depth_raw), view<float>(depth_out), points_out ? view<float>(points_out) : null_view<float>(), radius, ss,
        sc, K ? K[0] : 1.0f, K ? K[4] : 1.0f, K ? K[2] : 0.0f, K ? K[5] : 0.0f);
    return launch_status();
}
