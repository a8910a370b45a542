// emf_b200_opencv_binding.cpp -- the reference-side binding of libemf_b200.so.
//
// This file is NOT part of the product library.  It is what a maintainer of EM-Fusion adds to the
// reference build: it DEFINES the reference's own level-1 operators
//   emf::cuda::TSDF::{updateTSDF, computeTSDFGrads, raycastTSDF, getVolumeVals, copyValues}
//       declared in  include/EMFusion/core/cuda/TSDF.cuh:115-230      (defined in src/core/cuda/TSDF.cu)
//   emf::cuda::ObjTSDF::updateFgBgProbs
//       declared in  include/EMFusion/core/cuda/ObjTSDF.cuh:49-56     (defined in src/core/cuda/ObjTSDF.cu)
//   emf::cuda::EMFusion::computePoints
//       declared in  include/EMFusion/core/cuda/EMFusion.cuh:39-40    (defined in src/core/cuda/EMFusion.cu)
// with exactly the reference's signatures, on top of the C ABI of include/emf_b200.h, so that
// src/core/TSDF.cpp, src/core/ObjTSDF.cpp and src/core/EMFusion.cpp compile and link UNCHANGED against the
// B200 kernels (the six definitions are removed from -- or #if'd out of -- the three .cu files).
//
// It only touches the public surface of cv::cuda::GpuMat (data, step, rows, cols, type()), cv::Matx33f /
// cv::Vec3f / cv::Vec3i (::val) and cv::cuda::StreamAccessor::getStream, i.e. it compiles against real
// OpenCV (>= 4.x with the cuda modules) and, for this repo's CPU-only compile test, against the type
// stand-in in oracle/shim/ (tests/test_binding_compiles.py).
//
// Error behaviour: the reference operators return void and never check CUDA errors; argument errors are
// OpenCV assertions.  A non-zero C-ABI status is therefore turned into CV_Assert-style failure here.
#include "EMFusion/core/cuda/TSDF.cuh"
#include "EMFusion/core/cuda/ObjTSDF.cuh"
#if __has_include("EMFusion/core/cuda/EMFusion.cuh") && !defined(EMF_B200_BINDING_NO_FRAME_OPS)
#include "EMFusion/core/cuda/EMFusion.cuh"
#define EMF_B200_BINDING_HAS_FRAME_OPS 1
#endif
#include <opencv2/core/cuda_stream_accessor.hpp>

#include "emf_b200.h"

namespace {

inline emf_image img(const cv::cuda::GpuMat& m) {
    emf_image i;
    i.ptr = (void*)m.data;
    i.pitch = m.step;
    i.width = m.cols;
    i.height = m.rows;
    return i;
}
inline emf_pose pose(const cv::Matx33f& R, const cv::Vec3f& t) {
    emf_pose p;
    for (int k = 0; k < 9; ++k) p.R[k] = R.val[k];   // row-major, as the reference reinterprets it
    for (int k = 0; k < 3; ++k) p.t[k] = t.val[k];   // (src/core/cuda/TSDF.cu:415-420)
    return p;
}
inline emf_stream_t str(cv::cuda::Stream& s) { return (emf_stream_t)cv::cuda::StreamAccessor::getStream(s); }
inline void ok(int rc) { CV_Assert(rc == EMF_OK); }

}  // namespace

namespace emf {
namespace cuda {
namespace TSDF {

void updateTSDF(const cv::cuda::GpuMat& depth, const cv::cuda::GpuMat& assocWeights, cv::cuda::GpuMat& tsdfVol,
                cv::cuda::GpuMat& tsdfWeights, const cv::Matx33f& rel_rot_OC, const cv::Vec3f& rel_trans_OC,
                const cv::Matx33f& intr, const cv::Vec3i& volumeRes, const float voxelSize, const float truncdist,
                const float maxWeight, cv::cuda::Stream& stream) {
    const emf_image d = img(depth), a = img(assocWeights);
    const emf_pose T = pose(rel_rot_OC, rel_trans_OC);
    ok(emf_update_tsdf(&d, &a, (float*)tsdfVol.data, (float*)tsdfWeights.data, &T, intr.val, volumeRes.val, voxelSize,
                       truncdist, maxWeight, str(stream)));
}

// TSDF::updateGradients calls tsdfGrads.setTo(0) first (src/core/TSDF.cpp:121); emf_compute_tsdf_grads writes
// every element including the zero planes, so that fill becomes redundant but harmless.
void computeTSDFGrads(const cv::cuda::GpuMat& tsdfVol, cv::cuda::GpuMat& tsdfGrads, const cv::Vec3i& volumeRes,
                      cv::cuda::Stream& stream) {
    ok(emf_compute_tsdf_grads((const float*)tsdfVol.data, (float*)tsdfGrads.data, volumeRes.val, str(stream)));
}

void raycastTSDF(const cv::cuda::GpuMat& tsdfVol, const cv::cuda::GpuMat& tsdfGrads, const cv::cuda::GpuMat& tsdfWeights,
                 cv::cuda::GpuMat& raylengths, cv::cuda::GpuMat& vertices, cv::cuda::GpuMat& normals,
                 cv::cuda::GpuMat& mask, const cv::Matx33f& rel_rot_CO, const cv::Vec3f& rel_trans_CO,
                 const cv::Matx33f& intr, const cv::Vec3i& volumeRes, const float voxelSize, const float truncdist,
                 cv::cuda::Stream& stream) {
    const emf_image r = img(raylengths), v = img(vertices), n = img(normals), m = img(mask);
    const emf_pose T = pose(rel_rot_CO, rel_trans_CO);
    ok(emf_raycast_tsdf((const float*)tsdfVol.data, (const float*)tsdfGrads.data, (const float*)tsdfWeights.data,
                        /*fg_probs=*/nullptr, &r, &v, &n, &m, &T, intr.val, volumeRes.val, voxelSize, truncdist,
                        /*hit_voxel=*/nullptr, str(stream)));
}

// Only the 1-channel float instantiation is on the hot path (association: tsdfVol and fgProbs,
// src/core/TSDF.cpp:144, src/core/ObjTSDF.cpp:189); 2/3-channel volumes (tracker) stay with the reference.
void getVolumeVals(const cv::cuda::GpuMat& vol, const cv::cuda::GpuMat& points, const cv::Matx33f& rel_rot_CO,
                   const cv::Vec3f& rel_trans_CO, const cv::Vec3i& volumeRes, const float voxelSize,
                   cv::cuda::GpuMat& vals, cv::cuda::Stream& stream) {
    CV_Assert(vol.type() == CV_32FC1);
    const emf_image p = img(points), o = img(vals);
    const emf_pose T = pose(rel_rot_CO, rel_trans_CO);
    ok(emf_get_volume_vals((const float*)vol.data, &p, &T, volumeRes.val, voxelSize, &o, str(stream)));
}

// ObjTSDF::resize (src/core/ObjTSDF.cpp:139-146) calls this once per array after zero-filling the target; a maintainer who
// edits resize() itself replaces the four fills and four calls by one emf_resize_volume.
void copyValues(const cv::cuda::GpuMat& src, cv::cuda::GpuMat& dst, const cv::Vec3i& offset, const cv::Vec3i& srcRes,
                const cv::Vec3i& dstRes) {
    CV_Assert(src.depth() == CV_32F && src.type() == dst.type() && src.channels() <= 3);
    ok(emf_copy_values((const float*)src.data, (float*)dst.data, src.channels(), offset.val, srcRes.val, dstRes.val,
                       /*stream=*/nullptr));   // the reference launches it on the default stream
}

}  // namespace TSDF

namespace ObjTSDF {

void updateFgBgProbs(const cv::cuda::GpuMat& mask, const cv::cuda::GpuMat& occluded_mask, const cv::cuda::GpuMat& tsdfVol,
                     const cv::cuda::GpuMat& tsdfWeights, cv::cuda::GpuMat& fgBgProbs, const cv::Matx33f& rel_rot,
                     const cv::Vec3f& rel_trans, const cv::Matx33f& intr, const cv::Vec3i& volumeRes,
                     const float voxelSize, cv::cuda::Stream& stream) {
    const emf_image m = img(mask), o = img(occluded_mask);
    const emf_pose T = pose(rel_rot, rel_trans);
    ok(emf_update_fgbg_probs(&m, &o, (const float*)tsdfVol.data, (const float*)tsdfWeights.data, (float*)fgBgProbs.data,
                             &T, intr.val, volumeRes.val, voxelSize, str(stream)));
}

}  // namespace ObjTSDF

#ifdef EMF_B200_BINDING_HAS_FRAME_OPS
namespace EMFusion {

// The reference launcher ends with cudaDeviceSynchronize() (src/core/cuda/EMFusion.cu:60); the callers do not
// rely on it (everything downstream is stream-ordered on the default stream), so it is not reproduced.
void computePoints(const cv::cuda::GpuMat& depth, cv::cuda::GpuMat& points, const cv::Matx33f& params) {
    const emf_image d = img(depth), p = img(points);
    ok(emf_compute_points(&d, &p, params.val, nullptr));
}

}  // namespace EMFusion
#endif

}  // namespace cuda
}  // namespace emf
