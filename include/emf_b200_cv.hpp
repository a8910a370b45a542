// emf_b200_cv.hpp -- emf::TSDF / emf::ObjTSDF with the reference's OWN signatures, over libemf_b200.so.
//
// This is the class-level drop-in of BASELINE.json's north_star ("keeps the C++ TSDF / ObjTSDF class surface so
// src/core/EMFusion.cpp drops it in unchanged"): a maintainer of EM-Fusion puts this header in place of
//   include/EMFusion/core/TSDF.h:39-328      (class emf::TSDF)
//   include/EMFusion/core/ObjTSDF.h:33-217   (class emf::ObjTSDF)
// (an overlay include directory whose EMFusion/core/TSDF.h and ObjTSDF.h just include this file), drops src/core/TSDF.cpp,
// src/core/ObjTSDF.cpp and src/core/cuda/*.cu from the build and links libemf_b200.so.  Every public method
// src/core/EMFusion.cpp calls (reference src/core/EMFusion.cpp:30-32, 60, 545-547, 642-645, 673-722, 747-750, 830, 856,
// 866-883, 902, 927 and the getters) exists here with the same name, argument types, constness, default arguments and
// non-virtual hiding (ObjTSDF redeclares raycast / computeAssociation / syncTrack; only reset, getMesh, getTSDF,
// getWeightsVol are virtual).  tests/test_cv_adapter_compiles.py compiles this header next to the reference's own class
// headers and static_asserts that the member-function types agree, one by one.
//
// Types: cv::cuda::GpuMat (data, step, rows, cols), cv::Affine3f, cv::Matx33f, cv::Vec3f / Vec3i, cv::Size, cv::Mat,
// cv::viz::Mesh, cv::cuda::Stream (cv::cuda::StreamAccessor::getStream), emf::TSDFParams (the reference's data.h, kept).
// Needs OpenCV with the cuda modules (any 4.x); in this repo's image, which has none, the compile test uses the type
// stand-in of oracle/shim_full/.
//
// What a call does: exactly one C-ABI call of include/emf_b200.h on the caller's stream (asynchronous, like the
// reference's level-1 operators).  So an UNCHANGED EMFusion.cpp works, with the reference's launch structure -- one call per
// volume and stage, its OpenCV element-wise chains for the normaliser and the composite in between, its host barriers --
// and gets the per-kernel speed-ups only; the batched frame path (one launch per stage for all volumes, no host barrier:
// the 7x of DESIGN.md) needs the three-method patch of EMFusion.cpp in INTEGRATION.md section 2, because the calls it
// fuses are separated by OpenCV calls of the reference's own that a class adapter cannot defer.
// Tracking: the reference drives nine methods per iteration (computeGradients ... computePoseUpdate, src/core/EMFusion.cpp:
// 674-684); here computePoseUpdate launches ONE fused Levenberg-Marquardt iteration on the device (emf_track_iterate) with
// the points / association image the earlier calls of the iteration named, the other eight are bookkeeping, and syncTrack
// reads the pose back.
#ifndef EMF_B200_CV_HPP
#define EMF_B200_CV_HPP

#if !__has_include(<opencv2/core/cuda.hpp>)
#error "emf_b200_cv.hpp needs OpenCV's cuda module headers (or, for the compile test, oracle/shim_full on the include path)"
#endif
#include <opencv2/opencv.hpp>
#include <opencv2/viz.hpp>
#include <opencv2/core/cuda.hpp>
#include <opencv2/core/cuda_stream_accessor.hpp>

#include <memory>
#include <vector>

#include "EMFusion/core/data.h"   // emf::TSDFParams (the reference's own parameter struct stays)
#include "emf_b200.hpp"

#ifndef EMF_B200_CV_NAMESPACE
#define EMF_B200_CV_NAMESPACE emf   // (the compile test puts the classes elsewhere to compare them with the reference's)
#endif

namespace EMF_B200_CV_NAMESPACE {

namespace b200_detail {
inline emfb::Affine affine(const cv::Affine3f& a) {
    emfb::Affine r;
    const cv::Matx33f R = a.rotation();
    const cv::Vec3f t = a.translation();
    for (int k = 0; k < 9; ++k) r.R[k] = R.val[k];
    for (int k = 0; k < 3; ++k) r.t[k] = t.val[k];
    return r;
}
inline cv::Affine3f affine(const emfb::Affine& a) {
    return cv::Affine3f(cv::Matx33f(a.R[0], a.R[1], a.R[2], a.R[3], a.R[4], a.R[5], a.R[6], a.R[7], a.R[8]), cv::Vec3f(a.t[0], a.t[1], a.t[2]));
}
inline emfb::Matx33f matx(const cv::Matx33f& m) {
    emfb::Matx33f r;
    for (int k = 0; k < 9; ++k) r[k] = m.val[k];
    return r;
}
template <typename T>
inline emfb::Image<T> view(const cv::cuda::GpuMat& m) {      // non-owning view of the caller's image
    emfb::Image<T> im;
    im.width = m.cols; im.height = m.rows; im.channels = m.channels();
    im.external = emf_image{(void*)m.data, m.step, m.cols, m.rows};
    return im;
}
inline cudaStream_t str(cv::cuda::Stream& s) { return cv::cuda::StreamAccessor::getStream(s); }
inline emfb::TSDFParams params(const ::emf::TSDFParams& p) {
    emfb::TSDFParams q;
    q.maxTSDFWeight = p.maxTSDFWeight; q.assocSigma = p.assocSigma; q.alpha = p.alpha; q.uniPrior = p.uniPrior;
    q.tau = p.tau; q.eps1 = p.eps1; q.eps2 = p.eps2; q.nu_init = p.nu_init; q.huberThresh = p.huberThresh;
    return q;
}
// the state of one volume's tracker between the reference's per-iteration method calls
struct TrackRun {
    emfb::DeviceArray<emf_track_state> state{1};
    emfb::DeviceArray<float> records{EMF_TRACK_RECORD};
    emfb::DeviceArray<unsigned char> ws;
    emf_image points{}, assoc{};
    cudaStream_t stream = nullptr;
};
}  // namespace b200_detail

class TSDF {
public:
    TSDF(cv::Vec3i _volumeRes, const float _voxelSize, const float _truncdist, cv::Affine3f _pose, ::emf::TSDFParams _params,
         cv::Size frameSize)
        : params(_params),
          impl_(std::make_shared<emfb::TSDF>(emfb::Vec3i{_volumeRes.val[0], _volumeRes.val[1], _volumeRes.val[2]}, _voxelSize, _truncdist,
                                             b200_detail::affine(_pose), b200_detail::params(_params), frameSize.width, frameSize.height)) {}
    virtual ~TSDF() = default;

    virtual void reset(const cv::Affine3f& _pose) { impl_->reset(b200_detail::affine(_pose)); }
    void getCorners(cv::Vec3f& low, cv::Vec3f& high) const {
        emfb::Vec3f l, h;
        impl_->getCorners(l, h);
        low = cv::Vec3f(l[0], l[1], l[2]); high = cv::Vec3f(h[0], h[1], h[2]);
    }
    cv::Vec3f getVolumeSize() const { const emfb::Vec3f s = impl_->getVolumeSize(); return cv::Vec3f(s[0], s[1], s[2]); }
    cv::Vec3i getVolumeRes() const { const emfb::Vec3i r = impl_->getVolumeRes(); return cv::Vec3i(r[0], r[1], r[2]); }
    float getVoxelSize() const { return impl_->getVoxelSize(); }
    float getTruncDist() const { return impl_->getTruncDist(); }
    cv::Affine3f getPose() const { return b200_detail::affine(impl_->getPose()); }

    void integrate(const cv::cuda::GpuMat& depth, const cv::cuda::GpuMat& weights, const cv::Affine3f& cam_pose,
                   const cv::Matx33f& intr, cv::cuda::Stream& stream = cv::cuda::Stream::Null()) {
        impl_->integrate(b200_detail::view<float>(depth), b200_detail::view<float>(weights), b200_detail::affine(cam_pose),
                         b200_detail::matx(intr), b200_detail::str(stream));
    }
    void updateGradients(cv::cuda::Stream& stream = cv::cuda::Stream::Null()) { impl_->updateGradients(b200_detail::str(stream)); }
    void raycast(const cv::Affine3f& cam_pose, const cv::Matx33f& intr, cv::cuda::GpuMat& raylengths, cv::cuda::GpuMat& vertices,
                 cv::cuda::GpuMat& normals, cv::cuda::GpuMat& mask, cv::cuda::Stream& stream = cv::cuda::Stream::Null()) {
        auto r = b200_detail::view<float>(raylengths), v = b200_detail::view<float>(vertices), n = b200_detail::view<float>(normals);
        auto m = b200_detail::view<uint8_t>(mask);
        impl_->raycast(b200_detail::affine(cam_pose), b200_detail::matx(intr), r, v, n, m, b200_detail::str(stream));
    }
    void computeAssociation(const cv::cuda::GpuMat& points, const cv::Affine3f& cam_pose, cv::cuda::GpuMat& associationWeights,
                            cv::cuda::Stream& stream = cv::cuda::Stream::Null()) {
        auto a = b200_detail::view<float>(associationWeights);
        impl_->computeAssociation(b200_detail::view<float>(points), b200_detail::affine(cam_pose), a, b200_detail::str(stream));
    }

    // ---- tracker (src/core/TSDF.cpp:170-338)
    void prepareTracking(const cv::Affine3f& cam_pose, cv::cuda::Stream& stream = cv::cuda::Stream::Null()) {
        impl_->prepareTracking(b200_detail::affine(cam_pose));
        beginTrack(b200_detail::str(stream));
    }
    void computeGradients(const cv::cuda::GpuMat& points) { run_->points = b200_detail::view<float>(points).c(); }
    void computeTSDFVals(const cv::cuda::GpuMat& points) { run_->points = b200_detail::view<float>(points).c(); }
    void computeTSDFWeights(const cv::cuda::GpuMat& points) { run_->points = b200_detail::view<float>(points).c(); }
    void computeHuberWeights() {}
    void normalizeTSDFWeights() {}
    void combineWeights(const cv::cuda::GpuMat& associationWeights) { run_->assoc = b200_detail::view<float>(associationWeights).c(); }
    void computeHessians() {}
    void reduceHessians() {}
    void computePoseUpdate(const cv::cuda::GpuMat& points) {
        run_->points = b200_detail::view<float>(points).c();
        trackIteration();
    }
    void syncTrack(cv::Affine3f& cam_pose) {
        endTrack();
        emfb::Affine c;
        impl_->syncTrack(c);
        cam_pose = b200_detail::affine(c);
    }
    void getHuberWeights(cv::Mat& weights) const { weights = cv::Mat(); }        // (diagnostic images of saveOutput: not kept by the fused tracker)
    void getTrackingWeights(cv::Mat& weights) const { weights = cv::Mat(); }

    virtual cv::viz::Mesh getMesh() { return meshOf(*impl_); }
    virtual cv::Mat getTSDF() const { return volumeMat(impl_->getTSDF()); }
    virtual cv::Mat getWeightsVol() const { return volumeMat(impl_->getWeightsVol()); }

protected:
    explicit TSDF(std::shared_ptr<emfb::TSDF> impl, const ::emf::TSDFParams& p) : params(p), impl_(std::move(impl)) {}
    static cv::viz::Mesh meshOf(const emfb::TSDF& t) {          // 1 x n CV_32FC3 / CV_32FC3 / 1 x m CV_32SC1, as the reference downloads them
        const emfb::TSDF::Mesh h = t.getMesh();
        cv::viz::Mesh m;
        if (h.cloud.empty()) return m;
        m.cloud = cv::Mat(1, (int)(h.cloud.size() / 3), CV_32FC3);
        m.normals = cv::Mat(1, (int)(h.normals.size() / 3), CV_32FC3);
        m.polygons = cv::Mat(1, (int)h.polygons.size(), CV_32SC1);
        std::memcpy(m.cloud.ptr<float>(), h.cloud.data(), h.cloud.size() * sizeof(float));
        std::memcpy(m.normals.ptr<float>(), h.normals.data(), h.normals.size() * sizeof(float));
        std::memcpy(m.polygons.ptr<int>(), h.polygons.data(), h.polygons.size() * sizeof(int32_t));
        return m;
    }
    cv::Mat volumeMat(const std::vector<float>& v) const {
        const emfb::Vec3i r = impl_->getVolumeRes();
        cv::Mat m(r[1] * r[2], r[0], CV_32FC1);          // rows = Ry * Rz, cols = Rx (src/core/TSDF.cpp:35-42)
        std::memcpy(m.ptr<float>(), v.data(), v.size() * sizeof(float));
        return m;
    }
    void beginTrack(cudaStream_t stream) {
        run_ = std::make_shared<b200_detail::TrackRun>();
        run_->stream = stream;
        const size_t wsb = emf_track_workspace_bytes(1);
        run_->ws.allocate(wsb);
        emf_track_state st{};
        const emfb::Affine& T = impl_->relPoseCO();
        for (int k = 0; k < 9; ++k) st.R[k] = T.R[k];
        for (int k = 0; k < 3; ++k) st.t[k] = T.t[k];
        st.nu = params.nu_init; st.first_iteration = 1; st.evaluate_gradient = 1;
        emfb::cu(cudaMemcpyAsync(run_->state.data(), &st, sizeof st, cudaMemcpyHostToDevice, stream), "tracker state upload");
        emfb::ok(emf_track_workspace_init(run_->ws.data(), wsb, (emf_stream_t)stream), "emf_track_workspace_init");
    }
    void trackIteration() {
        const emf_volume v = impl_->descriptor();
        const emf_pose hint = impl_->relPoseCO().c();
        const emf_track_lm_params lm{params.tau, params.eps1, params.eps2, params.nu_init, params.huberThresh, params.maxTSDFWeight};
        emfb::Image<float>& iw = impl_->trackingWeightsImage();
        const emf_image iwc = iw.c();
        emfb::ok(emf_track_iterate(1, &v, run_->state.data(), &hint, &run_->points, nullptr, &run_->assoc, &lm, &iwc, run_->records.data(),
                                   run_->ws.data(), run_->ws.size(), 1, (emf_stream_t)run_->stream), "emf_track_iterate");
    }
    void endTrack() {
        if (!run_) return;
        emf_track_state st{};
        emfb::cu(cudaMemcpyAsync(&st, run_->state.data(), sizeof st, cudaMemcpyDeviceToHost, run_->stream), "tracker state download");
        emfb::cu(cudaStreamSynchronize(run_->stream), "sync");
        emfb::Affine T;
        for (int k = 0; k < 9; ++k) T.R[k] = (float)st.R[k];
        for (int k = 0; k < 3; ++k) T.t[k] = (float)st.t[k];
        impl_->setRelPoseCO(T);
        impl_->trackingConverged = st.converged != 0;
        run_.reset();
    }

    ::emf::TSDFParams params;
    std::shared_ptr<emfb::TSDF> impl_;                  // shared storage: copies of the object are shallow, like GpuMat's
    std::shared_ptr<b200_detail::TrackRun> run_;
};

class ObjTSDF : public TSDF {
public:
    ObjTSDF(cv::Vec3i _volumeRes, const float _voxelSize, const float _truncdist, cv::Affine3f _pose, ::emf::TSDFParams _params,
            cv::Size frameSize)
        : TSDF(std::make_shared<emfb::ObjTSDF>(emfb::Vec3i{_volumeRes.val[0], _volumeRes.val[1], _volumeRes.val[2]}, _voxelSize, _truncdist,
                                               b200_detail::affine(_pose), b200_detail::params(_params), frameSize.width, frameSize.height),
               _params) {}

    bool operator==(const ObjTSDF& other) const { return getID() == other.getID(); }
    bool operator!=(const ObjTSDF& other) const { return !(*this == other); }
    const int getID() const { return obj()->getID(); }

    virtual void reset(const cv::Affine3f& _pose) override { obj()->reset(b200_detail::affine(_pose)); }
    float getExProb() { return obj()->getExProb(); }
    void updateExProb(const bool exists) { obj()->updateExProb(exists); }
    void updateClassProbs(const std::vector<double>& _classProbs) {
        if (classProbs.size() != _classProbs.size()) classProbs.assign(_classProbs.size(), 0.0);     // src/core/ObjTSDF.cpp:68-78
        for (size_t i = 0; i < classProbs.size(); ++i) classProbs[i] += _classProbs[i];
    }
    cv::Vec3f resize(const cv::Vec3f& p10, const cv::Vec3f& p90, const float volPad) {
        const emfb::Vec3f o = obj()->resize(emfb::Vec3f{p10.val[0], p10.val[1], p10.val[2]}, emfb::Vec3f{p90.val[0], p90.val[1], p90.val[2]}, volPad);
        return cv::Vec3f(o[0], o[1], o[2]);
    }
    void integrateMask(const cv::cuda::GpuMat& mask, const cv::cuda::GpuMat& occluded_mask, const cv::Affine3f& cam_pose,
                       const cv::Matx33f& intr, cv::cuda::Stream& stream = cv::cuda::Stream::Null()) {
        obj()->integrateMask(b200_detail::view<uint8_t>(mask), b200_detail::view<uint8_t>(occluded_mask), b200_detail::affine(cam_pose),
                             b200_detail::matx(intr), b200_detail::str(stream));
    }
    void computeAssociation(const cv::cuda::GpuMat& points, const cv::Affine3f& cam_pose, cv::cuda::GpuMat& associationWeights,
                            cv::cuda::Stream& stream = cv::cuda::Stream::Null()) {
        auto a = b200_detail::view<float>(associationWeights);
        obj()->computeAssociation(b200_detail::view<float>(points), b200_detail::affine(cam_pose), a, b200_detail::str(stream));
    }
    void raycast(const cv::Affine3f& cam_pose, const cv::Matx33f& intr, cv::cuda::GpuMat& raylengths, cv::cuda::GpuMat& vertices,
                 cv::cuda::GpuMat& normals, cv::cuda::GpuMat& mask, cv::cuda::Stream& stream = cv::cuda::Stream::Null()) {
        auto r = b200_detail::view<float>(raylengths), v = b200_detail::view<float>(vertices), n = b200_detail::view<float>(normals);
        auto m = b200_detail::view<uint8_t>(mask);
        obj()->raycast(b200_detail::affine(cam_pose), b200_detail::matx(intr), r, v, n, m, b200_detail::str(stream));
    }
    void syncTrack(const cv::Affine3f& cam_pose) {
        endTrack();
        obj()->syncTrack(b200_detail::affine(cam_pose));
    }
    void getFgProbVals(cv::Mat& vals) const { vals = cv::Mat(); }
    int getClassID() const {
        int best = 0;
        for (size_t i = 1; i < classProbs.size(); ++i) if (classProbs[i] > classProbs[best]) best = (int)i;
        return best;
    }
    virtual cv::viz::Mesh getMesh() override { return meshOf(*impl_); }     // (descriptor() carries fgProbs: the fgVolMask test)
    cv::Mat getFgProbVol() { return volumeMat(obj()->getFgProbVol()); }
    cv::cuda::GpuMat getFgVolMask() { return cv::cuda::GpuMat(); }      // (not materialised: the fgProb > 0.5 test runs inside the raycast)

private:
    emfb::ObjTSDF* obj() const { return static_cast<emfb::ObjTSDF*>(impl_.get()); }
    std::vector<double> classProbs;
};

}  // namespace EMF_B200_CV_NAMESPACE

#endif  // EMF_B200_CV_HPP
