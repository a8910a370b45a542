// emf_b200.hpp -- C++ host mirror of EM-Fusion's volume classes over the C ABI (include/emf_b200.h).
//
// The reference's host side is C++ on OpenCV-CUDA / Eigen / Sophus, none of which exist in this image; this header is
// the same class surface without them: emfb::TSDF (reference include/EMFusion/core/TSDF.h:39-328, src/core/TSDF.cpp) and
// emfb::ObjTSDF (include/EMFusion/core/ObjTSDF.h:33-217, src/core/ObjTSDF.cpp) with the reference's method names,
// argument order and meaning -- integrate, updateGradients, raycast, computeAssociation, prepareTracking ... syncTrack,
// integrateMask, computeFgProbs, resize, the getters -- where cv::cuda::GpuMat becomes emfb::Image / emfb::DeviceArray
// (device pointer + pitch, RAII), cv::Affine3f becomes emfb::Affine (float, same composition order as the reference:
// rel_pose = cam_pose.inv() * pose, src/core/TSDF.cpp:112), cv::cuda::Stream becomes cudaStream_t.  Header-only, C++17,
// needs the CUDA runtime for allocations and libemf_b200.so for every kernel; nothing here computes on the host except
// pose algebra.  Non-virtual hiding as in the reference: ObjTSDF redeclares raycast / computeAssociation / syncTrack.
//
// Tracking: the nine per-iteration methods of the reference (computeGradients ... computePoseUpdate) are one call here,
// track(), which runs the whole Levenberg-Marquardt loop on the device (emf_track_iterate); prepareTracking / syncTrack
// keep their meaning.
#ifndef EMF_B200_HPP
#define EMF_B200_HPP

#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "emf_b200.h"

namespace emfb {

inline void ok(int rc, const char* what) {
    if (rc != EMF_OK) throw std::runtime_error(std::string(what) + " failed: " + std::to_string(rc));
}
inline void cu(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

using Vec3f = std::array<float, 3>;
using Vec3i = std::array<int, 3>;
using Matx33f = std::array<float, 9>;   // row-major, as cv::Matx33f::val

// cv::Affine3f: x' = R x + t, float arithmetic
struct Affine {
    Matx33f R{1, 0, 0, 0, 1, 0, 0, 0, 1};
    Vec3f t{0, 0, 0};
    static Affine translation(float x, float y, float z) { Affine a; a.t = {x, y, z}; return a; }
    Vec3f rotate(const Vec3f& v) const {
        return {R[0] * v[0] + R[1] * v[1] + R[2] * v[2], R[3] * v[0] + R[4] * v[1] + R[5] * v[2], R[6] * v[0] + R[7] * v[1] + R[8] * v[2]};
    }
    Affine inv() const {
        Affine r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.R[3 * i + j] = R[3 * j + i];
        const Vec3f q = r.rotate(t);
        r.t = {-q[0], -q[1], -q[2]};
        return r;
    }
    Affine operator*(const Affine& o) const {   // this * o: apply o first
        Affine r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.R[3 * i + j] = R[3 * i] * o.R[j] + R[3 * i + 1] * o.R[3 + j] + R[3 * i + 2] * o.R[6 + j];
        const Vec3f q = rotate(o.t);
        r.t = {q[0] + t[0], q[1] + t[1], q[2] + t[2]};
        return r;
    }
    emf_pose c() const {
        emf_pose p;
        std::memcpy(p.R, R.data(), sizeof p.R);
        std::memcpy(p.t, t.data(), sizeof p.t);
        return p;
    }
};

// a device allocation (cv::cuda::GpuMat's storage): zero-initialised, freed with the object, movable
template <typename T>
class DeviceArray {
public:
    DeviceArray() = default;
    explicit DeviceArray(size_t n) { allocate(n); }
    DeviceArray(const DeviceArray&) = delete;
    DeviceArray& operator=(const DeviceArray&) = delete;
    DeviceArray(DeviceArray&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    DeviceArray& operator=(DeviceArray&& o) noexcept {
        if (this != &o) { release(); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; }
        return *this;
    }
    ~DeviceArray() { release(); }
    void allocate(size_t n) {
        release();
        cu(cudaMalloc((void**)&p_, std::max<size_t>(n, 1) * sizeof(T)), "cudaMalloc");
        n_ = n;
        setZero();
    }
    void setZero(cudaStream_t s = nullptr) { if (p_) cu(cudaMemsetAsync(p_, 0, n_ * sizeof(T), s), "cudaMemset"); }
    T* data() const { return p_; }
    size_t size() const { return n_; }
    void upload(const T* host, size_t n) { cu(cudaMemcpy(p_, host, n * sizeof(T), cudaMemcpyHostToDevice), "upload"); }
    std::vector<T> download() const {
        std::vector<T> h(n_);
        cu(cudaMemcpy(h.data(), p_, n_ * sizeof(T), cudaMemcpyDeviceToHost), "download");
        return h;
    }

private:
    void release() { if (p_) cudaFree(p_); p_ = nullptr; n_ = 0; }
    T* p_ = nullptr;
    size_t n_ = 0;
};

// a W x H device image with `channels` elements of T per pixel (continuous)
template <typename T>
struct Image {
    DeviceArray<T> mem;
    int width = 0, height = 0, channels = 1;
    Image() = default;
    Image(int w, int h, int ch = 1) : mem((size_t)w * h * ch), width(w), height(h), channels(ch) {}
    emf_image c() const { return emf_image{(void*)mem.data(), (size_t)width * channels * sizeof(T), width, height}; }
};

// include/EMFusion/core/data.h:32-71
struct TSDFParams {
    float tau = 1e3f, eps1 = 1e-8f, eps2 = 1e-8f, nu_init = 2.0f, huberThresh = 0.2f, maxTSDFWeight = 64.0f;
    float assocSigma = 0.02f, alpha = 0.8f, uniPrior = 1.0f;
    emf_tsdf_params c() const { return emf_tsdf_params{maxTSDFWeight, assocSigma, alpha, uniPrior}; }
};

class TSDF {
public:
    // src/core/TSDF.cpp:29-72
    TSDF(Vec3i _volumeRes, float _voxelSize, float _truncdist, const Affine& _pose, const TSDFParams& _params, int frameW,
         int frameH)
        : params(_params), volumeRes(_volumeRes), voxelSize(_voxelSize), truncdist(_truncdist), pose(_pose), frameW_(frameW),
          frameH_(frameH), tsdfVol(numVoxels()), tsdfWeights(numVoxels()), intWeights(frameW, frameH) {}
    virtual ~TSDF() = default;

    virtual void reset(const Affine& _pose) {   // :74-79
        tsdfVol.setZero();
        tsdfWeights.setZero();
        tsdfGrads = DeviceArray<float>();
        pose = _pose;
    }
    void getCorners(Vec3f& low, Vec3f& high) const {
        for (int k = 0; k < 3; ++k) { high[k] = (volumeRes[k] - 1) * voxelSize / 2; low[k] = -high[k]; }
    }
    Vec3f getVolumeSize() const { return {volumeRes[0] * voxelSize, volumeRes[1] * voxelSize, volumeRes[2] * voxelSize}; }
    Vec3i getVolumeRes() const { return volumeRes; }
    float getVoxelSize() const { return voxelSize; }
    float getTruncDist() const { return truncdist; }
    Affine getPose() const { return pose; }
    size_t numVoxels() const { return (size_t)volumeRes[0] * volumeRes[1] * volumeRes[2]; }

    // :108-118
    void integrate(const Image<float>& depth, const Image<float>& weights, const Affine& cam_pose, const Matx33f& intr,
                   cudaStream_t stream = nullptr) {
        const emf_pose T = (cam_pose.inv() * pose).c();
        const emf_image d = depth.c(), w = weights.c();
        ok(emf_update_tsdf(&d, &w, tsdfVol.data(), tsdfWeights.data(), &T, intr.data(), volumeRes.data(), voxelSize, truncdist,
                           params.maxTSDFWeight, (emf_stream_t)stream), "emf_update_tsdf");
        gradsValid_ = false;
    }
    // :120-123 -- the raycast and the tracker take forward differences on the fly, so this only has to run for a consumer
    // that wants the float3 volume (getGrads)
    void updateGradients(cudaStream_t stream = nullptr) { (void)stream; gradsValid_ = false; }
    const float* getGrads(cudaStream_t stream = nullptr) {
        if (tsdfGrads.size() != 3 * numVoxels()) { tsdfGrads.allocate(3 * numVoxels()); gradsValid_ = false; }
        if (!gradsValid_) {
            ok(emf_compute_tsdf_grads(tsdfVol.data(), tsdfGrads.data(), volumeRes.data(), (emf_stream_t)stream), "emf_compute_tsdf_grads");
            gradsValid_ = true;
        }
        return tsdfGrads.data();
    }
    // :158-168
    void raycast(const Affine& cam_pose, const Matx33f& intr, Image<float>& raylengths, Image<float>& vertices,
                 Image<float>& normals, Image<uint8_t>& mask, cudaStream_t stream = nullptr) {
        raycastImpl(nullptr, cam_pose, intr, raylengths, vertices, normals, mask, stream);
    }
    // :125-136
    void computeAssociation(const Image<float>& points, const Affine& cam_pose, Image<float>& associationWeights,
                            cudaStream_t stream = nullptr) {
        associationImpl(nullptr, points, cam_pose, associationWeights, stream);
    }

    // :170-192
    void prepareTracking(const Affine& cam_pose) { rel_pose_CO = orthonormalised(pose.inv() * cam_pose); }
    // computeGradients ... computePoseUpdate for up to maxTrackingIter iterations (:194-338), on the device
    int track(const Image<float>& points, const Image<float>& associationWeights, const Matx33f& intr, int maxTrackingIter = 100,
              cudaStream_t stream = nullptr) {
        emf_track_state st{};
        for (int k = 0; k < 9; ++k) st.R[k] = rel_pose_CO.R[k];
        for (int k = 0; k < 3; ++k) st.t[k] = rel_pose_CO.t[k];
        st.nu = params.nu_init; st.first_iteration = 1; st.evaluate_gradient = 1;
        DeviceArray<emf_track_state> dst(1);
        DeviceArray<float> records(EMF_TRACK_RECORD);
        const size_t wsb = emf_track_workspace_bytes(1);
        DeviceArray<unsigned char> ws(wsb);
        cu(cudaMemcpyAsync(dst.data(), &st, sizeof st, cudaMemcpyHostToDevice, stream), "state upload");
        ok(emf_track_workspace_init(ws.data(), wsb, (emf_stream_t)stream), "emf_track_workspace_init");
        emf_volume v = cVolume(nullptr);
        const emf_pose hint = rel_pose_CO.c();
        const emf_image p = points.c(), a = associationWeights.c(), iw = intWeights.c();
        const emf_track_lm_params lm{params.tau, params.eps1, params.eps2, params.nu_init, params.huberThresh, params.maxTSDFWeight};
        int done = 0;
        while (done < maxTrackingIter) {
            const int k = std::min(8, maxTrackingIter - done);
            ok(emf_track_iterate(1, &v, dst.data(), &hint, &p, intr.data(), &a, &lm, &iw, records.data(), ws.data(), wsb, k,
                                 (emf_stream_t)stream), "emf_track_iterate");
            done += k;
            cu(cudaMemcpyAsync(&st, dst.data(), sizeof st, cudaMemcpyDeviceToHost, stream), "state download");
            cu(cudaStreamSynchronize(stream), "sync");
            if (st.converged) break;
        }
        for (int k = 0; k < 9; ++k) rel_pose_CO.R[k] = (float)st.R[k];
        for (int k = 0; k < 3; ++k) rel_pose_CO.t[k] = (float)st.t[k];
        trackingConverged = st.converged != 0;
        return st.iterations;
    }
    // :333-338
    void syncTrack(Affine& cam_pose) { cam_pose = pose * rel_pose_CO; }

    virtual std::vector<float> getTSDF() const { return tsdfVol.download(); }
    virtual std::vector<float> getWeightsVol() const { return tsdfWeights.download(); }

    TSDFParams params;
    bool trackingConverged = false;

protected:
    emf_volume cVolume(const float* fg) const {
        emf_volume v{};
        v.tsdf = tsdfVol.data(); v.weights = tsdfWeights.data(); v.fg_probs = fg;
        v.res[0] = volumeRes[0]; v.res[1] = volumeRes[1]; v.res[2] = volumeRes[2];
        v.voxel_size = voxelSize; v.truncdist = truncdist; v.id = id_;
        return v;
    }
    void raycastImpl(const float* fg, const Affine& cam_pose, const Matx33f& intr, Image<float>& raylengths, Image<float>& vertices,
                     Image<float>& normals, Image<uint8_t>& mask, cudaStream_t stream) {
        const emf_pose T = (pose.inv() * cam_pose).c();
        const emf_image r = raylengths.c(), v = vertices.c(), n = normals.c(), m = mask.c();
        ok(emf_raycast_tsdf(tsdfVol.data(), nullptr, tsdfWeights.data(), fg, &r, &v, &n, &m, &T, intr.data(), volumeRes.data(),
                            voxelSize, truncdist, nullptr, (emf_stream_t)stream), "emf_raycast_tsdf");
    }
    void associationImpl(const float* fg, const Image<float>& points, const Affine& cam_pose, Image<float>& out, cudaStream_t stream) {
        const emf_pose T = (pose.inv() * cam_pose).c();
        const emf_volume v = cVolume(fg);
        const emf_tsdf_params prm = params.c();
        const emf_image p = points.c(), o = out.c();
        ok(emf_compute_association(&v, &p, &T, &prm, &o, nullptr, (emf_stream_t)stream), "emf_compute_association");
    }
    // Q of the QR decomposition of the rotation block, columns flipped where diag(R) < 0 (:174-181): Gram-Schmidt on the columns
    static Affine orthonormalised(const Affine& T) {
        double c[3][3];
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) c[j][i] = T.R[3 * i + j];
        for (int j = 0; j < 3; ++j) {
            for (int k = 0; k < j; ++k) {
                const double d = c[j][0] * c[k][0] + c[j][1] * c[k][1] + c[j][2] * c[k][2];
                for (int i = 0; i < 3; ++i) c[j][i] -= d * c[k][i];
            }
            const double n = std::sqrt(c[j][0] * c[j][0] + c[j][1] * c[j][1] + c[j][2] * c[j][2]);
            for (int i = 0; i < 3; ++i) c[j][i] /= n;
        }
        Affine r = T;
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) r.R[3 * i + j] = (float)c[j][i];
        return r;
    }

    Vec3i volumeRes;
    float voxelSize, truncdist;
    Affine pose, rel_pose_CO;
    int frameW_, frameH_;
    int id_ = 0;
    DeviceArray<float> tsdfVol, tsdfWeights, tsdfGrads;
    Image<float> intWeights;
    bool gradsValid_ = false;
};

class ObjTSDF : public TSDF {
public:
    // src/core/ObjTSDF.cpp:30-54; ids from a static counter incremented only here (:28,34)
    ObjTSDF(Vec3i _volumeRes, float _voxelSize, float _truncdist, const Affine& _pose, const TSDFParams& _params, int frameW, int frameH)
        : TSDF(_volumeRes, _voxelSize, _truncdist, _pose, _params, frameW, frameH), fgBgProbs(2 * numVoxels()), fgProbs(numVoxels()),
          fgBox(6) {
        id_ = ++nextID();
        resetBox();
    }
    static int& nextID() { static int n = 0; return n; }
    bool operator==(const ObjTSDF& o) const { return id_ == o.id_; }
    bool operator!=(const ObjTSDF& o) const { return id_ != o.id_; }
    int getID() const { return id_; }

    void reset(const Affine& _pose) override {   // :56-59
        TSDF::reset(_pose);
        fgBgProbs.setZero();
        fgProbs.setZero();
        resetBox();
    }
    float getExProb() const { return (float)exCount / (float)(exCount + nonExCount); }
    void updateExProb(bool exists) { exCount += exists; nonExCount += !exists; }

    // :167-179
    void integrateMask(const Image<uint8_t>& mask, const Image<uint8_t>& occluded_mask, const Affine& cam_pose, const Matx33f& intr,
                       cudaStream_t stream = nullptr) {
        const emf_pose T = (cam_pose.inv() * pose).c();
        const emf_image m = mask.c(), o = occluded_mask.c();
        ok(emf_update_fgbg_probs(&m, &o, tsdfVol.data(), tsdfWeights.data(), fgBgProbs.data(), &T, intr.data(), volumeRes.data(),
                                 voxelSize, (emf_stream_t)stream), "emf_update_fgbg_probs");
        computeFgProbs(stream);
    }
    // :218-226
    void computeFgProbs(cudaStream_t stream = nullptr) {
        ok(emf_compute_fg_probs_box(fgBgProbs.data(), volumeRes.data(), fgProbs.data(), nullptr, fgBox.data(), (emf_stream_t)stream),
           "emf_compute_fg_probs_box");
    }
    // :181-201 (hides TSDF::computeAssociation, as in the reference)
    void computeAssociation(const Image<float>& points, const Affine& cam_pose, Image<float>& associationWeights,
                            cudaStream_t stream = nullptr) {
        associationImpl(fgProbs.data(), points, cam_pose, associationWeights, stream);
    }
    // :203-216 (hides TSDF::raycast): the fgProb > 0.5 weight mask is applied inside the kernel
    void raycast(const Affine& cam_pose, const Matx33f& intr, Image<float>& raylengths, Image<float>& vertices, Image<float>& normals,
                 Image<uint8_t>& mask, cudaStream_t stream = nullptr) {
        raycastImpl(fgProbs.data(), cam_pose, intr, raylengths, vertices, normals, mask, stream);
    }
    // :228-235 (hides TSDF::syncTrack)
    void syncTrack(const Affine& cam_pose) { pose = cam_pose * rel_pose_CO.inv(); }

    // :80-165
    Vec3f resize(const Vec3f& p10, const Vec3f& p90, float volPad, cudaStream_t stream = nullptr) {
        bool contained = true;
        for (int i = 0; i < 3; ++i) {
            const float hi = ((float)volumeRes[i] - 1.0f) / 2.0f * voxelSize;
            if (p10[i] < -hi || p90[i] > hi) { contained = false; break; }
        }
        if (contained) return {0, 0, 0};
        Vec3f newCenter;
        Vec3i pixOffset;
        for (int i = 0; i < 3; ++i) {
            newCenter[i] = (p10[i] + p90[i]) / 2.0f;
            pixOffset[i] = (int)std::nearbyint(newCenter[i] / voxelSize);     // cv::Vec3i(Vec3f): saturate_cast<int> = cvRound
            newCenter[i] = (float)pixOffset[i] * voxelSize;
        }
        const Vec3f shift = pose.rotate(newCenter);
        for (int i = 0; i < 3; ++i) pose.t[i] += shift[i];
        const float ext = std::max({p90[0] - p10[0], p90[1] - p10[1], p90[2] - p10[2]});
        const int n = ((int)std::ceil(volPad * ext / voxelSize) + 1) / 2 * 2;
        const Vec3i newRes{n, n, n};
        for (int i = 0; i < 3; ++i) pixOffset[i] -= (int)std::nearbyint((newRes[i] - volumeRes[i]) * 0.5);
        const size_t nv = (size_t)n * n * n;
        DeviceArray<float> nt(nv), nw(nv), nf(2 * nv);
        ok(emf_resize_volume(tsdfVol.data(), tsdfWeights.data(), fgBgProbs.data(), volumeRes.data(), nt.data(), nw.data(), nf.data(),
                             newRes.data(), pixOffset.data(), (emf_stream_t)stream), "emf_resize_volume");
        cu(cudaStreamSynchronize(stream), "sync");     // the old arrays are released below
        tsdfVol = std::move(nt); tsdfWeights = std::move(nw); fgBgProbs = std::move(nf);
        fgProbs.allocate(nv);
        tsdfGrads = DeviceArray<float>();
        volumeRes = newRes;
        computeFgProbs(stream);
        return newCenter;
    }
    std::vector<float> getFgProbVol() const { return fgProbs.download(); }

private:
    void resetBox() {
        const int32_t empty[6] = {1, 1, 1, 0, 0, 0};
        fgBox.upload(empty, 6);
    }
    DeviceArray<float> fgBgProbs, fgProbs;
    DeviceArray<int32_t> fgBox;
    int exCount = 1, nonExCount = 0;
};

}  // namespace emfb
#endif  // EMF_B200_HPP
