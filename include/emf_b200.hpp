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
#include <memory>
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
    emf_image external{};    // non-owning view of somebody else's image (ptr != nullptr): c() returns it
    emf_image c() const {
        if (external.ptr) return external;
        return emf_image{(void*)mem.data(), (size_t)width * channels * sizeof(T), width, height};
    }
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
        trackIterations = st.iterations;
        return st.iterations;
    }
    // :333-338
    void syncTrack(Affine& cam_pose) { cam_pose = pose * rel_pose_CO; }

    // getMesh (:356-373; ObjTSDF :247-268 through descriptor()'s fgProbs): marching cubes over the voxels with weight > 0
    // (objects: and fgProb > 0.5); cv::viz::Mesh's three arrays on the host
    struct Mesh { std::vector<float> cloud, normals; std::vector<int32_t> polygons; };
    Mesh getMesh(cudaStream_t stream = nullptr) const {
        const emf_volume v = descriptor();
        const size_t wsb = emf_mesh_workspace_bytes(v.res);
        DeviceArray<unsigned char> ws(wsb);
        ok(emf_mesh_count(&v, ws.data(), wsb, (emf_stream_t)stream), "emf_mesh_count");
        int32_t counts[2] = {0, 0};
        cu(cudaMemcpyAsync(counts, ws.data(), sizeof counts, cudaMemcpyDeviceToHost, stream), "mesh counts");
        cu(cudaStreamSynchronize(stream), "sync");
        if (counts[0] < 0 || counts[1] < 0) throw std::runtime_error("mesh larger than 2^31 - 1 elements");
        Mesh m;
        if (counts[0] == 0) return m;
        DeviceArray<float> dv((size_t)3 * counts[0]), dn((size_t)3 * counts[0]);
        DeviceArray<int32_t> dt((size_t)counts[1]);
        ok(emf_mesh_extract(&v, ws.data(), wsb, dv.data(), dn.data(), dt.data(), (emf_stream_t)stream), "emf_mesh_extract");
        cu(cudaStreamSynchronize(stream), "sync");
        m.cloud = dv.download(); m.normals = dn.download(); m.polygons = dt.download();
        return m;
    }
    virtual std::vector<float> getTSDF() const { return tsdfVol.download(); }
    virtual std::vector<float> getWeightsVol() const { return tsdfWeights.download(); }
    // the volume as the C ABI sees it (for the batched / engine entry points)
    virtual emf_volume descriptor() const { return cVolume(nullptr); }
    void setPose(const Affine& p) { pose = p; }
    const Affine& relPoseCO() const { return rel_pose_CO; }
    void setRelPoseCO(const Affine& T) { rel_pose_CO = T; }
    Image<float>& trackingWeightsImage() { return intWeights; }    // (scratch of the fused tracker iteration: emf_track_iterate)

    TSDFParams params;
    bool trackingConverged = false;
    int trackIterations = 0;          // iterations of the last track() (diagnostic)

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
    emf_volume descriptor() const override {
        emf_volume v = cVolume(fgProbs.data());
        v.fg_box = fgBox.data();
        return v;
    }

private:
    void resetBox() {
        const int32_t empty[6] = {1, 1, 1, 0, 0, 0};
        fgBox.upload(empty, 6);
    }
    DeviceArray<float> fgBgProbs, fgProbs;
    DeviceArray<int32_t> fgBox;
    int exCount = 1, nonExCount = 0;
};

// include/EMFusion/core/data.h:76-199 (the fields the hot path consumes)
struct Params {
    int frameW = 640, frameH = 480;
    Matx33f intr{525.f, 0, 319.5f, 0, 525.f, 239.5f, 0, 0, 1};
    Vec3i globalVolumeDims{512, 512, 512};
    float globalVoxelSize = 0.01f, globalRelTruncDist = 10.0f;
    Vec3i objVolumeDims{64, 64, 64};
    float objRelTruncDist = 10.0f;
    Affine volumePose = Affine::translation(0, 0, 2.56f);
    int visibilityThresh = 40 * 40, boundary = 20, maxTrackingIter = 100;
    TSDFParams tsdfParams;
};

// The hot methods of emf::EMFusion (include/EMFusion/core/EMFusion.h, src/core/EMFusion.cpp) over the native frame engine:
// processFrame = computePoints, computeAssociationWeights, [performTracking], raycast, integrateDepth (:70-129 minus Mask R-CNN),
// each also callable on its own.  `background` and `objects` are the reference's members (:452-454).
class EMFusion {
public:
    explicit EMFusion(const Params& p)
        : params(p),
          background(p.globalVolumeDims, p.globalVoxelSize, p.globalRelTruncDist * p.globalVoxelSize, p.volumePose, p.tsdfParams,
                     p.frameW, p.frameH) {
        emf_engine_config cfg{};
        cfg.width = p.frameW; cfg.height = p.frameH;
        std::memcpy(cfg.K, p.intr.data(), sizeof cfg.K);
        cfg.params = p.tsdfParams.c();
        cfg.boundary = p.boundary; cfg.visibility_thresh = p.visibilityThresh;
        engine_ = emf_engine_create(&cfg);
        if (!engine_) throw std::runtime_error("emf_engine_create failed");
        dirty_ = true;
    }
    EMFusion(const EMFusion&) = delete;
    EMFusion& operator=(const EMFusion&) = delete;
    ~EMFusion() { if (engine_) emf_engine_destroy(engine_); }

    // Anything that re-allocates a volume's arrays behind the engine's back (ObjTSDF::resize as called by updateObj,
    // src/core/EMFusion.cpp:1010-1018) must be followed by invalidate(): the next frame re-submits the volume table.
    void invalidate() { dirty_ = true; }
    // updateObj's resize (:1010-1018) on an object owned by this instance
    template <class... A> void resizeObj(size_t i, A&&... a) { objects.at(i)->resize(std::forward<A>(a)...); dirty_ = true; }

    // createObj (:908-920, the volume part): a new object volume; its association image starts at 1
    ObjTSDF& createObj(const Affine& obj_pose, float voxelSize) {
        objects.emplace_back(new ObjTSDF(params.objVolumeDims, voxelSize, params.objRelTruncDist * voxelSize, obj_pose, params.tsdfParams,
                                         params.frameW, params.frameH));
        dirty_ = true;
        created_.push_back((int)objects.size());   // engine index (background = 0)
        return *objects.back();
    }

    void processFrame(const Image<float>& depth, cudaStream_t stream = nullptr) {
        depth_ = &depth;
        frame(EMF_FRAME_POINTS, stream);
        if (frameCount > 0) {
            computeAssociationWeights(stream);
            if (trackingEnabled) { performTracking(stream); computeAssociationWeights(stream); }
            raycast(stream);
            integrateDepth(stream);
        } else {
            frame(EMF_FRAME_INTEGRATE | EMF_FRAME_INTEGRATE_ALL, stream);
        }
        ++frameCount;
    }
    void computeAssociationWeights(cudaStream_t stream = nullptr) { frame(EMF_FRAME_ASSOC, stream); }          // :635-670
    void raycast(cudaStream_t stream = nullptr) { frame(EMF_FRAME_RAYCAST | EMF_FRAME_COMPOSITE, stream); }    // :726-795
    void integrateDepth(cudaStream_t stream = nullptr) { frame(EMF_FRAME_INTEGRATE, stream); }                 // :865-889
    // :672-722 -- the background first, then all objects together, each as one device-resident Levenberg-Marquardt run
    void performTracking(cudaStream_t stream = nullptr) {
        sync(stream);
        const Image<float> pts = view(EMF_IMG_POINTS, 0, 3);
        {
            const Image<float> a = view(EMF_IMG_VOL_ASSOC, 0, 1);
            background.prepareTracking(pose);
            background.track(pts, a, params.intr, params.maxTrackingIter, stream);
            background.syncTrack(pose);
        }
        computeAssociationWeights(stream);
        for (size_t i = 0; i < objects.size(); ++i) {
            const Image<float> a = view(EMF_IMG_VOL_ASSOC, (int)i + 1, 1);
            objects[i]->prepareTracking(pose);
            objects[i]->track(pts, a, params.intr, params.maxTrackingIter, stream);
            objects[i]->syncTrack(pose);
        }
    }
    // engine-owned images (valid until the object list changes): EMF_IMG_* of include/emf_b200.h
    std::vector<float> downloadF(int what, int index, int channels) { sync(nullptr); return view(what, index, channels).mem_download(); }
    std::vector<uint8_t> downloadSegmentation() {
        sync(nullptr);
        emf_image im;
        ok(emf_engine_image(engine_, EMF_IMG_SEG, 0, &im), "emf_engine_image");
        std::vector<uint8_t> h((size_t)im.width * im.height);
        cu(cudaMemcpy2D(h.data(), im.width, im.ptr, im.pitch, im.width, im.height, cudaMemcpyDeviceToHost), "download seg");
        return h;
    }
    std::vector<int32_t> visibilityCounts() {
        std::vector<int32_t> c(objects.size());
        if (!c.empty()) ok(emf_engine_vis_counts(engine_, c.data(), (int)c.size()), "emf_engine_vis_counts");
        return c;
    }

    Params params;
    TSDF background;
    std::vector<std::unique_ptr<ObjTSDF>> objects;
    Affine pose;                 // camera pose (the caller sets it when tracking is bypassed)
    int frameCount = 0;
    bool trackingEnabled = false;

private:
    // a non-owning view of an engine image with the interface the volume classes take
    struct ViewF : Image<float> {
        emf_image im{};
        std::vector<float> mem_download() const {
            std::vector<float> h((size_t)im.width * im.height * channels);
            cu(cudaMemcpy2D(h.data(), (size_t)im.width * channels * 4, im.ptr, im.pitch, (size_t)im.width * channels * 4, im.height,
                            cudaMemcpyDeviceToHost), "download");
            return h;
        }
    };
    ViewF view(int what, int index, int channels) {
        ViewF v;
        ok(emf_engine_image(engine_, what, index, &v.im), "emf_engine_image");
        v.width = v.im.width; v.height = v.im.height; v.channels = channels;
        v.external = v.im;
        return v;
    }
    void sync(cudaStream_t s) { cu(cudaStreamSynchronize(s), "sync"); }
    void syncVolumes(cudaStream_t stream) {
        std::vector<emf_volume> v{background.descriptor()};
        for (auto& o : objects) v.push_back(o->descriptor());
        ok(emf_engine_set_volumes(engine_, (int)v.size(), v.data(), 1, (emf_stream_t)stream), "emf_engine_set_volumes");
        for (int idx : created_)
            if (frameCount > 0) ok(emf_engine_force_integrate(engine_, idx), "emf_engine_force_integrate");   // :918
        created_.clear();
        dirty_ = false;
    }
    void frame(unsigned flags, cudaStream_t stream) {
        if (dirty_) syncVolumes(stream);
        if (!depth_) throw std::runtime_error("no depth frame");
        std::vector<emf_pose> co{(background.getPose().inv() * pose).c()}, oc{(pose.inv() * background.getPose()).c()};
        for (auto& o : objects) { co.push_back((o->getPose().inv() * pose).c()); oc.push_back((pose.inv() * o->getPose()).c()); }
        const emf_image d = depth_->c();
        ok(emf_engine_frame(engine_, &d, co.data(), oc.data(), flags, (emf_stream_t)stream), "emf_engine_frame");
    }
    emf_engine* engine_ = nullptr;
    const Image<float>* depth_ = nullptr;
    std::vector<int> created_;
    bool dirty_ = true;
};

}  // namespace emfb
#endif  // EMF_B200_HPP
