// cvfull.hpp -- type-level stand-in for the parts of OpenCV (core, cuda, viz) and Sophus that the reference's CLASS headers
// (include/EMFusion/core/{data,TSDF,ObjTSDF}.h) and include/emf_b200_cv.hpp touch.  TEST INFRASTRUCTURE for the CPU-only
// compile test tests/test_cv_adapter_compiles.py: OpenCV-with-CUDA, Eigen and Sophus are not installable in this image.
// Types carry just enough behaviour (pose algebra, headers of matrices) for the adapter to compile and for its signature
// static_asserts against the reference's own declarations to be meaningful; nothing here is linked into the product.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstring>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) (((depth) & 7) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_Assert(expr) do { if (!(expr)) throw std::string("CV_Assert: " #expr); } while (0)
typedef unsigned char uchar;

namespace cv {
struct Size { int width = 0, height = 0; Size() = default; Size(int w, int h) : width(w), height(h) {} };
struct Vec3f {
    float val[3];
    Vec3f() : val{0, 0, 0} {}
    Vec3f(float x, float y, float z) : val{x, y, z} {}
    float& operator[](int i) { return val[i]; }
    const float& operator[](int i) const { return val[i]; }
};
struct Vec3i {
    int val[3];
    Vec3i() : val{0, 0, 0} {}
    Vec3i(int x, int y, int z) : val{x, y, z} {}
    static Vec3i all(int v) { return Vec3i(v, v, v); }
    int& operator[](int i) { return val[i]; }
    const int& operator[](int i) const { return val[i]; }
};
struct Matx33f {
    float val[9];
    Matx33f() : val{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
    Matx33f(float a, float b, float c, float d, float e, float f, float g, float h, float i) : val{a, b, c, d, e, f, g, h, i} {}
    float operator()(int r, int c) const { return val[3 * r + c]; }
};
struct Affine3f {
    Matx33f R;
    Vec3f t;
    Affine3f() = default;
    Affine3f(const Matx33f& r, const Vec3f& tt) : R(r), t(tt) {}
    Matx33f rotation() const { return R; }
    Vec3f translation() const { return t; }
    Affine3f translate(const Vec3f& d) const { Affine3f a = *this; for (int k = 0; k < 3; ++k) a.t.val[k] += d.val[k]; return a; }
    Affine3f inv() const {
        Affine3f r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.R.val[3 * i + j] = R.val[3 * j + i];
        for (int i = 0; i < 3; ++i) r.t.val[i] = -(r.R.val[3 * i] * t.val[0] + r.R.val[3 * i + 1] * t.val[1] + r.R.val[3 * i + 2] * t.val[2]);
        return r;
    }
    Affine3f operator*(const Affine3f& o) const {
        Affine3f r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
            r.R.val[3 * i + j] = R.val[3 * i] * o.R.val[j] + R.val[3 * i + 1] * o.R.val[3 + j] + R.val[3 * i + 2] * o.R.val[6 + j];
        for (int i = 0; i < 3; ++i) r.t.val[i] = R.val[3 * i] * o.t.val[0] + R.val[3 * i + 1] * o.t.val[1] + R.val[3 * i + 2] * o.t.val[2] + t.val[i];
        return r;
    }
};
class Mat {
public:
    int rows = 0, cols = 0, flags = 0;
    std::vector<unsigned char> store;
    unsigned char* data = nullptr;
    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        rows = r; cols = c; flags = type;
        const size_t es = (size_t)(((type & 7) == CV_8U) ? 1 : 4) * (size_t)((type >> CV_CN_SHIFT) + 1);
        store.assign((size_t)r * c * es, 0); data = store.data();
    }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data) + (size_t)r * cols; }
    bool empty() const { return data == nullptr; }
};
namespace viz { struct Mesh { Mat cloud, normals, polygons; }; }
namespace cuda {
class Stream {
public:
    Stream() : s_(nullptr) {}
    explicit Stream(cudaStream_t s) : s_(s) {}
    static Stream& Null() { static Stream n; return n; }
    void waitForCompletion() { cudaStreamSynchronize(s_); }
    cudaStream_t s_;
};
struct StreamAccessor { static cudaStream_t getStream(const Stream& s) { return s.s_; } };
struct Event {};
class GpuMat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
    int flags = 0;
    GpuMat() = default;
    GpuMat(int r, int c, int type, void* p, size_t st) : rows(r), cols(c), step(st), data((unsigned char*)p), flags(type) {}
    int type() const { return flags; }
    int channels() const { return (flags >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return (size_t)(((flags & 7) == CV_8U) ? 1 : 4) * channels(); }
    bool empty() const { return data == nullptr; }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows == 1; }
};
}  // namespace cuda
}  // namespace cv
namespace Sophus { struct SE3f {}; }
