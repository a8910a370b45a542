// Type-only stand-in for the handful of OpenCV / OpenCV-CUDA types that
// /root/reference/src/core/cuda/{TSDF,ObjTSDF}.cu and their headers touch.
// TEST INFRASTRUCTURE ONLY (oracle/): it lets the reference kernels be
// compiled *in place and unchanged* (see oracle/Makefile) so that the product
// kernels can be diffed against them on a B200.  There is no arithmetic in
// here apart from a byte-fill (GpuMat::setTo with an all-zero scalar, which
// the reference launchers call before two of their kernels) and a trivial
// host-side sum used only by the (off-path) marching-cubes launcher.
// Not part of the product; nothing under emfusion_b200/ includes this file.
#pragma once
#include <cuda_runtime.h>
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <memory>
#include <vector>

#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAKE_TYPE CV_MAKETYPE
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_Assert(expr) do { if (!(expr)) { std::fprintf(stderr, "CV_Assert failed: %s\n", #expr); std::abort(); } } while (0)

typedef unsigned char uchar;
typedef unsigned short ushort;

namespace cv {

struct Scalar {
    double val[4];
    Scalar() : val{0, 0, 0, 0} {}
    Scalar(double a, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double operator[](int i) const { return val[i]; }
};

struct Matx33f { float val[9]; };
struct Vec3f { float val[3]; };
struct Vec3i { int val[3]; };

template <typename T> struct DataType;
template <> struct DataType<unsigned char> { enum { depth = CV_8U }; };
template <> struct DataType<bool> { enum { depth = CV_8U }; };
template <> struct DataType<int> { enum { depth = CV_32S }; };
template <> struct DataType<float> { enum { depth = CV_32F }; };

namespace cuda {

class Stream {
public:
    Stream() : s_(nullptr) {}
    explicit Stream(cudaStream_t s) : s_(s) {}
    static Stream& Null() { static Stream n; return n; }
    cudaStream_t s_;
};

struct StreamAccessor {
    static cudaStream_t getStream(const Stream& s) { return s.s_; }
};

template <typename T> struct PtrStep {
    T* data;
    size_t step;  // bytes per row
    __host__ __device__ PtrStep() : data(nullptr), step(0) {}
    __host__ __device__ PtrStep(T* d, size_t s) : data(d), step(s) {}
    __host__ __device__ T* ptr(int y = 0) { return (T*)((char*)data + y * step); }
    __host__ __device__ const T* ptr(int y = 0) const { return (const T*)((const char*)data + y * step); }
    __host__ __device__ T& operator()(int y, int x) { return ptr(y)[x]; }
    __host__ __device__ const T& operator()(int y, int x) const { return ptr(y)[x]; }
};

template <typename T> struct PtrStepSz : public PtrStep<T> {
    int cols, rows;
    __host__ __device__ PtrStepSz() : cols(0), rows(0) {}
    __host__ __device__ PtrStepSz(int r, int c, T* d, size_t s) : PtrStep<T>(d, s), cols(c), rows(r) {}
};

// Non-owning or owning (shared) 2-D device matrix header.
class GpuMat {
public:
    int rows, cols;
    size_t step;
    unsigned char* data;
    int flags;  // CV type
    std::shared_ptr<void> owner;

    GpuMat() : rows(0), cols(0), step(0), data(nullptr), flags(0) {}
    GpuMat(int r, int c, int type, void* d, size_t s)
        : rows(r), cols(c), step(s), data((unsigned char*)d), flags(type) {}

    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize() const {
        static const int sz[] = {1, 1, 2, 2, 4, 4, 8};
        return (size_t)sz[depth()] * channels();
    }
    bool empty() const { return data == nullptr; }

    template <typename T> T* ptr(int y = 0) { return (T*)(data + y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + y * step); }

    template <typename T> operator PtrStep<T>() const { return PtrStep<T>((T*)data, step); }
    template <typename T> operator PtrStepSz<T>() const { return PtrStepSz<T>(rows, cols, (T*)data, step); }

    // reshape of a continuous matrix (the only kind this shim creates)
    GpuMat reshape(int cn, int new_rows = 0) const {
        GpuMat m = *this;
        size_t total_scalars = (size_t)rows * cols * channels();
        if (cn == 0) cn = channels();
        if (new_rows == 0) new_rows = rows;
        m.rows = new_rows;
        m.cols = (int)(total_scalars / ((size_t)new_rows * cn));
        m.flags = CV_MAKETYPE(depth(), cn);
        m.step = m.cols * m.elemSize();
        return m;
    }

    // Real fill; only all-equal-byte patterns (zero) are needed by the reference launchers.
    GpuMat& setTo(const Scalar& v, Stream& stream = Stream::Null()) {
        bool zero = true;
        for (int i = 0; i < channels(); ++i) zero = zero && (v.val[i] == 0.0);
        CV_Assert(zero && "shim GpuMat::setTo supports only zero fill");
        if (step == cols * elemSize())
            cudaMemsetAsync(data, 0, step * rows, stream.s_);
        else
            cudaMemset2DAsync(data, step, 0, cols * elemSize(), rows, stream.s_);
        return *this;
    }
    void create(int r, int c, int type) {
        static const int sz[] = {1, 1, 2, 2, 4, 4, 8};
        size_t es = (size_t)sz[CV_MAT_DEPTH(type)] * CV_MAT_CN(type);
        void* p = nullptr;
        cudaMalloc(&p, (size_t)r * c * es + 16);
        owner = std::shared_ptr<void>(p, [](void* q) { cudaFree(q); });
        rows = r; cols = c; flags = type; step = c * es; data = (unsigned char*)p;
    }
};

inline void createContinuous(int rows, int cols, int type, GpuMat& m) { m.create(rows, cols, type); }
inline GpuMat createContinuous(int rows, int cols, int type) { GpuMat m; m.create(rows, cols, type); return m; }

// Only used by the off-path marching-cubes launcher; host-side int/float sum.
inline Scalar sum(const GpuMat& m) {
    std::vector<unsigned char> h(m.step * m.rows);
    cudaMemcpy(h.data(), m.data, h.size(), cudaMemcpyDeviceToHost);
    double acc = 0;
    for (int y = 0; y < m.rows; ++y)
        for (int x = 0; x < m.cols * m.channels(); ++x) {
            const unsigned char* row = h.data() + y * m.step;
            switch (m.depth()) {
            case CV_8U: acc += row[x]; break;
            case CV_32S: acc += ((const int*)row)[x]; break;
            case CV_32F: acc += ((const float*)row)[x]; break;
            default: break;
            }
        }
    return Scalar(acc);
}

}  // namespace cuda
}  // namespace cv
