// xchg.cu -- the multi-GPU exchanges of a frame over NVLink peer memory, without a communication library on the path.
//
// The two exchanges of the sharded frame are (1) the per-pixel association normaliser
// (reference src/core/EMFusion.cpp:653-657: a sum over ALL volumes, i.e. over all ranks) and (2) the composite of the
// per-rank raycasts on the rank that owns the background (:760-794), followed by the visibility counters every rank's
// integrate is gated with (:869-872).  With NCCL these are an all-reduce, a gather and a broadcast: three library kernels with
// their own launch and rendezvous latencies, and a gather that moves every rank's whole pre-composite (8.9 MB) although the
// merge reads 5 bytes per pixel and part plus the winner's 24.
//
// Here every rank owns one exchange buffer (cudaMalloc, exported with cudaIpcGetMemHandle and opened by every peer, so a
// peer's buffer is an ordinary device pointer that loads and stores travel to over NVLink / NVSwitch):
//   * a producer writes its data into its OWN buffer with its normal kernels, then raises a flag in the consumers' buffers
//     (k_signal: __threadfence_system + one 32-bit store per consumer; flags carry the frame number, so nothing is reset);
//   * a consumer's stream waits for its local flags (k_wait: one thread per producer spinning on local memory, bounded by a
//     time-out that sets an error word instead of hanging the GPU), then its consuming kernel reads the producers' buffers
//     directly: k_sum_parts adds the partial normalisers in rank order (the same order, hence the same bits, on every
//     rank); emf_composite_merge (raycast.cu) is simply handed peer pointers; k_scatter_u32 stores the visibility counters
//     into every rank's buffer.
// Everything stays on the frame's stream; the host never synchronises.
#include "common.cuh"
#include <string.h>

namespace emfb {

struct SignalArgs { uint32_t* flag[16]; int n; };
__global__ void k_signal_args(const __grid_constant__ SignalArgs A, uint32_t value) {
    __threadfence_system();
    const int i = threadIdx.x;
    if (i < A.n && A.flag[i]) {
        volatile uint32_t* f = A.flag[i];
        *f = value;
        __threadfence_system();
    }
}

// flags[i] >= value for every i (frame numbers only grow); gives up after timeout_ns and records it in *err
__global__ void k_wait(const uint32_t* flags, int n, uint32_t value, uint32_t* err, unsigned long long timeout_ns) {
    const int i = threadIdx.x;
    if (i < n) {
        const volatile uint32_t* f = flags + i;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int32_t)(*f - value) < 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) { atomicExch(err, 1u + (uint32_t)i); break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

struct SumArgs { const float* part[16]; int n; };
__global__ void __launch_bounds__(256) k_sum_parts(const __grid_constant__ SumArgs A, Img<float> out) {
    const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= out.w || y >= out.h) return;
    const size_t i = (size_t)y * out.w + x;
    if (x + 3 < out.w && ((out.w & 3) == 0)) {
        float4 s = *reinterpret_cast<const float4*>(A.part[0] + i);
        for (int r = 1; r < A.n; ++r) {
            const float4 p = *reinterpret_cast<const float4*>(A.part[r] + i);
            s.x = fadd(s.x, p.x); s.y = fadd(s.y, p.y); s.z = fadd(s.z, p.z); s.w = fadd(s.w, p.w);
        }
        float* o = out.row(y) + x;
        o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.w;
    } else {
        for (int k = 0; k < 4 && x + k < out.w; ++k) {
            float s = A.part[0][i + k];
            for (int r = 1; r < A.n; ++r) s = fadd(s, A.part[r][i + k]);
            out.at(y, x + k) = s;
        }
    }
}

struct ScatterArgs { uint32_t* dst[16]; int n; };
__global__ void k_scatter_u32(const uint32_t* __restrict__ src, int count, const __grid_constant__ ScatterArgs A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t v = src[i];
    for (int r = 0; r < A.n; ++r) A.dst[r][i] = v;
}

}  // namespace emfb

using namespace emfb;

extern "C" EMF_API int emf_xchg_alloc(size_t bytes, void** ptr_out, unsigned char handle_out[64]) {
    if (!bytes || !ptr_out || !handle_out) return EMF_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return EMF_ERR_CUDA;
    if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaFree(p); return EMF_ERR_CUDA; }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); cudaGetLastError(); return EMF_ERR_CUDA; }
    memcpy(handle_out, &h, 64);
    *ptr_out = p;
    return EMF_OK;
}

extern "C" EMF_API int emf_xchg_open(const unsigned char handle[64], void** ptr_out) {
    if (!handle || !ptr_out) return EMF_ERR_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return EMF_ERR_CUDA; }
    *ptr_out = p;
    return EMF_OK;
}

extern "C" EMF_API int emf_xchg_close(void* peer_ptr) {
    if (!peer_ptr) return EMF_ERR_INVALID;
    return cudaIpcCloseMemHandle(peer_ptr) == cudaSuccess ? EMF_OK : EMF_ERR_CUDA;
}

extern "C" EMF_API int emf_xchg_free(void* ptr) {
    if (!ptr) return EMF_ERR_INVALID;
    return cudaFree(ptr) == cudaSuccess ? EMF_OK : EMF_ERR_CUDA;
}

extern "C" EMF_API int emf_xchg_signal(int n, uint32_t* const* flags, uint32_t value, emf_stream_t stream) {
    if (n <= 0 || n > 16 || !flags) return EMF_ERR_INVALID;
    SignalArgs A;
    A.n = n;
    for (int i = 0; i < n; ++i) A.flag[i] = flags[i];
    k_signal_args<<<1, 32, 0, (cudaStream_t)stream>>>(A, value);
    return launch_status();
}

extern "C" EMF_API int emf_xchg_wait(const uint32_t* flags, int n, uint32_t value, uint32_t* err, double timeout_s,
                                     emf_stream_t stream) {
    if (n <= 0 || n > 32 || !flags || !err || !(timeout_s > 0.0)) return EMF_ERR_INVALID;
    k_wait<<<1, 32, 0, (cudaStream_t)stream>>>(flags, n, value, err, (unsigned long long)(timeout_s * 1e9));
    return launch_status();
}

extern "C" EMF_API int emf_xchg_sum_images(int n_parts, const float* const* parts, const emf_image* out, emf_stream_t stream) {
    if (n_parts <= 0 || n_parts > 16 || !parts || !image_ok(out, 4)) return EMF_ERR_INVALID;
    SumArgs A;
    A.n = n_parts;
    for (int i = 0; i < n_parts; ++i) { if (!parts[i] || !aligned16(parts[i])) return EMF_ERR_INVALID; A.part[i] = parts[i]; }
    const dim3 grid((out->width + 127) / 128, (out->height + 7) / 8);
    k_sum_parts<<<grid, 256, 0, (cudaStream_t)stream>>>(A, view<float>(out));
    return launch_status();
}

extern "C" EMF_API int emf_xchg_scatter_u32(const uint32_t* src, int count, int n_dst, uint32_t* const* dst, emf_stream_t stream) {
    if (!src || count <= 0 || n_dst <= 0 || n_dst > 16 || !dst) return EMF_ERR_INVALID;
    ScatterArgs A;
    A.n = n_dst;
    for (int i = 0; i < n_dst; ++i) { if (!dst[i]) return EMF_ERR_INVALID; A.dst[i] = dst[i]; }
    k_scatter_u32<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(src, count, A);
    return launch_status();
}
