// mcubes.cu -- marching cubes over a TSDF volume: emf::TSDF::getMesh / emf::ObjTSDF::getMesh.
//
// Replaces emf::cuda::TSDF::marchingCubes (reference src/core/cuda/TSDF.cu:855-1152) and the mask / scratch passes around it
// (src/core/TSDF.cpp:356-373, src/core/ObjTSDF.cpp:247-268).  The reference keeps three scratch volumes per TSDF (cubeClasses
// u8, vertIdxBuffer i32, triIdxBuffer i32: 9 B/voxel), clears them and the mask volume per call, classifies, sums on the HOST
// (a full download per sum), scans twice with thrust and then emits.  Here:
//  * pass 1 (k_mc_count): one CTA per row of cubes (fixed y, z); the cube class is formed from the eight corner signs, masked
//    by `weight > 0` (and `fgProb > 0.5` for objects) at all eight corners, looked up in the case tables, and the row's vertex
//    and triangle-index counts are written: 8 B per ROW of scratch instead of 9 B per voxel;
//  * k_mc_scan: exclusive scan over the rows (one CTA), totals left on the device for the caller's allocation;
//  * pass 2 (k_mc_emit): the same traversal; a block-wide scan inside the row gives every cube its base index, vertices are
//    interpolated on the crossed edges and triangles emitted.
// Output layout and values as the reference's: one vertex per crossed edge and cube (not shared between cubes), cubes in
// (z, y, x) order, edges in ascending order; position p1 + mu (p2 - p1) with mu = -v1 / (v2 - v1) and the reference's three
// 1e-5 shortcuts; "normals" = the same interpolation of the forward-difference gradients at the two corners, NOT normalised
// (the reference's `float3 /= float` is a no-op, include/EMFusion/core/cuda/common.cuh:170-173); triangles as VTK polygons
// (3, i0, i1, i2).  Case tables: mc_tables.h, derived by scripts/gen_mc_tables.py (same crossed edges, counts and oriented patch
// boundaries as the reference's tables in all 256 cases; the interior diagonals of patches with more than three vertices are ours).
#include "common.cuh"
#include "mc_tables.h"

namespace emfb {

constexpr int kMcThreads = 128;

struct McParams {
    const float* tsdf;
    const float* weights;
    const float* fg_probs;     // nullable
    int rx, ry, rz;
    float voxel;
    int2* row_counts;          // per row of cubes: (vertices, triangle ints); after k_mc_scan: exclusive prefix
    int n_rows;
    int32_t* totals;           // [0] vertices, [1] triangle ints
    float* vertices;           // pass 2
    float* normals;
    int32_t* triangles;
};

__device__ __forceinline__ bool mc_masked_in(const McParams& P, int64_t i) {
    if (!(__ldg(P.weights + i) > 0.0f)) return false;
    return !P.fg_probs || __ldg(P.fg_probs + i) > 0.5f;
}

// class of the cube whose corner 0 is voxel (x, y, z); 0 when the mask excludes one of its corners (TSDF.cu:883-906)
__device__ __forceinline__ int mc_class(const McParams& P, int x, int y, int z, float vals[8]) {
    const int64_t plane = (int64_t)P.rx * P.ry;
    const int64_t i0 = ((int64_t)z * P.ry + y) * P.rx + x;
    // corners in the reference's order: (x,y,z) (x+1,y,z) (x+1,y,z+1) (x,y,z+1) (x,y+1,z) (x+1,y+1,z) (x+1,y+1,z+1) (x,y+1,z+1)
    const int64_t idx[8] = {i0, i0 + 1, i0 + plane + 1, i0 + plane, i0 + P.rx, i0 + P.rx + 1, i0 + plane + P.rx + 1, i0 + plane + P.rx};
    int cls = 0;
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        ok = ok && mc_masked_in(P, idx[c]);
        vals[c] = __ldg(P.tsdf + idx[c]);
        cls |= (vals[c] < 0.0f) << c;
    }
    return ok ? cls : 0;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    int base = 0;
    total = 0;
#pragma unroll
    for (int k = 0; k < kMcThreads / 32; ++k) { if (k < wid) base += s_warp[k]; total += s_warp[k]; }
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(kMcThreads) k_mc_count(const __grid_constant__ McParams P) {
    __shared__ int s_v[kMcThreads / 32], s_t[kMcThreads / 32];
    const int row = blockIdx.x;
    const int z = row / (P.ry - 1), y = row - z * (P.ry - 1);
    int nv = 0, nt = 0;
    for (int x = threadIdx.x; x < P.rx - 1; x += kMcThreads) {
        float vals[8];
        const int cls = mc_class(P, x, y, z, vals);
        nv += __popc((unsigned)kMcEdges[cls]);
        nt += 4 * kMcNumTris[cls];
    }
    int tv, tt;
    block_exclusive_scan(nv, s_v, tv);
    block_exclusive_scan(nt, s_t, tt);
    if (threadIdx.x == 0) P.row_counts[row] = make_int2(tv, tt);
}

// exclusive scan over the rows, one CTA of 1024 threads (a 512^3 volume has 261 121 rows: 256 per thread)
__global__ void __launch_bounds__(1024) k_mc_scan(const __grid_constant__ McParams P) {
    __shared__ long long s_a[1024], s_b[1024];
    const int t = threadIdx.x;
    const int per = (P.n_rows + 1023) / 1024;
    const int r0 = min(t * per, P.n_rows), r1 = min(r0 + per, P.n_rows);
    long long a = 0, b = 0;
    for (int r = r0; r < r1; ++r) { const int2 c = P.row_counts[r]; a += c.x; b += c.y; }
    s_a[t] = a; s_b[t] = b;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        long long ua = 0, ub = 0;
        if (t >= o) { ua = s_a[t - o]; ub = s_b[t - o]; }
        __syncthreads();
        s_a[t] += ua; s_b[t] += ub;
        __syncthreads();
    }
    long long ba = s_a[t] - a, bb = s_b[t] - b;
    for (int r = r0; r < r1; ++r) {
        const int2 c = P.row_counts[r];
        P.row_counts[r] = make_int2((int)ba, (int)bb);
        ba += c.x; bb += c.y;
    }
    if (t == 1023) {       // (totals beyond 2^31 - 1 cannot be indexed by the int32 triangle list: reported as -1)
        P.totals[0] = s_a[1023] > 0x7fffffffLL ? -1 : (int)s_a[1023];
        P.totals[1] = s_b[1023] > 0x7fffffffLL ? -1 : (int)s_b[1023];
    }
}

// forward-difference gradient at a voxel, zero on the last plane of each axis (TSDF.cu:436-447)
__device__ __forceinline__ float3 mc_grad(const McParams& P, int x, int y, int z) {
    if (x >= P.rx - 1 || y >= P.ry - 1 || z >= P.rz - 1) return make_float3(0.f, 0.f, 0.f);
    const float* p = P.tsdf + ((int64_t)z * P.ry + y) * P.rx + x;
    const float f = __ldg(p);
    return make_float3(fsub(__ldg(p + 1), f), fsub(__ldg(p + P.rx), f), fsub(__ldg(p + (int64_t)P.rx * P.ry), f));
}

// vertexInterp (TSDF.cu:909-919): the comparisons are made in double, as `fabs(float) < 0.00001` is
__device__ __forceinline__ float3 mc_interp(float3 p1, float3 p2, float v1, float v2) {
    if ((double)fabsf(v1) < 0.00001) return p1;
    if ((double)fabsf(v2) < 0.00001) return p2;
    if ((double)fabsf(fsub(v1, v2)) < 0.00001) return p1;
    const float mu = fdiv(-v1, fsub(v2, v1));
    return make_float3(ffma(mu, fsub(p2.x, p1.x), p1.x), ffma(mu, fsub(p2.y, p1.y), p1.y), ffma(mu, fsub(p2.z, p1.z), p1.z));
}

__global__ void __launch_bounds__(kMcThreads) k_mc_emit(const __grid_constant__ McParams P) {
    __shared__ int s_v[kMcThreads / 32], s_t[kMcThreads / 32];
    const int row = blockIdx.x;
    const int z = row / (P.ry - 1), y = row - z * (P.ry - 1);
    int vbase = P.row_counts[row].x, tbase = P.row_counts[row].y;
    const float hx = fmul((float)(P.rx - 1), 0.5f), hy = fmul((float)(P.ry - 1), 0.5f), hz = fmul((float)(P.rz - 1), 0.5f);
    const float s = P.voxel;
    for (int xb = 0; xb < P.rx - 1; xb += kMcThreads) {      // (block-uniform trip count: the scans below are collective)
        const int x = xb + threadIdx.x;
        float vals[8];
        int cls = 0;
        if (x < P.rx - 1) cls = mc_class(P, x, y, z, vals);
        const unsigned edges = kMcEdges[cls];
        const int nt = kMcNumTris[cls];
        int tv, tt;
        const int vo = vbase + block_exclusive_scan(__popc(edges), s_v, tv);
        const int to = tbase + block_exclusive_scan(4 * nt, s_t, tt);
        vbase += tv; tbase += tt;
        if (!edges) continue;
        // corner positions ((i - (R-1)/2.f) * voxelSize, TSDF.cu:945-969) and corner gradients
        const float px0 = fmul(fsub((float)x, hx), s), px1 = fmul(fsub((float)(x + 1), hx), s);
        const float py0 = fmul(fsub((float)y, hy), s), py1 = fmul(fsub((float)(y + 1), hy), s);
        const float pz0 = fmul(fsub((float)z, hz), s), pz1 = fmul(fsub((float)(z + 1), hz), s);
        const float3 ps[8] = {make_float3(px0, py0, pz0), make_float3(px1, py0, pz0), make_float3(px1, py0, pz1), make_float3(px0, py0, pz1),
                              make_float3(px0, py1, pz0), make_float3(px1, py1, pz0), make_float3(px1, py1, pz1), make_float3(px0, py1, pz1)};
        const int cx[8] = {0, 1, 1, 0, 0, 1, 1, 0}, cy[8] = {0, 0, 0, 0, 1, 1, 1, 1}, cz[8] = {0, 0, 1, 1, 0, 0, 1, 1};
        const int ea[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3}, eb[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
        int slot[12];
        int k = 0;
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            slot[e] = k;
            if ((edges >> e) & 1u) {
                const int a = ea[e], b = eb[e];
                const float3 v = mc_interp(ps[a], ps[b], vals[a], vals[b]);
                const float3 n = mc_interp(mc_grad(P, x + cx[a], y + cy[a], z + cz[a]), mc_grad(P, x + cx[b], y + cy[b], z + cz[b]), vals[a], vals[b]);
                float* vp = P.vertices + 3 * (size_t)(vo + k);
                float* np = P.normals + 3 * (size_t)(vo + k);
                vp[0] = v.x; vp[1] = v.y; vp[2] = v.z;
                np[0] = n.x; np[1] = n.y; np[2] = n.z;
                ++k;
            }
        }
        const unsigned long long tris = kMcTris[cls];
        for (int t = 0; t < nt; ++t) {
            int32_t* tp = P.triangles + (size_t)to + 4 * t;
            tp[0] = 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int e = (int)((tris >> (4 * (3 * t + c))) & 15ull);
                int sl = 0;
#pragma unroll
                for (int q = 0; q < 12; ++q) if (q == e) sl = slot[q];
                tp[1 + c] = vo + sl;
            }
        }
    }
}

static int fill_mc(McParams& P, const emf_volume* vol, void* workspace, size_t workspace_bytes) {
    if (!vol || !vol->tsdf || !vol->weights || !res_ok(vol->res) || !workspace) return EMF_ERR_INVALID;
    const int64_t rows = (int64_t)(vol->res[1] - 1) * (vol->res[2] - 1);
    if (rows <= 0 || rows > 0x7fffffff) return EMF_ERR_UNSUPPORTED;
    if (workspace_bytes < emf_mesh_workspace_bytes(vol->res) || ((uintptr_t)workspace & 15) != 0) return EMF_ERR_INVALID;
    P.tsdf = vol->tsdf; P.weights = vol->weights; P.fg_probs = vol->fg_probs;
    P.rx = vol->res[0]; P.ry = vol->res[1]; P.rz = vol->res[2];
    P.voxel = vol->voxel_size;
    P.totals = (int32_t*)workspace;
    P.row_counts = (int2*)((char*)workspace + 256);
    P.n_rows = (int)rows;
    P.vertices = nullptr; P.normals = nullptr; P.triangles = nullptr;
    return EMF_OK;
}

}  // namespace emfb

using namespace emfb;

extern "C" EMF_API size_t emf_mesh_workspace_bytes(const int res[3]) {
    if (!res_ok(res)) return 0;
    return 256 + (size_t)(res[1] - 1) * (res[2] - 1) * sizeof(int2);
}

extern "C" EMF_API int emf_mesh_count(const emf_volume* vol, void* workspace, size_t workspace_bytes, emf_stream_t stream) {
    McParams P;
    const int rc = fill_mc(P, vol, workspace, workspace_bytes);
    if (rc != EMF_OK) return rc;
    k_mc_count<<<(unsigned)P.n_rows, kMcThreads, 0, (cudaStream_t)stream>>>(P);
    k_mc_scan<<<1, 1024, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}

extern "C" EMF_API int emf_mesh_extract(const emf_volume* vol, void* workspace, size_t workspace_bytes, float* vertices, float* normals,
                                int32_t* triangles, emf_stream_t stream) {
    McParams P;
    const int rc = fill_mc(P, vol, workspace, workspace_bytes);
    if (rc != EMF_OK) return rc;
    if (!vertices || !normals || !triangles) return EMF_ERR_INVALID;
    P.vertices = vertices; P.normals = normals; P.triangles = triangles;
    k_mc_emit<<<(unsigned)P.n_rows, kMcThreads, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}
