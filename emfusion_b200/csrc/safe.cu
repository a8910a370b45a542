// safe.cu -- "safe sample" bitmaps of a TSDF volume, derived from the constant-segment bitmaps that the
// integrate kernel maintains (integrate.cu).
//
// No counterpart in the reference: this is acceleration state for the raycast (raycast.cu).  The
// reference's march (src/core/cuda/TSDF.cu:523-572) samples the volume at every step; wherever all eight
// corners of a sample hold the same value c in {+1, 0, -1} the trilinear result is exactly c (checked
// exhaustively over every fp32 fraction), so a sample taken while the previous sample was already c
// changes nothing of the march state.
//
// const_bits[m] has one bit per 4-voxel x-segment: "all four voxels equal c_m".  safe_bits[m] is its
// erosion by the box  x: -1..+1 segments, y: -1..+2 rows, z: -1..+2 slices  (everything outside the
// volume counts as not constant).  If the bit of the segment holding voxel (X, Y, Z) is set, every voxel
// in [X-4, X+7] x [Y-1, Y+2] x [Z-1, Z+2] (at least) equals c_m, hence every trilinear sample whose base
// voxel lies within one voxel of (X, Y, Z) returns exactly c_m.
//
// One CTA = a 32 (y) x 4 (z) tile of one word column of one map: the 35 x 7 input rows are eroded along x
// with shifts while they are staged in shared memory, then along y and z from there.
#include "common.cuh"

namespace emfb {

struct SafeVol {
    const uint32_t* cbits;
    uint32_t* sbits;
    int wpr, ry, rz;
    int segs;            // segments per row (Rx / 4)
    int tiles_y, tiles_z;
    int first_cta;
    size_t map_words;
};
struct SafeParams {
    SafeVol v[EMF_MAX_VOLUMES];
    int n_vol;
};

constexpr int kTy = 32, kTz = 4;

__global__ void __launch_bounds__(kTy * kTz) k_safe_bits(const __grid_constant__ SafeParams P) {
    __shared__ uint32_t s_x[kTz + 3][kTy + 3];    // x-eroded input rows
    __shared__ uint32_t s_y[kTz + 3][kTy];        // ... then y-eroded
    int lo = 0, hi = P.n_vol - 1;
    const int b = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (P.v[mid].first_cta <= b) lo = mid; else hi = mid - 1;
    }
    const SafeVol& V = P.v[lo];
    int r = b - V.first_cta;
    const int tz = r % V.tiles_z; r /= V.tiles_z;
    const int ty = r % V.tiles_y; r /= V.tiles_y;
    const int wx = r % V.wpr;
    const int m = r / V.wpr;
    const uint32_t* __restrict__ src = V.cbits + (size_t)m * V.map_words;
    const int y0 = ty * kTy - 1, z0 = tz * kTz - 1;
    // bits of this word that are real segments
    const int nbits = min(32, V.segs - wx * 32);
    const uint32_t valid = nbits >= 32 ? 0xffffffffu : ((1u << nbits) - 1u);
    for (int i = threadIdx.x; i < (kTz + 3) * (kTy + 3); i += kTy * kTz) {
        const int dz = i / (kTy + 3), dy = i - dz * (kTy + 3);
        const int y = y0 + dy, z = z0 + dz;
        uint32_t e = 0;
        if (y >= 0 && y < V.ry && z >= 0 && z < V.rz) {
            const uint32_t* row = src + ((size_t)z * V.ry + y) * V.wpr;
            const uint32_t w = __ldg(row + wx) & valid;
            const uint32_t prev = wx > 0 ? __ldg(row + wx - 1) : 0u;
            const uint32_t next = wx + 1 < V.wpr ? __ldg(row + wx + 1) : 0u;
            e = w & ((w << 1) | (prev >> 31)) & ((w >> 1) | (next << 31));
        }
        s_x[dz][dy] = e;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (kTz + 3) * kTy; i += kTy * kTz) {
        const int dz = i / kTy, yy = i - dz * kTy;
        s_y[dz][yy] = s_x[dz][yy] & s_x[dz][yy + 1] & s_x[dz][yy + 2] & s_x[dz][yy + 3];
    }
    __syncthreads();
    const int yy = threadIdx.x % kTy, zz = threadIdx.x / kTy;
    const int y = ty * kTy + yy, z = tz * kTz + zz;
    if (y < V.ry && z < V.rz)
        V.sbits[(size_t)m * V.map_words + ((size_t)z * V.ry + y) * V.wpr + wx] =
            s_y[zz][yy] & s_y[zz + 1][yy] & s_y[zz + 2][yy] & s_y[zz + 3][yy];
}

}  // namespace emfb

using namespace emfb;

extern "C" EMF_API int emf_update_safe_bits(int n_vol, const emf_volume* vols, emf_stream_t stream) {
    if (n_vol <= 0 || !vols) return EMF_ERR_INVALID;
    if (n_vol > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    SafeParams P;
    int n = 0;
    int64_t total = 0;
    for (int i = 0; i < n_vol; ++i) {
        const emf_volume& v = vols[i];
        if (!v.const_bits || !v.safe_bits) continue;
        if (!res_ok(v.res) || v.res[0] % 4) return EMF_ERR_INVALID;
        SafeVol& d = P.v[n++];
        d.cbits = v.const_bits; d.sbits = v.safe_bits;
        d.wpr = emf_bitmap_words_per_row(v.res[0]);
        d.ry = v.res[1]; d.rz = v.res[2];
        d.segs = v.res[0] / 4;
        d.tiles_y = (d.ry + kTy - 1) / kTy; d.tiles_z = (d.rz + kTz - 1) / kTz;
        d.map_words = (size_t)d.wpr * d.ry * d.rz;
        d.first_cta = (int)total;
        total += (int64_t)3 * d.wpr * d.tiles_y * d.tiles_z;
        if (total > 0x7fffffff) return EMF_ERR_UNSUPPORTED;
    }
    if (n == 0) return EMF_OK;
    P.n_vol = n;
    k_safe_bits<<<(unsigned)total, kTy * kTz, 0, (cudaStream_t)stream>>>(P);
    return launch_status();
}

extern "C" EMF_API int emf_reset_bitmaps(const emf_volume* vol, emf_stream_t stream) {
    if (!vol || !res_ok(vol->res)) return EMF_ERR_INVALID;
    const size_t words = (size_t)emf_bitmap_words_per_row(vol->res[0]) * vol->res[1] * vol->res[2];
    cudaStream_t s = (cudaStream_t)stream;
    if (vol->const_bits) {
        cudaMemsetAsync(vol->const_bits, 0x00, words * 4, s);
        cudaMemsetAsync(vol->const_bits + words, 0xff, words * 4, s);
        cudaMemsetAsync(vol->const_bits + 2 * words, 0x00, words * 4, s);
    }
    if (vol->safe_bits) cudaMemsetAsync(vol->safe_bits, 0x00, 3 * words * 4, s);
    return launch_status();
}
