// bricks.cu -- the brick map of a TSDF volume: acceleration state of the raycast (no reference counterpart).
//
// The reference's march (src/core/cuda/TSDF.cu:523-572) samples the volume at every step, and its step size
// never grows back once a sample with |tsdf| < 1 was seen -- in particular every ray that starts in the
// never-observed apex of the frustum (tsdf == 0) crosses the whole volume in half-voxel steps (~290 samples
// per background ray on the 512^3 bench scene).  Wherever all eight corners of a sample hold the same value
// c in {+1, 0, -1} the trilinear result is exactly c (checked over every fp32 fraction), and a sample that
// returns the value the ray already carries changes nothing of the march state.  The brick map certifies such
// regions so the raycast can skip those samples (raycast.cu) while reproducing the ray parameter bit for bit
// (seq_add.h).
//
// Input: emf_volume::const_bits -- three bitmaps (all +1 / all 0 / all -1), one bit per 4-voxel x-segment,
// maintained by the integrate kernel (integrate.cu).
// Output: emf_volume::brick_map -- one byte per 8^3 brick: (m << 4) | (P << 3) | D, m = 1..3 the constant the whole
// brick holds (0 = mixed), D = 1..7 the Chebyshev radius in bricks of the cube of bricks around it that all hold the
// same constant (D = 1: only the brick itself), P = 1 if the 2 x 2 x 2 block of bricks (b .. b+1 per axis) does.
// Everything outside the volume counts as "constant" (samples there are skipped by the reference as well).
//
//   k_brick_codes: AND of the 2 x 8 x 8 segment bits of every brick  -> three bit planes (32 bricks per word)
//   k_brick_dist : D by repeated 3x3x3 erosion of the planes; one CTA = a 32 x 8 x 8 tile of bricks plus a
//                  6-brick halo, rows held as 64-bit windows, x-erosion by shifts, y/z through shared memory.
#include "common.cuh"

namespace emfb {

constexpr int kBrick = 8;        // voxels per brick edge
constexpr int kMaxD = 7;
constexpr int kHalo = kMaxD - 1;
constexpr int kDTy = 8, kDTz = 8;                 // output tile (bricks) in y, z; 32 in x
constexpr int kDRows = kDTy + 2 * kHalo;          // 20

struct BrickVol {
    const uint32_t* cbits;   // 3 maps
    uint32_t* planes;        // 3 bit planes: [m][bz][by][wpb]
    uint8_t* map;            // [bz][by][bx]
    int wpr;                 // words per voxel row in a segment bitmap
    size_t map_words;        // words per segment bitmap
    int ry, rz, nseg;        // voxel rows, segments per row
    int nbx, nby, nbz, wpb;  // bricks, plane words per brick row
    int first_a, first_b;    // first thread-group of k_brick_codes / first CTA of k_brick_dist
    int tiles_y, tiles_z;
};
struct BrickParams {
    BrickVol v[EMF_MAX_VOLUMES];
    int n_vol;
};

__device__ __forceinline__ const BrickVol& find_vol(const BrickParams& P, int item, bool second) {
    int lo = 0, hi = P.n_vol - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((second ? P.v[mid].first_b : P.v[mid].first_a) <= item) lo = mid; else hi = mid - 1;
    }
    return P.v[lo];
}

// even bits of a 64-bit word -> 32-bit word
__device__ __forceinline__ uint32_t compress_even(uint64_t x) {
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
    x = (x | (x >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)x;
}

// one thread = one plane word: (map m, brick row (by, bz), word wx)
__global__ void __launch_bounds__(128) k_brick_codes(const __grid_constant__ BrickParams P) {
    const int item = blockIdx.x * 128 + threadIdx.x;
    const BrickVol& V = find_vol(P, item, false);
    int r = item - V.first_a;
    const int per_map = V.nbz * V.nby * V.wpb;
    if (r >= 3 * per_map) return;
    const int m = r / per_map; r -= m * per_map;
    const int wx = r % V.wpb; r /= V.wpb;
    const int by = r % V.nby, bz = r / V.nby;
    const uint32_t* __restrict__ src = V.cbits + (size_t)m * V.map_words;
    uint32_t a0 = 0xffffffffu, a1 = 0xffffffffu;
    const int w0 = 2 * wx, w1 = 2 * wx + 1;
    const int y1 = min(V.ry, (by + 1) * kBrick), z1 = min(V.rz, (bz + 1) * kBrick);
    for (int z = bz * kBrick; z < z1; ++z)
        for (int y = by * kBrick; y < y1; ++y) {
            const uint32_t* row = src + ((size_t)z * V.ry + y) * V.wpr;
            a0 &= __ldg(row + w0);
            if (w1 < V.wpr) a1 &= __ldg(row + w1);
        }
    uint64_t a = ((uint64_t)a1 << 32) | a0;
    // segments that do not exist (row padding) count as constant
    const int first_seg = 64 * wx;
    const int nreal = V.nseg - first_seg;            // real segments in this 64-bit window
    if (nreal < 64) a |= nreal <= 0 ? ~0ull : (~0ull << nreal);
    V.planes[(size_t)m * per_map + ((size_t)bz * V.nby + by) * V.wpb + wx] = compress_even(a & (a >> 1));
}

// one CTA = 32 x 8 x 8 bricks (+ halo); thread = one brick row of the (8 + 12)^2 window
__global__ void __launch_bounds__(kDRows * kDRows) k_brick_dist(const __grid_constant__ BrickParams P) {
    __shared__ uint64_t s_e[kDRows][kDRows];
    const BrickVol& V = find_vol(P, blockIdx.x, true);
    int r = blockIdx.x - V.first_b;
    const int tz = r % V.tiles_z; r /= V.tiles_z;
    const int ty = r % V.tiles_y;
    const int tx = r / V.tiles_y;
    const int ly = threadIdx.x % kDRows, lz = threadIdx.x / kDRows;
    const int by = ty * kDTy - kHalo + ly, bz = tz * kDTz - kHalo + lz;
    const bool row_in = by >= 0 && by < V.nby && bz >= 0 && bz < V.nbz;
    const bool inner = ly >= kHalo && ly < kHalo + kDTy && lz >= kHalo && lz < kHalo + kDTz && row_in;
    const int per_map = V.nbz * V.nby * V.wpb;
    uint32_t dcode[32 / 4];    // result bytes of this row's 32 bricks (inner threads)
#pragma unroll
    for (int i = 0; i < 8; ++i) dcode[i] = 0;

    for (int m = 0; m < 3; ++m) {
        // 64-bit window: bit j <-> brick x = 32 * tx - 16 + j ; everything outside the volume is "constant"
        uint64_t e = ~0ull;
        if (row_in) {
            const uint32_t* row = V.planes + (size_t)m * per_map + ((size_t)bz * V.nby + by) * V.wpb;
            const uint32_t wm = tx > 0 ? __ldg(row + tx - 1) : 0xffffffffu;
            const uint32_t w0 = __ldg(row + tx);
            const uint32_t wp = tx + 1 < V.wpb ? __ldg(row + tx + 1) : 0xffffffffu;
            e = ((uint64_t)(wm >> 16)) | ((uint64_t)w0 << 16) | ((uint64_t)wp << 48);
        }
        // P: this brick and its +x / +y / +z neighbours (2 x 2 x 2 block)
        uint32_t pflag;
        {
            const uint64_t ex = e & ((e >> 1) | (1ull << 63));
            __syncthreads();
            s_e[lz][ly] = ex;
            __syncthreads();
            uint64_t p = ex;
            if (ly < kDRows - 1) p &= s_e[lz][ly + 1];
            if (lz < kDRows - 1) { p &= s_e[lz + 1][ly]; if (ly < kDRows - 1) p &= s_e[lz + 1][ly + 1]; }
            pflag = (uint32_t)(p >> 16);
        }
        uint32_t c0 = 0, c1 = 0, c2 = 0;   // bit-sliced per-brick counter of the erosion levels that still hold
        for (int k = 1; k <= kMaxD; ++k) {
            const uint32_t lvl = (uint32_t)(e >> 16);
            const uint32_t k0 = c0 & lvl; c0 ^= lvl;
            const uint32_t k1 = c1 & k0; c1 ^= k0;
            c2 ^= k1;
            if (k == kMaxD) break;
            // erode by one brick in x, y, z (window / tile edges lose one valid bit / row per pass: the halo)
            e = e & ((e << 1) | 1ull) & ((e >> 1) | (1ull << 63));
            __syncthreads();
            s_e[lz][ly] = e;
            __syncthreads();
            if (ly > 0) e &= s_e[lz][ly - 1];
            if (ly < kDRows - 1) e &= s_e[lz][ly + 1];
            __syncthreads();
            s_e[lz][ly] = e;
            __syncthreads();
            if (lz > 0) e &= s_e[lz - 1][ly];
            if (lz < kDRows - 1) e &= s_e[lz + 1][ly];
        }
        if (inner) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t d = ((c0 >> i) & 1u) | (((c1 >> i) & 1u) << 1) | (((c2 >> i) & 1u) << 2);
                if (d) dcode[i >> 2] |= (((uint32_t)(m + 1) << 4) | (((pflag >> i) & 1u) << 3) | d) << (8 * (i & 3));
            }
        }
    }
    if (inner) {
        uint8_t* out = V.map + ((size_t)bz * V.nby + by) * V.nbx + 32 * tx;
        const int n = min(32, V.nbx - 32 * tx);
        for (int i = 0; i < n; ++i) out[i] = (uint8_t)(dcode[i >> 2] >> (8 * (i & 3)));
    }
}

static inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

}  // namespace emfb

using namespace emfb;

extern "C" EMF_API size_t emf_brick_map_bytes(const int res[3]) {
    if (!res_ok(res)) return 0;
    const size_t nbx = (res[0] + kBrick - 1) / kBrick, nby = (res[1] + kBrick - 1) / kBrick, nbz = (res[2] + kBrick - 1) / kBrick;
    const size_t wpb = (nbx + 31) / 32;
    return align16(nbx * nby * nbz) + 3 * nby * nbz * wpb * sizeof(uint32_t);
}

extern "C" EMF_API int emf_update_brick_maps(int n_vol, const emf_volume* vols, emf_stream_t stream) {
    if (n_vol <= 0 || !vols) return EMF_ERR_INVALID;
    if (n_vol > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    BrickParams P;
    int n = 0;
    int64_t total_a = 0, total_b = 0;
    for (int i = 0; i < n_vol; ++i) {
        const emf_volume& v = vols[i];
        if (!v.const_bits || !v.brick_map) continue;
        if (!res_ok(v.res) || v.res[0] % 4) return EMF_ERR_INVALID;
        BrickVol& d = P.v[n++];
        d.cbits = v.const_bits;
        d.wpr = emf_bitmap_words_per_row(v.res[0]);
        d.ry = v.res[1]; d.rz = v.res[2];
        d.nseg = v.res[0] / 4;
        d.map_words = (size_t)d.wpr * d.ry * d.rz;
        d.nbx = (v.res[0] + kBrick - 1) / kBrick; d.nby = (v.res[1] + kBrick - 1) / kBrick; d.nbz = (v.res[2] + kBrick - 1) / kBrick;
        d.wpb = (d.nbx + 31) / 32;
        d.map = v.brick_map;
        d.planes = (uint32_t*)(v.brick_map + align16((size_t)d.nbx * d.nby * d.nbz));
        d.tiles_y = (d.nby + kDTy - 1) / kDTy; d.tiles_z = (d.nbz + kDTz - 1) / kDTz;
        d.first_a = (int)total_a; d.first_b = (int)total_b;
        // k_brick_codes: whole 128-thread groups per volume so that a CTA never straddles two volumes' index ranges
        total_a += ((int64_t)3 * d.nbz * d.nby * d.wpb + 127) / 128 * 128;
        total_b += (int64_t)d.wpb * d.tiles_y * d.tiles_z;
        if (total_a > 0x7fffffff || total_b > 0x7fffffff) return EMF_ERR_UNSUPPORTED;
    }
    if (n == 0) return EMF_OK;
    P.n_vol = n;
    k_brick_codes<<<(unsigned)(total_a / 128), 128, 0, (cudaStream_t)stream>>>(P);
    k_brick_dist<<<(unsigned)total_b, kDRows * kDRows, 0, (cudaStream_t)stream>>>(P);
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
        if (vol->brick_map) return emf_update_brick_maps(1, vol, stream);
    } else if (vol->brick_map) {
        cudaMemsetAsync(vol->brick_map, 0x00, emf_brick_map_bytes(vol->res), s);
    }
    return launch_status();
}
