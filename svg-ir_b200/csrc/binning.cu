// Tile binning: a hand-written tile|depth radix sort organised MSD-first.
//
// The reference (rasterizer_impl.cu:70-138, 307-348) builds (tile<<32 | depth_bits, surfel) pairs
// with an InclusiveSum + duplicateWithKeys, sorts them with a 6-pass CUB LSD radix sort over 44-45
// key bits and then detects tile boundaries.  Here the most significant digit -- the tile id -- is
// resolved by a counting pass (tile_count, filled by the preprocess kernel), an exclusive scan over
// the T tiles (which *is* identifyTileRanges: ranges[t] = (start,end)) and a scatter of
// (depth_bits<<32 | surfel) into each tile's bucket.  Each bucket is then sorted on that 64-bit
// composite inside shared memory.  Because a surfel appears at most once per tile and CUB's sort is
// stable with instances emitted in ascending surfel order, the reference's order inside a tile is
// exactly ascending (depth_bits, surfel) -- a total order, so the result is bit-identical without
// needing a stable algorithm.
//
// HBM traffic per instance: 8 B (scatter) + 8 B read + 4 B write (bucket sort) = 20 B, against
// ~156 B for the 6-pass LSD sort it replaces.  Algorithmic bytes: R*12*2*ceil(44/8).
#include "common.cuh"

namespace svgir {

#define SMALL_CAP 2048
#define MEDIUM_CAP 16384

#define ORDER_BUCKETS 512
__device__ __forceinline__ int order_bucket(uint32_t count) {
    const uint32_t k = count >> 3;
    return ORDER_BUCKETS - 1 - (int)(k < ORDER_BUCKETS - 1 ? k : ORDER_BUCKETS - 1);
}

// Exclusive scan over tile counts; single CTA of 1024 threads.
__global__ void __launch_bounds__(1024) tile_scan_kernel(int T, const uint32_t* __restrict__ count,
                                                         uint32_t* __restrict__ cursor,
                                                         uint2* __restrict__ ranges,
                                                         uint32_t* __restrict__ big, int32_t* __restrict__ num_rendered,
                                                         long long cap_R) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    __shared__ uint32_t n_medium, n_large;
    __shared__ uint32_t hist[ORDER_BUCKETS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { carry_s = 0; n_medium = 0; n_large = 0; }
    for (int i = tid; i < ORDER_BUCKETS; i += 1024) hist[i] = 0;
    __syncthreads();
    // big[0] = #medium, big[1] = #large, big[2 .. 2+T) medium list, big[2+T .. 2+2T) large list,
    // big[2+2T .. 2+3T) compositing order of the tiles (heaviest first, see below)
    for (int base = 0; base < T; base += 1024) {
        const int t = base + tid;
        const uint32_t v = t < T ? count[t] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_tot[lane];
            uint32_t winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += n;
            }
            warp_tot[lane] = winc - w;  // exclusive prefix of warp totals
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t start = carry + warp_tot[wid] + inc - v;
        if (t < T) {
            cursor[t] = start;
            ranges[t] = v ? make_uint2(start, start + v) : make_uint2(0u, 0u);
            if (v > SMALL_CAP) {
                if (v <= MEDIUM_CAP) big[2 + atomicAdd(&n_medium, 1u)] = (uint32_t)t;
                else big[2 + T + atomicAdd(&n_large, 1u)] = (uint32_t)t;
            }
            atomicAdd(&hist[order_bucket(v)], 1u);
        }
        __syncthreads();
        if (tid == 1023) carry_s = start + v;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t R = carry_s;
        num_rendered[0] = (int32_t)R;
        num_rendered[1] = ((long long)R > cap_R) ? 1 : 0;
        big[0] = n_medium;
        big[1] = n_large;
    }
    // Compositing order: tiles by descending instance count (counting sort on count/8, bucket 0 = heaviest). The
    // compositors map blockIdx.x through it, so the hardware block scheduler hands out the long tiles first and the
    // last wave of the grid consists of short ones (longest-processing-time-first; the order inside a bucket is
    // arbitrary and only affects scheduling).
    __syncthreads();
    if (wid == 0) {   // exclusive scan of the histogram, 16 buckets per lane
        uint32_t loc[ORDER_BUCKETS / 32], sum = 0;
#pragma unroll
        for (int i = 0; i < ORDER_BUCKETS / 32; i++) { loc[i] = sum; sum += hist[lane * (ORDER_BUCKETS / 32) + i]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        const uint32_t base = inc - sum;
#pragma unroll
        for (int i = 0; i < ORDER_BUCKETS / 32; i++) hist[lane * (ORDER_BUCKETS / 32) + i] = base + loc[i];
    }
    __syncthreads();
    uint32_t* order = big + 2 + 2 * T;
    for (int t = tid; t < T; t += 1024) order[atomicAdd(&hist[order_bucket(count[t])], 1u)] = (uint32_t)t;
}

// Scatter (depth_bits<<32 | surfel) into the tile buckets (duplicateWithKeys, rasterizer_impl.cu:70-111).
__global__ void __launch_bounds__(256) emit_kernel(int P, int gx, const int32_t* __restrict__ radii,
                                                   const ushort4* __restrict__ rect,
                                                   const float4* __restrict__ rec,
                                                   uint32_t* __restrict__ cursor,
                                                   uint64_t* __restrict__ keys,
                                                   const int32_t* __restrict__ num_rendered) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P || num_rendered[1]) return;
    if (radii[idx] <= 0) return;
    const ushort4 r = rect[idx];
    const uint32_t dbits = __float_as_uint(rec[(size_t)idx * REC_F4 + 1].z);
    const uint64_t kv = ((uint64_t)dbits << 32) | (uint32_t)idx;
    for (int ty = r.y; ty < r.w; ty++)
        for (int tx = r.x; tx < r.z; tx++) {
            const uint32_t pos = atomicAdd(&cursor[ty * gx + tx], 1u);
            keys[pos] = kv;
        }
}

// Bitonic sorting network in the all-ascending formulation: for every merge size k the first
// sub-step compares i with its mirror i^(k-1), the remaining ones compare i with i^j.  Every
// compare-exchange leaves the larger key at the higher index, so keys at indices >= n can be
// *virtual* +inf padding that is never read or written.
template <int NT>
__device__ __forceinline__ void bitonic_sort(uint64_t* a, int n, int m) {
    for (int k = 2; k <= m; k <<= 1) {
        const int hk = k >> 1;
        for (int i = threadIdx.x; i < (m >> 1); i += NT) {
            const int blk = i / hk, off = i - blk * hk;
            const int lo = blk * k + off, hi = blk * k + (k - 1 - off);
            if (hi < n) {
                const uint64_t x = a[lo], y = a[hi];
                if (x > y) { a[lo] = y; a[hi] = x; }
            }
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (m >> 1); i += NT) {
                const int lo = 2 * i - (i & (j - 1));
                const int hi = lo + j;
                if (hi < n) {
                    const uint64_t x = a[lo], y = a[hi];
                    if (x > y) { a[lo] = y; a[hi] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int n) {
    int m = 1;
    while (m < n) m <<= 1;
    return m;
}

template <int NT>
__device__ __forceinline__ void write_bucket(const uint64_t* buf, int tile, const uint2 rg,
                                             uint32_t* __restrict__ point_list,
                                             uint64_t* __restrict__ sorted_keys) {
    const int n = (int)(rg.y - rg.x);
    for (int i = threadIdx.x; i < n; i += NT) {
        const uint64_t kv = buf[i];
        point_list[rg.x + i] = (uint32_t)kv;
        if (sorted_keys) sorted_keys[rg.x + i] = ((uint64_t)(uint32_t)tile << 32) | (kv >> 32);
    }
}

template <int NT>
__device__ __forceinline__ void sort_bucket_smem(uint64_t* buf, int tile, const uint2 rg,
                                                 const uint64_t* __restrict__ keys,
                                                 uint32_t* __restrict__ point_list,
                                                 uint64_t* __restrict__ sorted_keys) {
    const int n = (int)(rg.y - rg.x);
    for (int i = threadIdx.x; i < n; i += NT) buf[i] = keys[rg.x + i];
    __syncthreads();
    bitonic_sort<NT>(buf, n, next_pow2(n));
    write_bucket<NT>(buf, tile, rg, point_list, sorted_keys);
    __syncthreads();
}

// One CTA per tile; buckets of <= SMALL_CAP instances sorted in 16 KB of static shared memory.
__global__ void __launch_bounds__(256) sort_small_kernel(const uint2* __restrict__ ranges,
                                                         const uint64_t* __restrict__ keys,
                                                         uint32_t* __restrict__ point_list,
                                                         uint64_t* __restrict__ sorted_keys,
                                                         const int32_t* __restrict__ num_rendered,
                                                         const uint32_t* __restrict__ tile_order) {
    __shared__ uint64_t buf[SMALL_CAP];
    if (num_rendered[1]) return;
    const int tile = (int)tile_order[blockIdx.x];   // fullest buckets first
    const uint2 rg = ranges[tile];
    const int n = (int)(rg.y - rg.x);
    if (n == 0 || n > SMALL_CAP) return;
    sort_bucket_smem<256>(buf, tile, rg, keys, point_list, sorted_keys);
}

// Persistent CTAs over the medium-tile work list; 128 KB dynamic shared memory.
__global__ void __launch_bounds__(1024) sort_medium_kernel(int T, const uint2* __restrict__ ranges,
                                                           const uint32_t* __restrict__ big,
                                                           const uint64_t* __restrict__ keys,
                                                           uint32_t* __restrict__ point_list,
                                                           uint64_t* __restrict__ sorted_keys,
                                                           const int32_t* __restrict__ num_rendered) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    uint64_t* buf = reinterpret_cast<uint64_t*>(dyn_smem);
    if (num_rendered[1]) return;
    const int n_medium = (int)big[0];
    for (int w = blockIdx.x; w < n_medium; w += gridDim.x) {
        const int tile = (int)big[2 + w];
        sort_bucket_smem<1024>(buf, tile, ranges[tile], keys, point_list, sorted_keys);
    }
}

// Buckets larger than MEDIUM_CAP: the same network run by one CTA in place in global memory
// (`keys` is scratch). Rare: only for tiles with > 16384 overlapping surfels.
__global__ void __launch_bounds__(1024) sort_large_kernel(int T, const uint2* __restrict__ ranges,
                                                          const uint32_t* __restrict__ big,
                                                          uint64_t* __restrict__ keys,
                                                          uint32_t* __restrict__ point_list,
                                                          uint64_t* __restrict__ sorted_keys,
                                                          const int32_t* __restrict__ num_rendered) {
    if (num_rendered[1]) return;
    const int n_large = (int)big[1];
    for (int w = blockIdx.x; w < n_large; w += gridDim.x) {
        const int tile = (int)big[2 + T + w];
        const uint2 rg = ranges[tile];
        const int n = (int)(rg.y - rg.x);
        __syncthreads();
        bitonic_sort<1024>(keys + rg.x, n, next_pow2(n));
        write_bucket<1024>(keys + rg.x, tile, rg, point_list, sorted_keys);
        __syncthreads();
    }
}

int launch_tile_scan(const svgir_raster_cfg& c, svgir_raster_state& st, cudaStream_t s) {
    const int T = ((c.W + TILE - 1) / TILE) * ((c.H + TILE - 1) / TILE);
    { TimedScope ts_("tile_scan", s); tile_scan_kernel<<<1, 1024, 0, s>>>(T, st.tile_count, st.tile_cursor, (uint2*)st.ranges,
                                         st.big_tiles, st.num_rendered, (long long)st.cap_R); }
    return check_launch("tile_scan", c.debug, s);
}

int launch_binning(const svgir_raster_cfg& c, svgir_raster_state& st, const int32_t* radii,
                   cudaStream_t s) {
    const int gx = (c.W + TILE - 1) / TILE, gy = (c.H + TILE - 1) / TILE;
    const int T = gx * gy;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(sort_medium_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 MEDIUM_CAP * 8) != cudaSuccess) {
            set_error("cudaFuncSetAttribute(sort_medium_kernel) failed");
            return SVGIR_ERR_CUDA;
        }
        attr_set = true;
    }
    // cursor[t] was set to start[t] by the scan; tile_scan also re-evaluated the overflow flag
    // against cap_R, which may have changed since svgir_raster_preprocess.
    { TimedScope ts_("tile_scan", s);
      tile_scan_kernel<<<1, 1024, 0, s>>>(T, st.tile_count, st.tile_cursor, (uint2*)st.ranges,
                                         st.big_tiles, st.num_rendered, (long long)st.cap_R); }
    { TimedScope ts_("emit", s); emit_kernel<<<(c.P + 255) / 256, 256, 0, s>>>(c.P, gx, radii, (const ushort4*)st.rect,
                                                  (const float4*)st.rec, st.tile_cursor, st.keys,
                                                  st.num_rendered); }
    int rc = check_launch("emit", c.debug, s);
    if (rc) return rc;
    { TimedScope ts_("sort_small", s); sort_small_kernel<<<T, 256, 0, s>>>((const uint2*)st.ranges, st.keys, st.point_list,
                                        st.sorted_keys, st.num_rendered, st.big_tiles + 2 + 2 * T); }
    { TimedScope ts_("sort_medium", s); sort_medium_kernel<<<148, 1024, MEDIUM_CAP * 8, s>>>(T, (const uint2*)st.ranges, st.big_tiles,
                                                         st.keys, st.point_list, st.sorted_keys,
                                                         st.num_rendered); }
    { TimedScope ts_("sort_large", s); sort_large_kernel<<<148, 1024, 0, s>>>(T, (const uint2*)st.ranges, st.big_tiles, st.keys,
                                           st.point_list, st.sorted_keys, st.num_rendered); }
    return check_launch("tile_sort", c.debug, s);
}

}  // namespace svgir
