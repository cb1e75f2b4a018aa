// LBVH over surfels and the visibility (opacity) ray trace -- the B200-native equivalent of the
// reference's submodules/bvh (thrust LBVH, construct.cu:147-265; thread-per-ray trace, trace.cu:196-286).
//
// How it differs in HOW (the WHAT is bit-identical for the build, rounding-identical for the trace):
//  * no thrust / device_vector allocations: every kernel runs on the caller's stream out of one
//    caller-provided workspace, nothing synchronises the host;
//  * Morton sort = hand-written 8-bit LSD radix sort (block histogram -> scan -> stable scatter with
//    warp match ranks) on (code, index) pairs: same permutation as thrust::stable_sort_by_key;
//  * node boxes are merged bottom-up with one atomic flag per internal node, and the same pass writes
//    a traversal copy of the tree: ONE 64-byte record per internal node holding both children's ids
//    and boxes (4 x LDG.128 per visit instead of 2 x 20 B node + 2 x 24 B box gathers);
//  * leaf data (normal, opacity, mean, Sigma^-1) is packed into one 64-byte record per surfel, with the
//    two cheap rejection inputs (normal, opacity) in the first 16 bytes;
//  * the traversal stack lives in shared memory (interleaved per thread, conflict free) instead of two
//    nested local-memory stacks, and leaf children are evaluated when met instead of being pushed.
#include <cfloat>
#include "common.cuh"

namespace svgir {

static constexpr int SORT_THREADS = 256;
static constexpr int SORT_ITEMS = 4;
static constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

// ------------------------------------------------------------------------------------------------
// RayTracer.__init__ (submodules/bvh/__init__.py:29-57): node / box initialisation and the leaf box
// of the 8 corners mu +- 3 s_a a +- 3 s_b b +- 3 s_c c. The reference evaluates this with separate
// torch kernels, i.e. every product and sum is rounded on its own; the _rn intrinsics reproduce that
// bit for bit (utils/general_utils.py:82-103 for the rotation).
__global__ void bvh_leaf_aabb_kernel(int P, const float* __restrict__ means, const float* __restrict__ scales,
                                     const float4* __restrict__ rots, int32_t* __restrict__ nodes,
                                     float* __restrict__ aabbs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float4 q4 = rots[i];
    const float n = sqrt_(add_(add_(add_(mul_(q4.x, q4.x), mul_(q4.y, q4.y)), mul_(q4.z, q4.z)), mul_(q4.w, q4.w)));
    const float r = div_(q4.x, n), x = div_(q4.y, n), y = div_(q4.z, n), z = div_(q4.w, n);
    float R[3][3];
    R[0][0] = sub_(1.f, mul_(2.f, add_(mul_(y, y), mul_(z, z))));
    R[0][1] = mul_(2.f, sub_(mul_(x, y), mul_(r, z)));
    R[0][2] = mul_(2.f, add_(mul_(x, z), mul_(r, y)));
    R[1][0] = mul_(2.f, add_(mul_(x, y), mul_(r, z)));
    R[1][1] = sub_(1.f, mul_(2.f, add_(mul_(x, x), mul_(z, z))));
    R[1][2] = mul_(2.f, sub_(mul_(y, z), mul_(r, x)));
    R[2][0] = mul_(2.f, sub_(mul_(x, z), mul_(r, y)));
    R[2][1] = mul_(2.f, add_(mul_(y, z), mul_(r, x)));
    R[2][2] = sub_(1.f, mul_(2.f, add_(mul_(x, x), mul_(y, y))));
    const float s0 = mul_(3.f, scales[3 * i]), s1 = mul_(3.f, scales[3 * i + 1]), s2 = mul_(3.f, scales[3 * i + 2]);
    float* bb = aabbs + (size_t)(P - 1 + i) * 6;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float m = means[3 * i + k];
        const float A = mul_(R[k][0], s0), B = mul_(R[k][1], s1), C = mul_(R[k][2], s2);
        float lo = FLT_MAX, hi = -FLT_MAX;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            float v = (c & 4) ? sub_(m, A) : add_(m, A);
            v = (c & 2) ? sub_(v, B) : add_(v, B);
            v = (c & 1) ? sub_(v, C) : add_(v, C);
            lo = fminf(lo, v);
            hi = fmaxf(hi, v);
        }
        bb[k] = lo;
        bb[3 + k] = hi;
    }
    int32_t* nd = nodes + (size_t)(P - 1 + i) * 5;
    nd[0] = -1; nd[1] = -1; nd[2] = -1; nd[3] = -1; nd[4] = 1;
    if (i < P - 1) {
        int32_t* ni = nodes + (size_t)i * 5;
        ni[0] = -1; ni[1] = -1; ni[2] = -1; ni[3] = -1; ni[4] = 0;
        float* bi = aabbs + (size_t)i * 6;
        bi[0] = bi[1] = bi[2] = 100000.f;
        bi[3] = bi[4] = bi[5] = -100000.f;
    }
}

// monotone float <-> uint map so that atomicMin/Max on the bits orders like the floats
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// scene box = merge of all leaf boxes starting from (+1e5, -1e5) (construct.cu:159-168); also keeps a
// copy of the un-permuted leaf boxes for the gather after the sort.
__global__ void bvh_scene_box_kernel(int P, const float* __restrict__ leaf_aabbs, float* __restrict__ leaf_copy,
                                     unsigned* __restrict__ box_ord) {
    float lo[3] = {100000.f, 100000.f, 100000.f}, hi[3] = {-100000.f, -100000.f, -100000.f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const float2* src = reinterpret_cast<const float2*>(leaf_aabbs + (size_t)i * 6);
        float2* dst = reinterpret_cast<float2*>(leaf_copy + (size_t)i * 6);
        const float2 a = src[0], b = src[1], c = src[2];
        dst[0] = a; dst[1] = b; dst[2] = c;
        lo[0] = fminf(lo[0], a.x); lo[1] = fminf(lo[1], a.y); lo[2] = fminf(lo[2], b.x);
        hi[0] = fmaxf(hi[0], b.y); hi[1] = fmaxf(hi[1], c.x); hi[2] = fmaxf(hi[2], c.y);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(box_ord + k, f2ord(lo[k]));
            atomicMax(box_ord + 3 + k, f2ord(hi[k]));
        }
    }
}

__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {  // construct.cu:6-14
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// morton_code_calculator (construct.cu:22-52): centroid -> unit cube of the scene box -> 10 bits/axis
__global__ void bvh_morton_kernel(int P, const float* __restrict__ leaf_copy, const unsigned* __restrict__ box_ord,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float wlo = ord2f(box_ord[k]), whi = ord2f(box_ord[3 + k]);
        float c = mul_(add_(leaf_copy[(size_t)i * 6 + 3 + k], leaf_copy[(size_t)i * 6 + k]), 0.5f);  // utility.cuh:11-19
        c = sub_(c, wlo);
        c = div_(c, sub_(whi, wlo));
        q[k] = (uint32_t)fminf(fmaxf(mul_(c, 1024.0f), 0.0f), 1023.0f);
    }
    const uint32_t code = expand_bits(q[0]) * 4 + expand_bits(q[1]) * 2 + expand_bits(q[2]);
    keys[i] = code;
    idx[i] = (uint32_t)i;
}

// ---- stable LSD radix sort, 8-bit digits ---------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(int n, int nblk, int shift, const uint32_t* __restrict__ keys,
                                                                 uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const int g = base + r * SORT_THREADS + threadIdx.x;
        if (g < n) atomicAdd(&h[(keys[g] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `n` counters in place (single CTA; n = 256 * nblk, a few 100k at most)
__global__ void __launch_bounds__(1024) sort_scan_kernel(int n, uint32_t* __restrict__ hist) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024 * 4) {
        const int i0 = base + threadIdx.x * 4;
        uint32_t v[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { v[k] = (i0 + k < n) ? hist[i0 + k] : 0u; s += v[k]; }
        uint32_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        if (w == 0) {
            uint32_t t = warp_tot[lane], ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += u;
            }
            warp_tot[lane] = ti - t;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        uint32_t ex = carry + warp_tot[w] + inc - s;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < n) hist[i0 + k] = ex;
            ex += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[w] + inc;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(int n, int nblk, int shift, const uint32_t* __restrict__ keys_in,
                                                                    const uint32_t* __restrict__ vals_in, const uint32_t* __restrict__ offs,
                                                                    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t warp_cnt[SORT_THREADS / 32][256];
    __shared__ uint32_t run[256];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    run[tid] = 0;
    uint32_t key[SORT_ITEMS], rank[SORT_ITEMS];
    const int base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const int g = base + r * SORT_THREADS + tid;
        const bool valid = g < n;
        key[r] = valid ? keys_in[g] : 0u;
        const uint32_t digit = valid ? ((key[r] >> shift) & 255u) : 256u + lane;  // invalid lanes match nobody
#pragma unroll
        for (int k = 0; k < SORT_THREADS / 32; k++) warp_cnt[k][tid] = 0;
        __syncthreads();
        const unsigned m = __match_any_sync(0xffffffffu, digit);
        const uint32_t below = __popc(m & ((1u << lane) - 1u));
        if (valid && below == 0) warp_cnt[w][digit] = __popc(m);
        __syncthreads();
        if (valid) {
            uint32_t pre = run[digit];
            for (int k = 0; k < w; k++) pre += warp_cnt[k][digit];
            rank[r] = pre + below;
        }
        __syncthreads();
        uint32_t tot = 0;
#pragma unroll
        for (int k = 0; k < SORT_THREADS / 32; k++) tot += warp_cnt[k][tid];
        run[tid] += tot;
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const int g = base + r * SORT_THREADS + tid;
        if (g < n) {
            const uint32_t digit = (key[r] >> shift) & 255u;
            const uint32_t dst = offs[(size_t)digit * nblk + blockIdx.x] + rank[r];
            keys_out[dst] = key[r];
            vals_out[dst] = vals_in[g];
        }
    }
}

// sorted leaves: object id, permuted box, 64-bit code (m << 31) | idx (construct.cu:182-203)
__global__ void bvh_leaves_kernel(int P, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ idx,
                                  const float* __restrict__ leaf_copy, int32_t* __restrict__ nodes,
                                  float* __restrict__ aabbs, uint64_t* __restrict__ morton) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t obj = idx[i];
    morton[i] = ((uint64_t)keys[i] << 31) | (uint64_t)obj;
    const float2* src = reinterpret_cast<const float2*>(leaf_copy + (size_t)obj * 6);
    float2* dst = reinterpret_cast<float2*>(aabbs + (size_t)(P - 1 + i) * 6);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    int32_t* nd = nodes + (size_t)(P - 1 + i) * 5;
    nd[1] = -1; nd[2] = -1; nd[3] = (int32_t)obj; nd[4] = 1;
    if (P == 1) nd[0] = -1;
}

__device__ __forceinline__ int cub64(uint64_t a, uint64_t b) { return __clzll((long long)(a ^ b)); }

// Karras ranges / splits (construct.cu:54-145, 204-227). count = size of the node's leaf range, which
// is what the reference's atomic bottom-up counting (construct.cu:240,257) converges to.
__global__ void bvh_internal_kernel(int P, const uint64_t* __restrict__ code, int32_t* __restrict__ nodes,
                                    int* __restrict__ flags) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P - 1) return;
    flags[idx] = 0;
    int lo, hi;
    if (idx == 0) {
        lo = 0; hi = P - 1;
    } else {
        const uint64_t self = code[idx];
        const int L = cub64(self, code[idx - 1]), R = cub64(self, code[idx + 1]);
        const int d = (R > L) ? 1 : -1;
        const int dmin = min(L, R);
        int lmax = 2, delta = -1;
        long long it = (long long)idx + (long long)d * lmax;
        if (0 <= it && it < P) delta = cub64(self, code[it]);
        while (delta > dmin) {
            lmax <<= 1;
            it = (long long)idx + (long long)d * lmax;
            delta = -1;
            if (0 <= it && it < P) delta = cub64(self, code[it]);
        }
        int l = 0;
        for (int t = lmax >> 1; t > 0; t >>= 1) {
            it = (long long)idx + (long long)(l + t) * d;
            delta = -1;
            if (0 <= it && it < P) delta = cub64(self, code[it]);
            if (delta > dmin) l += t;
        }
        const int j = idx + l * d;
        lo = min(idx, j); hi = max(idx, j);
    }
    int split;
    {
        const uint64_t fc = code[lo], lc = code[hi];
        if (fc == lc) {
            split = (lo + hi) >> 1;
        } else {
            const int dn = cub64(fc, lc);
            split = lo;
            int stride = hi - lo;
            do {
                stride = (stride + 1) >> 1;
                const int mid = split + stride;
                if (mid < hi && cub64(fc, code[mid]) > dn) split = mid;
            } while (stride > 1);
        }
    }
    int left = split, right = split + 1;
    if (lo == split) left += P - 1;
    if (hi == split + 1) right += P - 1;
    int32_t* nd = nodes + (size_t)idx * 5;
    nd[1] = left; nd[2] = right; nd[3] = -1; nd[4] = hi - lo + 1;
    if (idx == 0) nd[0] = -1;
    nodes[(size_t)left * 5] = idx;
    nodes[(size_t)right * 5] = idx;
}

// Bottom-up box merge (construct.cu:229-263): the second thread to arrive at a node merges its
// children's boxes. The same thread writes the node's traversal record.
// packed[4*id] = { left, right (>=0 internal id, <0 ~object), lmin.x, lmin.y | lmin.z, lmax.xyz |
//                  rmin.xyz, rmax.x | rmax.y, rmax.z, 0, 0 }
__global__ void bvh_merge_kernel(int P, const int32_t* __restrict__ nodes, float* aabbs, int* flags,
                                 float4* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    int node = P - 1 + i;
    int parent = nodes[(size_t)node * 5];
    while (parent >= 0) {
        __threadfence();
        if (atomicAdd(flags + parent, 1) == 0) return;
        __threadfence();
        const int l = nodes[(size_t)parent * 5 + 1], r = nodes[(size_t)parent * 5 + 2];
        float lb[6], rb[6];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            lb[k] = __ldcg(aabbs + (size_t)l * 6 + k);
            rb[k] = __ldcg(aabbs + (size_t)r * 6 + k);
        }
        float* pb = aabbs + (size_t)parent * 6;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            __stcg(pb + k, fminf(lb[k], rb[k]));
            __stcg(pb + 3 + k, fmaxf(lb[3 + k], rb[3 + k]));
        }
        const int le = l < P - 1 ? l : ~nodes[(size_t)l * 5 + 3];
        const int re = r < P - 1 ? r : ~nodes[(size_t)r * 5 + 3];
        float4* rec = packed + (size_t)parent * 4;
        rec[0] = make_float4(__int_as_float(le), __int_as_float(re), lb[0], lb[1]);
        rec[1] = make_float4(lb[2], lb[3], lb[4], lb[5]);
        rec[2] = make_float4(rb[0], rb[1], rb[2], rb[3]);
        rec[3] = make_float4(rb[4], rb[5], 0.f, 0.f);
        node = parent;
        parent = nodes[(size_t)node * 5];
    }
}

// leaf record: { n.xyz, opacity | mu.xyz, S0 | S1..S4 | S5, 0, 0, 0 }, Sigma^-1 as strip_symmetric orders it
__global__ void bvh_pack_leaves_kernel(int P, const float* __restrict__ means, const float* __restrict__ cov_inv,
                                       const float* __restrict__ opac, const float* __restrict__ normals,
                                       float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float* c = cov_inv + (size_t)i * 6;
    out[(size_t)i * 4 + 0] = make_float4(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2], opac[i]);
    out[(size_t)i * 4 + 1] = make_float4(means[3 * i], means[3 * i + 1], means[3 * i + 2], c[0]);
    out[(size_t)i * 4 + 2] = make_float4(c[1], c[2], c[3], c[4]);
    out[(size_t)i * 4 + 3] = make_float4(c[5], 0.f, 0.f, 0.f);
}

// ---- trace ----------------------------------------------------------------------------------------
// utility.cuh:35-83 verbatim in behaviour (IEEE divisions, the same swap / reject order): returns tmax,
// or -1 on a miss. Only "tmax > 0" is used by the traversal (trace.cu:259-272).
__device__ __forceinline__ float ray_box_tmax(float lx, float ly, float lz, float ux, float uy, float uz,
                                              float ox, float oy, float oz, float dx, float dy, float dz) {
    float tmin = div_(sub_(lx, ox), dx), tmax = div_(sub_(ux, ox), dx);
    if (tmin > tmax) { const float t = tmin; tmin = tmax; tmax = t; }
    float tymin = div_(sub_(ly, oy), dy), tymax = div_(sub_(uy, oy), dy);
    if (tymin > tymax) { const float t = tymin; tymin = tymax; tymax = t; }
    if (tmin > tymax || tymin > tmax) return -1.f;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = div_(sub_(lz, oz), dz), tzmax = div_(sub_(uz, oz), dz);
    if (tzmin > tzmax) { const float t = tzmin; tzmin = tzmax; tzmax = t; }
    if (tmin > tzmax || tzmin > tmax) return -1.f;
    if (tzmax < tmax) tmax = tzmax;
    return tmax;
}

static constexpr int TRACE_THREADS = 128;
static constexpr int TRACE_STACK = 64;

struct RayState {
    float ox, oy, oz, dx, dy, dz, T;
    int count;
    bool dead;
};

// trace.cu:231-252 leaf test (opacity / back-face / t / power rejections, alpha, early-out at T < 0.9)
__device__ __forceinline__ void leaf_eval(const float4* __restrict__ leaves, int obj, RayState& r) {
    const float4 v0 = __ldg(leaves + (size_t)obj * 4);
    if (v0.w < 1.f / 255.f) return;
    if (v0.x * r.dx + v0.y * r.dy + v0.z * r.dz > 0.f) return;
    const float4 v1 = __ldg(leaves + (size_t)obj * 4 + 1);
    const float4 v2 = __ldg(leaves + (size_t)obj * 4 + 2);
    const float c5 = __ldg(reinterpret_cast<const float*>(leaves + (size_t)obj * 4 + 3));
    const float c0 = v1.w, c1 = v2.x, c2 = v2.y, c3 = v2.z, c4 = v2.w;
    const float mx = v1.x - r.ox, my = v1.y - r.oy, mz = v1.z - r.oz;
    const float t1 = c0 * mx * r.dx + c1 * mx * r.dy + c2 * mx * r.dz + c1 * my * r.dx + c3 * my * r.dy + c4 * my * r.dz +
                     c2 * mz * r.dx + c4 * mz * r.dy + c5 * mz * r.dz;
    const float t2 = c0 * r.dx * r.dx + c1 * r.dx * r.dy + c2 * r.dx * r.dz + c1 * r.dy * r.dx + c3 * r.dy * r.dy +
                     c4 * r.dy * r.dz + c2 * r.dz * r.dx + c4 * r.dz * r.dy + c5 * r.dz * r.dz;
    const float t = t1 / t2;
    if (t <= 0.01f) return;  // reference: `t < 0.01` evaluated in double, and float(0.01) < 0.01; NaN falls through as there
    const float px = r.ox + t * r.dx, py = r.oy + t * r.dy, pz = r.oz + t * r.dz;
    const float ex = v1.x - px, ey = v1.y - py, ez = v1.z - pz;
    const float power = -0.5f * (ex * ex * c0 + ey * ey * c3 + ez * ez * c5 + 2 * ex * ey * c1 + 2 * ex * ez * c2 +
                                 2 * ey * ez * c4);
    if (power > 0.f) return;
    r.count += 1;
    const float alpha = v0.w * __expf(power);
    r.T *= 1 - alpha;
    if (r.T <= 0.9f) r.dead = true;  // reference: `ray_opacity < 0.9` in double; float(0.9) < 0.9
}

// One thread per ray. ray r: origin rays_o[r / o_div] (+ t_off * d, as RayTracer.trace_visibility adds
// it: submodules/bvh/__init__.py:63), direction rays_d[r].
__global__ void __launch_bounds__(TRACE_THREADS) bvh_trace_opacity_kernel(
    long long n_rays, int P, int root_obj, const float4* __restrict__ packed, const float4* __restrict__ leaves,
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, int o_div, float t_off,
    int32_t* __restrict__ contributes, float* __restrict__ visibility) {
    __shared__ int stack[TRACE_STACK][TRACE_THREADS];
    const long long ray = (long long)blockIdx.x * TRACE_THREADS + threadIdx.x;
    if (ray >= n_rays) return;
    const int tid = threadIdx.x;
    RayState r;
    r.dx = rays_d[3 * ray]; r.dy = rays_d[3 * ray + 1]; r.dz = rays_d[3 * ray + 2];
    const long long oi = o_div > 1 ? ray / o_div : ray;
    r.ox = rays_o[3 * oi]; r.oy = rays_o[3 * oi + 1]; r.oz = rays_o[3 * oi + 2];
    if (t_off != 0.f) {
        r.ox = add_(r.ox, mul_(r.dx, t_off)); r.oy = add_(r.oy, mul_(r.dy, t_off)); r.oz = add_(r.oz, mul_(r.dz, t_off));
    }
    r.T = 1.0f; r.count = 0; r.dead = false;
    int sp = 0;
    if (P == 1) leaf_eval(leaves, root_obj, r);
    else stack[sp++][tid] = 0;
    while (sp > 0 && !r.dead) {
        const int id = stack[--sp][tid];
        const float4 a = __ldg(packed + (size_t)id * 4), b = __ldg(packed + (size_t)id * 4 + 1);
        const float4 c = __ldg(packed + (size_t)id * 4 + 2), d = __ldg(packed + (size_t)id * 4 + 3);
        const int l = __float_as_int(a.x), rr = __float_as_int(a.y);
        const float lt = ray_box_tmax(a.z, a.w, b.x, b.y, b.z, b.w, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz);
        const float rt = ray_box_tmax(c.x, c.y, c.z, c.w, d.x, d.y, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz);
        // the reference pops the child with the smaller tmax first (trace.cu:259-272); leaves are
        // evaluated at once (the result does not depend on the visiting order)
        const bool l_first = !(lt > rt);
        const int first = l_first ? l : rr, second = l_first ? rr : l;
        const float ft = l_first ? lt : rt, st = l_first ? rt : lt;
        if (st > 0.f && second >= 0 && sp < TRACE_STACK) stack[sp++][tid] = second;
        if (ft > 0.f) {
            if (first < 0) leaf_eval(leaves, ~first, r);
            else if (sp < TRACE_STACK) stack[sp++][tid] = first;
        }
        if (st > 0.f && second < 0 && !r.dead) leaf_eval(leaves, ~second, r);
    }
    contributes[ray] = r.dead ? 0 : r.count;  // an early-out ray keeps the zero it was initialised with (bvh.cu:103)
    visibility[ray] = r.dead ? 0.0f : r.T;
}

struct BvhWs {
    float* leaf_copy; uint32_t *k0, *k1, *v0, *v1, *hist; unsigned* box; int* flags;
};
static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t carve(int P, char* base, BvhWs* w) {
    const int nblk = (P + SORT_TILE - 1) / SORT_TILE;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += al(bytes); return p; };
    char* p;
    p = take((size_t)P * 24); if (w) w->leaf_copy = (float*)p;
    p = take((size_t)P * 4); if (w) w->k0 = (uint32_t*)p;
    p = take((size_t)P * 4); if (w) w->k1 = (uint32_t*)p;
    p = take((size_t)P * 4); if (w) w->v0 = (uint32_t*)p;
    p = take((size_t)P * 4); if (w) w->v1 = (uint32_t*)p;
    p = take((size_t)256 * nblk * 4); if (w) w->hist = (uint32_t*)p;
    p = take(32); if (w) w->box = (unsigned*)p;
    p = take((size_t)(P > 1 ? P - 1 : 1) * 4); if (w) w->flags = (int*)p;
    return off;
}

__global__ void bvh_box_init_kernel(unsigned* box) {
    if (threadIdx.x < 3) box[threadIdx.x] = f2ord(100000.f);
    else if (threadIdx.x < 6) box[threadIdx.x] = f2ord(-100000.f);
}

}  // namespace svgir

using namespace svgir;

extern "C" {

size_t svgir_bvh_workspace_bytes(int P) { return P > 0 ? carve(P, nullptr, nullptr) : 0; }

int svgir_bvh_leaf_aabbs(int P, const float* means3D, const float* scales, const float* rotations,
                         int32_t* nodes, float* aabbs, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return SVGIR_OK;
    if (!means3D || !scales || !rotations || !nodes || !aabbs) { set_error("bvh_leaf_aabbs: null pointer"); return SVGIR_ERR_INVALID; }
    if ((uintptr_t)rotations & 15) { set_error("bvh_leaf_aabbs: rotations must be 16-byte aligned"); return SVGIR_ERR_INVALID; }
    { TimedScope t_("bvh_leaf_aabb", s);
      bvh_leaf_aabb_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, scales, (const float4*)rotations, nodes, aabbs); }
    return check_launch("bvh_leaf_aabb", false, s);
}

int svgir_bvh_build(const svgir_bvh* b, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!b || b->P <= 0) return SVGIR_OK;
    const int P = b->P;
    if (!b->nodes || !b->aabbs || !b->morton || !b->packed || !b->workspace) { set_error("bvh_build: null pointer"); return SVGIR_ERR_INVALID; }
    if (b->workspace_bytes < svgir_bvh_workspace_bytes(P)) { set_error("bvh_build: workspace too small"); return SVGIR_ERR_CAPACITY; }
    if (((uintptr_t)b->workspace & 255) || ((uintptr_t)b->packed & 15) || ((uintptr_t)b->aabbs & 7)) { set_error("bvh_build: misaligned buffer"); return SVGIR_ERR_INVALID; }
    BvhWs w;
    carve(P, (char*)b->workspace, &w);
    const float* leaf = b->aabbs + (size_t)(P - 1) * 6;
    const int g = (P + 255) / 256;
    { TimedScope t_("bvh_scene_box", s);
      bvh_box_init_kernel<<<1, 32, 0, s>>>(w.box);
      bvh_scene_box_kernel<<<min(g, 148 * 8), 256, 0, s>>>(P, leaf, w.leaf_copy, w.box); }
    { TimedScope t_("bvh_morton", s); bvh_morton_kernel<<<g, 256, 0, s>>>(P, w.leaf_copy, w.box, w.k0, w.v0); }
    const int nblk = (P + SORT_TILE - 1) / SORT_TILE;
    uint32_t *ki = w.k0, *ko = w.k1, *vi = w.v0, *vo = w.v1;
    for (int shift = 0; shift < 32; shift += 8) {  // 30 code bits -> 4 digits
        TimedScope t_("bvh_sort", s);
        sort_hist_kernel<<<nblk, SORT_THREADS, 0, s>>>(P, nblk, shift, ki, w.hist);
        sort_scan_kernel<<<1, 1024, 0, s>>>(256 * nblk, w.hist);
        sort_scatter_kernel<<<nblk, SORT_THREADS, 0, s>>>(P, nblk, shift, ki, vi, w.hist, ko, vo);
        uint32_t* t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    { TimedScope t_("bvh_leaves", s); bvh_leaves_kernel<<<g, 256, 0, s>>>(P, ki, vi, w.leaf_copy, b->nodes, b->aabbs, b->morton); }
    if (P > 1) {
        { TimedScope t_("bvh_internal", s); bvh_internal_kernel<<<(P - 1 + 255) / 256, 256, 0, s>>>(P, b->morton, b->nodes, w.flags); }
        { TimedScope t_("bvh_merge", s); bvh_merge_kernel<<<g, 256, 0, s>>>(P, b->nodes, b->aabbs, w.flags, (float4*)b->packed); }
    }
    return check_launch("bvh_build", false, s);
}

int svgir_bvh_pack_leaves(int P, const float* means3D, const float* symm_inv, const float* opacity,
                          const float* normals, float* leaf_records, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return SVGIR_OK;
    if (!means3D || !symm_inv || !opacity || !normals || !leaf_records) { set_error("bvh_pack_leaves: null pointer"); return SVGIR_ERR_INVALID; }
    if ((uintptr_t)leaf_records & 15) { set_error("bvh_pack_leaves: leaf_records must be 16-byte aligned"); return SVGIR_ERR_INVALID; }
    { TimedScope t_("bvh_pack_leaves", s);
      bvh_pack_leaves_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, symm_inv, opacity, normals, (float4*)leaf_records); }
    return check_launch("bvh_pack_leaves", false, s);
}

int svgir_bvh_trace_opacity(const svgir_bvh* b, long long n_rays, const float* rays_o, const float* rays_d,
                            int rays_per_origin, float origin_offset, const float* leaf_records,
                            int32_t* contributes, float* visibility, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n_rays <= 0) return SVGIR_OK;
    if (!b || b->P <= 0 || !b->nodes || !b->packed) { set_error("bvh_trace_opacity: tree missing"); return SVGIR_ERR_INVALID; }
    if (!rays_o || !rays_d || !leaf_records || !contributes || !visibility) { set_error("bvh_trace_opacity: null pointer"); return SVGIR_ERR_INVALID; }
    if (rays_per_origin < 1) { set_error("bvh_trace_opacity: rays_per_origin < 1"); return SVGIR_ERR_INVALID; }
    const long long blocks = (n_rays + TRACE_THREADS - 1) / TRACE_THREADS;
    if (blocks > 0x7fffffffLL) { set_error("bvh_trace_opacity: too many rays for one launch"); return SVGIR_ERR_INVALID; }
    int root_obj = 0;  // P == 1: the root is the only leaf and its object id is 0
    { TimedScope t_("bvh_trace_opacity", s);
      bvh_trace_opacity_kernel<<<(unsigned)blocks, TRACE_THREADS, 0, s>>>(n_rays, b->P, root_obj, (const float4*)b->packed,
                                                                        (const float4*)leaf_records, rays_o, rays_d,
                                                                        rays_per_origin, origin_offset, contributes, visibility); }
    return check_launch("bvh_trace_opacity", false, s);
}

}  // extern "C"
