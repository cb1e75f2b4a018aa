// In-place fp32 sum all-reduce of the flat per-surfel gradient buffer over NVLink peer memory.
//
// New code (the reference is single-GPU: train.py:108-143 renders one view per iteration; SURVEY.md 8(e) asks for
// view-sharded data parallelism with one gradient exchange per step). Each rank's gradient bucket lives in a
// symmetric allocation mapped into every rank of the box, so the backward kernels' output IS the collective's
// input. One kernel per rank:
//   1. put/wait flag barrier with the same-numbered CTA of every peer (all gradients of this step are written);
//   2. rank r owns slice r: sum it over the `world` buffers and write the sum into every buffer -- with NVSwitch
//      multicast one multimem.ld_reduce (the switch adds) and one multimem.st (the switch broadcasts) per 16 bytes,
//      i.e. numel*4/world bytes in and out per GPU; without multicast, `world` 128-bit peer loads and stores;
//   3. flag barrier (every rank's buffer is complete before anything reads it).
// The flags use a compare-and-swap put (0 -> 1 on the peer) / wait (1 -> 0 locally) pair, so they are back at zero
// when the kernel ends and the same launch can be replayed from a CUDA graph.
#include <cstdlib>
#include "common.cuh"

namespace svgir {

struct PeerArgs {
    int world, rank;
    float* bufs[SVGIR_MAX_PEERS];
    unsigned int* flags[SVGIR_MAX_PEERS];
    float* mc;
    long long offset, numel;   // floats
    int bank;
};

__device__ __forceinline__ void put_flag(unsigned int* remote) {
    // release: everything this CTA wrote (made visible by the preceding __threadfence_system) precedes the flag
    while (atomicCAS_system(remote, 0u, 1u) != 0u) {}
}
__device__ __forceinline__ void wait_flag(unsigned int* local) {
    while (atomicCAS_system(local, 1u, 0u) != 1u) {}
}

// Barrier between CTA b of this rank and CTA b of every peer. `phase` selects one of two flag banks.
__device__ __forceinline__ void peer_barrier(const PeerArgs& a, int phase) {
    __syncthreads();
    if (threadIdx.x < a.world && (int)threadIdx.x != a.rank) {
        const int peer = threadIdx.x;
        const int slot = ((a.bank * 2 + phase) * SVGIR_PEER_BLOCKS + blockIdx.x) * SVGIR_MAX_PEERS;
        __threadfence_system();
        put_flag(a.flags[peer] + slot + a.rank);     // "rank a.rank, CTA b arrived" in the peer's area
        wait_flag(a.flags[a.rank] + slot + peer);    // the peer's CTA b arrived here
        __threadfence_system();
    }
    __syncthreads();
}

// WEAK = plain (weak) accesses: every address is read once and written once per launch, strictly between the two
// flag barriers (whose CAS + __threadfence_system order them against the peers), and L1 is invalidated at kernel
// start, so system-scope relaxed accesses are not required for correctness; they are kept as the measured
// alternative (SVGIR_PEER_WEAK=0).
template <bool WEAK>
__device__ __forceinline__ float4 ld_peer(const float* p) {
    float4 v;
    if (WEAK) asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
template <bool WEAK>
__device__ __forceinline__ void st_peer(float* p, float4 v) {
    if (WEAK) asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <bool WEAK>
__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
    float4 v;
    if (WEAK) asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                           : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    else asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
template <bool WEAK>
__device__ __forceinline__ void mc_st(float* p, float4 v) {
    if (WEAK) asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int WORLD, bool MC, bool WEAK>
__global__ void __launch_bounds__(512) peer_allreduce_kernel(const PeerArgs a) {
    peer_barrier(a, 0);
    // slice of this rank, in float4 units
    const long long o4 = a.offset / 4, n4 = (a.numel + 3) / 4;
    const long long per = (n4 + WORLD - 1) / WORLD;
    const long long lo = o4 + per * a.rank, hi = min(o4 + n4, lo + per);
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int U = MC ? 8 : (WORLD <= 2 ? 8 : (WORLD <= 4 ? 4 : 2));   // independent 16-byte requests in flight per thread: U (x WORLD)
    for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * U) {
        if (MC) {
            float4 s[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const long long i = i0 + u * stride;
                if (i < hi) s[u] = mc_ld_reduce<WEAK>(a.mc + 4 * i);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const long long i = i0 + u * stride;
                if (i < hi) mc_st<WEAK>(a.mc + 4 * i, s[u]);
            }
        } else {
            float4 v[U][WORLD];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const long long i = i0 + u * stride;
                if (i < hi) {
#pragma unroll
                    for (int r = 0; r < WORLD; r++) v[u][r] = ld_peer<WEAK>(a.bufs[r] + 4 * i);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const long long i = i0 + u * stride;
                if (i < hi) {
                    float4 s = v[u][0];
#pragma unroll
                    for (int r = 1; r < WORLD; r++) { s.x += v[u][r].x; s.y += v[u][r].y; s.z += v[u][r].z; s.w += v[u][r].w; }
#pragma unroll
                    for (int r = 0; r < WORLD; r++) st_peer<WEAK>(a.bufs[r] + 4 * i, s);
                }
            }
        }
    }
    peer_barrier(a, 1);
}

template <int WORLD>
static void launch_world(const PeerArgs& a, bool mc, bool weak, int grid, cudaStream_t s) {
    if (mc) {
        if (weak) peer_allreduce_kernel<WORLD, true, true><<<grid, 512, 0, s>>>(a);
        else peer_allreduce_kernel<WORLD, true, false><<<grid, 512, 0, s>>>(a);
    } else {
        if (weak) peer_allreduce_kernel<WORLD, false, true><<<grid, 512, 0, s>>>(a);
        else peer_allreduce_kernel<WORLD, false, false><<<grid, 512, 0, s>>>(a);
    }
}

}  // namespace svgir

using namespace svgir;

extern "C" int svgir_peer_allreduce_range(const svgir_peer_comm* comm, long long offset, long long numel, int bank,
                                          int grid_req, void* stream) {
    if (!comm || comm->world < 1 || comm->world > SVGIR_MAX_PEERS || comm->rank < 0 || comm->rank >= comm->world) {
        set_error("peer_allreduce: bad comm (world must be 1..%d)", SVGIR_MAX_PEERS);
        return SVGIR_ERR_INVALID;
    }
    if (numel < 0 || (numel & 3) || offset < 0 || (offset & 3)) {
        set_error("peer_allreduce: offset and numel must be non-negative multiples of 4");
        return SVGIR_ERR_INVALID;
    }
    if (bank < 0 || bank >= SVGIR_PEER_BANKS) { set_error("peer_allreduce: bank %d outside [0,%d)", bank, SVGIR_PEER_BANKS); return SVGIR_ERR_INVALID; }
    if (comm->world == 1 || numel == 0) return SVGIR_OK;
    PeerArgs a;
    a.world = comm->world; a.rank = comm->rank; a.mc = comm->multicast; a.offset = offset; a.numel = numel; a.bank = bank;
    for (int i = 0; i < SVGIR_MAX_PEERS; i++) {
        a.bufs[i] = i < comm->world ? comm->bufs[i] : nullptr;
        a.flags[i] = i < comm->world ? comm->flags[i] : nullptr;
        if (i < comm->world && (!a.bufs[i] || !a.flags[i] || ((uintptr_t)a.bufs[i] & 15))) {
            set_error("peer_allreduce: buffer/flag pointer of rank %d missing or not 16-byte aligned", i);
            return SVGIR_ERR_INVALID;
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    // grid <= SVGIR_PEER_BLOCKS (< one CTA per SM): all CTAs co-resident, each pairs with its peers' twin. The NVLink
    // ports saturate with few CTAs (B200 x8, 104 MB: multicast 0.262 ms at 32 CTAs, 0.279 ms at 128; peer
    // loads/stores 0.32 ms at any grid; NCCL 0.405 ms -- profiles/r01j_peer_allreduce_n8.json)
    const bool mc = a.mc != nullptr;
    int grid = grid_req > 0 ? grid_req : (mc ? 32 : 64);
    if (grid > SVGIR_PEER_BLOCKS) grid = SVGIR_PEER_BLOCKS;
    // tuning knobs (read per call so one process can A/B them): SVGIR_PEER_WEAK=0 -> system-scope relaxed accesses,
    // SVGIR_PEER_GRID=<n> -> CTA count of default-grid launches (must be the same on every rank)
    const char* ev = getenv("SVGIR_PEER_WEAK");
    const bool weak = !(ev && ev[0] == '0');
    if (grid_req <= 0 && (ev = getenv("SVGIR_PEER_GRID")) != nullptr) { const int g = atoi(ev); if (g >= 1 && g <= SVGIR_PEER_BLOCKS) grid = g; }
    { TimedScope ts_("peer_allreduce", s);
      switch (comm->world) {
          case 2: launch_world<2>(a, mc, weak, grid, s); break;
          case 3: launch_world<3>(a, mc, weak, grid, s); break;
          case 4: launch_world<4>(a, mc, weak, grid, s); break;
          case 5: launch_world<5>(a, mc, weak, grid, s); break;
          case 6: launch_world<6>(a, mc, weak, grid, s); break;
          case 7: launch_world<7>(a, mc, weak, grid, s); break;
          default: launch_world<8>(a, mc, weak, grid, s); break;
      } }
    return check_launch("peer_allreduce", false, s);
}

extern "C" int svgir_peer_allreduce(const svgir_peer_comm* comm, long long numel, void* stream) {
    return svgir_peer_allreduce_range(comm, 0, numel, 0, 0, stream);
}
