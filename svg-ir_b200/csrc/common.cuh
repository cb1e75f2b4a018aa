// Shared device helpers for the svgir_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "svgir_b200.h"

#define TILE 16
#define TILE_PIX 256
#define REC_F4 (SVGIR_REC_FLOATS / 4)

namespace svgir {

void set_error(const char* fmt, ...);
int check_launch(const char* what, bool debug, cudaStream_t stream);

// Optional per-kernel device timing (svgir_timing_enable / svgir_timing_collect in the C ABI):
// when enabled, every launch is bracketed by a pair of CUDA events on the launching stream.
void timing_begin(const char* name, cudaStream_t s);
void timing_end(cudaStream_t s);
struct TimedScope {
    cudaStream_t s;
    TimedScope(const char* name, cudaStream_t st) : s(st) { timing_begin(name, st); }
    ~TimedScope() { timing_end(s); }
};

// ---- exactly-rounded fp32 building blocks ------------------------------------------------
// The binning-relevant chain of the preprocess (projection, culls, covariance, radius, rect,
// depth key) must reproduce the reference build's roundings bit for bit (SURVEY.md 8(a) a7-a9,
// Appendix A.1b). nvcc never contracts or reorders the explicit _rn intrinsics, so the
// contraction pattern read from the reference SASS is spelled out with them.
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqrt_(float a) { return __fsqrt_rn(a); }

// m[i]*x + m[4+i]*y + m[8+i]*z : fma(z,m8, fma(x,m0, rn(y*m4)))
__device__ __forceinline__ float dot3_col(const float* __restrict__ m, int i, float x, float y, float z) {
    return fma_(z, m[8 + i], fma_(x, m[i], mul_(y, m[4 + i])));
}
// a0*b0 + a1*b1 + a2*b2 : fma(a2,b2, fma(a0,b0, rn(a1*b1)))
__device__ __forceinline__ float dot3_glm(float a0, float b0, float a1, float b1, float a2, float b2) {
    return fma_(a2, b2, fma_(a0, b0, mul_(a1, b1)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 128-bit fire-and-forget fp32 vector reduction into global memory (REDG.E.ADD.F32x4, sm_90+);
// addr must be 16-byte aligned
__device__ __forceinline__ void red_add_f32x4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 128-bit read-only streaming load
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- mbarrier (shared::cta) helpers for the compositors' staging ring ---------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive-on triggered by the hardware once every cp.async this thread issued so far has landed; .noinc: the
// arrival counts against the barrier's initial count (one per thread and phase)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    // try_wait suspends the warp in hardware up to the time hint; the polling loop around it only runs when the
    // hardware returns early. Replacing it by a nanosleep back-off measured slower (profiles/r02/composite_bwd_experiments.txt)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---- per-pair blend evaluation shared by the forward and backward compositors ------------
// rec layout: see svgir_raster_state.rec in include/svgir_b200.h
struct PairEval {
    float alpha, G, dx, dy;
};

// svgss: forward.cu:530-547 with the reference build's contraction; rgss: rgss forward.cu:433.
template <bool RGSS>
__device__ __forceinline__ bool eval_alpha(float px, float py, float mx, float my, float cx, float cy,
                                           float cz, float o, PairEval& e) {
    e.dx = sub_(mx, px);
    e.dy = sub_(my, py);
    float power;
    if (!RGSS) {
        float dist = fma_(e.dy, mul_(e.dx, add_(cy, cy)), fma_(e.dx, mul_(e.dx, cx), mul_(e.dy, mul_(e.dy, cz))));
        power = mul_(dist, -0.5f);
    } else {
        // rgss forward.cu:433 / backward.cu:581 as the reference's sm_100 build contracts it (read from its SASS:
        // FMUL dy*cz, FMUL dx*cx, FMUL dx*cy, FMUL dy*(dy*cz), FMUL dy*(dx*cy), FFMA dx*(dx*cx)+.., FFMA (..)*-0.5 - ..)
        power = fma_(fma_(e.dx, mul_(e.dx, cx), mul_(e.dy, mul_(e.dy, cz))), -0.5f, -mul_(e.dy, mul_(e.dx, cy)));
    }
    if (power > 0.0f) return false;
    e.G = expf(power);
    e.alpha = fminf(0.99f, mul_(o, e.G));
    return e.alpha >= 1.0f / 255.0f;
}

// ---- packed fp32x2 FMA (FFMA2, sm_100+): two fused multiply-adds per issued instruction -----
// Same FMA-pipe throughput as two FFMAs (profiles/r01e_ffma2_microbench.txt) but half the issue slots,
// which is what the issue-bound compositors are short of.
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// ---- per-warp footprint cull ------------------------------------------------------------------
// A pixel can only be hit (alpha = min(0.99, o*exp(power)) >= 1/255, forward.cu:541-547) inside the
// ellipse  cx dx^2 + 2 cy dx dy + cz dy^2 <= 2 ln(255 o).  Returns false only when the (slightly
// inflated) bounding box of that ellipse misses the pixel rectangle [X0,X0+XW] x [Y0,Y0+YH]; any
// non-finite or indefinite input returns true, so the exact per-pixel test still decides every hit.
__device__ __forceinline__ bool footprint_overlaps(float mx, float my, float cx, float cy, float cz, float o,
                                                   float X0, float Y0, float XW, float YH) {
    if (o < 1.0f / 255.0f) return false;  // alpha <= o (exp(power) <= 1)
    const float L = fmaf(2.0f * __logf(255.0f * o), 1.001f, 1e-3f);
    const float det = cx * cz - cy * cy;
    if (!(det > 0.f)) return true;
    const float k = L / det;
    const float ex = sqrtf(k * cz) * 1.0005f + 0.05f, ey = sqrtf(k * cx) * 1.0005f + 0.05f;
    return !(mx + ex < X0) && !(mx - ex > X0 + XW) && !(my + ey < Y0) && !(my - ey > Y0 + YH);
}

// 16x16 tile, 8 warps: warp w owns the 8x4 pixel block at ((w&1)*8, (w>>1)*4) -- a compact footprint
// (vs. 16x2 rows) means fewer warps touched per surfel and more hits per touched warp.
#define WARP_PX_W 8
#define WARP_PX_H 4

// Launchers (one per translation unit)
int launch_preprocess(const svgir_raster_cfg& c, const svgir_raster_in& in, svgir_raster_state& st,
                      svgir_raster_out& out, cudaStream_t s);
int launch_tile_scan(const svgir_raster_cfg& c, svgir_raster_state& st, cudaStream_t s);
int launch_binning(const svgir_raster_cfg& c, svgir_raster_state& st, const int32_t* radii, cudaStream_t s);
int launch_composite_fwd(const svgir_raster_cfg& c, const svgir_raster_in& in, svgir_raster_state& st,
                         svgir_raster_out& out, cudaStream_t s);
int launch_pseudo_normal(const svgir_raster_cfg& c, svgir_raster_out& out, cudaStream_t s);
int launch_composite_bwd(const svgir_raster_cfg& c, const svgir_raster_in& in,
                         const svgir_raster_state& st, svgir_raster_grads& g, cudaStream_t s);
int launch_preprocess_bwd(const svgir_raster_cfg& c, const svgir_raster_in& in,
                          const svgir_raster_state& st, const int32_t* radii, svgir_raster_grads& g,
                          cudaStream_t s);
int launch_preprocess_bwd_params(const svgir_raster_cfg& c, const svgir_raster_in& in, const svgir_raster_state& st,
                                 const float* geo_grad, const svgir_param_grads& pg, cudaStream_t s);

}  // namespace svgir
