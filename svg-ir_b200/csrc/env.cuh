// Environment-map helpers shared by the shading kernels (shading.cu) and the radiance-consistency kernels
// (radiance.cu): the lat-long bilinear lookup of DirectLightMap.direct_light (scene/direct_light_map.py:70-83)
// and the fire-and-forget gradient scatter into a [He,We,4] accumulator.
#pragma once
#include "common.cuh"

namespace svgir {

#ifndef PI_F
#define PI_F 3.14159265358979323846f
#endif

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// lat-long bilinear lookup (grid_sample, align_corners=True, zero padding). env [He][We][3].
// Returns texel corner (x0,y0) and weights so the backward can scatter.
struct EnvTap { int x0, y0; float wx1, wy1; };

__device__ __forceinline__ EnvTap env_coords(float dx, float dy, float dz, int He, int We) {
    const float phi = acosf(dz) - 1e-6f;
    const float theta = atan2f(dy, dx);
    const float qy = (phi * (1.f / PI_F)) * 2.f - 1.f;   // one rounding away from phi / pi: < 1e-7 of a texel
    const float qx = -theta * (1.f / PI_F);
    const float ix = (qx + 1.f) / 2.f * (float)(We - 1);
    const float iy = (qy + 1.f) / 2.f * (float)(He - 1);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    EnvTap t;
    t.x0 = (int)fx0; t.y0 = (int)fy0;
    t.wx1 = ix - fx0; t.wy1 = iy - fy0;
    return t;
}

// tap cache encoding (svgir_env_taps): word 0 = x0 | y0 << 16 (signed halves: -1 .. We-1 fit), words 1, 2 = wx1, wy1
__device__ __forceinline__ float env_tap_pack(const EnvTap& t) {
    return __int_as_float((t.x0 & 0xffff) | (t.y0 << 16));
}
__device__ __forceinline__ EnvTap env_tap_unpack(float w0, float wx1, float wy1) {
    const int b = __float_as_int(w0);
    EnvTap t;
    t.x0 = (int)(short)(b & 0xffff); t.y0 = b >> 16;
    t.wx1 = wx1; t.wy1 = wy1;
    return t;
}

__device__ __forceinline__ void env_fetch(const float* env, int He, int We, const EnvTap& t, float out[3]) {
    const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
    out[0] = out[1] = out[2] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = t.x0 + (k & 1), y = t.y0 + (k >> 1);
        if (x < 0 || x > We - 1 || y < 0 || y > He - 1) continue;
        const float w = ((k & 1) ? t.wx1 : wx0) * ((k >> 1) ? t.wy1 : wy0);
        const float* e = env + ((size_t)y * We + x) * 3;
        out[0] = fmaf(e[0], w, out[0]);
        out[1] = fmaf(e[1], w, out[1]);
        out[2] = fmaf(e[2], w, out[2]);
    }
}

// The same lookup on a padded copy of the map, env4 [He][We] float4 (what the shading kernels keep in shared memory):
// four 128-bit loads instead of twelve scalar ones, out-of-range taps handled by a clamped address and a zero weight
// (grid_sample's zero padding) instead of a branch. Same products, same order of additions as env_fetch.
__device__ __forceinline__ void env_fetch4(const float4* env4, int He, int We, const EnvTap& t, float out[3]) {
    const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
    out[0] = out[1] = out[2] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = t.x0 + (k & 1), y = t.y0 + (k >> 1);
        const bool in = x >= 0 && x <= We - 1 && y >= 0 && y <= He - 1;
        const int xc = min(max(x, 0), We - 1), yc = min(max(y, 0), He - 1);
        const float w = ((k & 1) ? t.wx1 : wx0) * ((k >> 1) ? t.wy1 : wy0);
        const float4 e = env4[yc * We + xc];
        if (in) {   // predicated adds: skipping (not adding a zero product) keeps -0 / NaN behaviour of env_fetch
            out[0] = fmaf(e.x, w, out[0]);
            out[1] = fmaf(e.y, w, out[1]);
            out[2] = fmaf(e.z, w, out[2]);
        }
    }
}

// Env-map gradient scatter. sm_100 has no native shared-memory float atomic (atomicAdd on shared
// compiles to a compare-and-swap loop that costs ~2 ms at the training shape), so the four bilinear
// taps of a sample go straight to L2 as four fire-and-forget vector reductions
// (REDG.E.ADD.F32x4) into a [He,We,4] accumulator; a tiny kernel folds it into d_env afterwards. Reductions on one
// address serialise in L2 and 10^7 samples land on a few hundred texels, so the accumulator is replicated and every
// CTA scatters into replica blockIdx % copies (measured on the radiance-consistency backward: 3.4 ms -> 0.8 ms).
__device__ __forceinline__ void red_add_v4(float* addr, float x, float y, float z) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(0.f) : "memory");
}

// defined in shading.cu
void launch_env_activate(int nenv, const float* param, float* act, int env_mode, cudaStream_t s);
void launch_env_grad_finalize(int ntex, int copies, int env_mode, const float* acc, const float* env_param, float* d_env, cudaStream_t s);

}  // namespace svgir
