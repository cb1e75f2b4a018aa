// Fused PBR render_equation: env-map importance samples x spatially-varying GGX BRDF, per surfel.
//
// Replaces the reference's live shading path -- rendering_equation4 + GGX_specular4
// (gaussian_renderer/svgss.py:537-631) and DirectLightMap/EnvLight.direct_light
// (scene/direct_light_map.py:70-106, scene/envmap.py:54-72) -- which runs as ~40 elementwise torch
// kernels each streaming [N,Ns,12] fp32 through HBM, forward and again backward with autograd
// saving most intermediates.  Here one kernel reads every per-sample input exactly once
// (N*Ns*32 B), keeps the (soft-plus'd) environment map in shared memory and writes only [N,12]
// results; the backward kernel recomputes the per-sample terms instead of loading saved ones.
//
// Mapping: one warp per surfel, one lane per light sample.  Every per-sample buffer
// ([N,Ns,3] dirs and radiance, [N,Ns] visibility and areas) is read with fully coalesced loads exactly
// once; the vertex-independent work of a sample (env lookup, half vector, Fresnel power) is done once
// by its lane, which then evaluates all 4 vertices; the 24-48 per-(vertex,channel) sums are combined
// with one "transposed" warp reduction (~1 shuffle per value instead of 5).  The grid is persistent
// (a few CTAs per SM striding over the surfels) so the env map is staged into shared memory once per CTA.
//
// Algebra: with T = L*area*max(N.w,0), pbr = mean((f_d+f_s) T) = f_d*mean(T) + F0*mean(D T) +
// (1-F0)*mean(p D T) where f_s = (F0 + (1-F0) p) D, p = 2^((-5.55473 VoH - 6.98316) VoH) and
// D = a^2/clamp(4 pi nom0^2 nom1 nom2).  Accumulating mean(T), mean(D T), mean(p D T) separately
// for the env ("direct") and cached-radiance ("indirect") light gives every output of the
// reference, and makes the optional per-vertex metallic (F0 = 0.04(1-m) + base m,
// f_d = (1-m) base/pi; render_equation.cu:55-190 of the legacy kernel) free.
//
// Roofline: HBM-bound, algorithmic bytes N*(Ns*32 + 124) + N*4*(60+S).
#include "common.cuh"
#include "env.cuh"

namespace svgir {

#define SH_THREADS 256   // forward: ~120 regs -> 2 CTAs/SM
#define SHB_THREADS 128  // backward: ~165 regs -> 3 CTAs/SM
#ifndef SHB_MIN_CTAS
#define SHB_MIN_CTAS 3
#endif
#define VUF_FLOATS 12    // forward per-vertex uniform block in shared memory
#define VU_FLOATS 24     // backward per-vertex uniform block (see shade_bwd_kernel)

struct SampleShared {  // vertex-independent per-sample quantities, produced by one lane of the quad
    float wx, wy, wz;      // raw incident direction
    float lx, ly, lz;      // normalised
    float hx, hy, hz;      // half vector
    float hlen;            // |(L+V)/2|
    float voh_raw, p;      // V.H before clamp, 2^FMi
    float gr, gg, gb;      // area * clamp(env)*vis
    float lr, lg, lb;      // area * radiance
};

__global__ void env_activate_kernel(int n, const float* __restrict__ param, float* __restrict__ act, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) act[i] = mode == 0 ? softplus_f(param[i]) : param[i];
}

// ---------------------------------------------------------------------------------------------
struct ShadeArgs {
    int N, Ns, He, We;
    float env_scale;          // 2.0 for the learnable map (direct_light_map.py:83), 1.0 for HDR maps
    const float* env_act;     // [He,We,3] activated env
    const float* transform;   // optional [3,3] applied to dirs before the lookup (envmap.py:58-61)
    const float* view3x3;     // optional: world->view rotation rows (viewmatrix[:3,:3]) for the packed normals
    const float* base_color;  // [N,12] channel-major
    const float* roughness;   // [N,4]
    const float* metallic;    // [N,4] or null
    const float* normals;     // [N,4,3]
    const float* viewdirs;    // [N,3]
    const float* radiance;    // [N,Ns,3]
    const float* visibility;  // [N,Ns]
    const float* dirs;        // [N,Ns,3]
    const float* areas;       // [N,Ns]
    const int32_t* list;       // optional work list: shade only surfels list[0 .. *list_count)
    const int32_t* list_count;
    const float* means3D;      // viewdirs == nullptr: view vector = campos - means3D[n] (normalised in the kernel)
    const float* campos;
    const int32_t* skip_flag;  // backward only: return immediately when *skip_flag != 0
    const float* taps;         // optional [N,Ns,3] cached env taps (svgir_env_taps): no acos / atan2 per sample
    int view_stride;           // row stride of view3x3: 3, or 4 when it points at the 4x4 world-view matrix itself
};

// Surfel handled by work slot `slot` (identity without a work list); a.N once the work is exhausted.
__device__ __forceinline__ int surfel_at(const ShadeArgs& a, int slot, int count) {
    return slot < count ? (a.list ? __ldg(a.list + slot) : slot) : a.N;
}

// Sum over the warp of NP (a power of two <= 32) per-lane values, "transposed": instead of 5 shuffles
// per value, each step exchanges half of the remaining slots, so the whole reduction costs ~NP
// shuffles.  Returns on every lane L the total of slot  L >> (5 - log2(NP)).  CUR / O are template
// parameters so that every slot index is a compile-time constant (registers, no predication).
template <int NP, int CUR, int O>
struct WarpSlotReduce {
    static __device__ __forceinline__ void run(float (&v)[NP], int lane) {
        const unsigned full = 0xffffffffu;
        if constexpr (CUR > 1) {
            constexpr int HALF = CUR / 2;
            const bool up = (lane & O) != 0;
#pragma unroll
            for (int i = 0; i < HALF; i++) {
                const float send = up ? v[i] : v[i + HALF];
                const float keep = up ? v[i + HALF] : v[i];
                v[i] = keep + __shfl_xor_sync(full, send, O);
            }
            if constexpr (O > 1) WarpSlotReduce<NP, HALF, O / 2>::run(v, lane);
        } else {
            v[0] += __shfl_xor_sync(full, v[0], O);
            if constexpr (O > 1) WarpSlotReduce<NP, 1, O / 2>::run(v, lane);
        }
    }
};
template <int NP>
__device__ __forceinline__ float warp_reduce_slots(float (&v)[NP], int lane) {
    WarpSlotReduce<NP, NP, 16>::run(v, lane);
    return v[0];
}

// 1/sqrt(x) to ~1 ulp (MUFU.RSQ + one Newton step): same quality as the reference's x / max(|x|, eps)
// (an IEEE sqrt followed by an IEEE divide) at a quarter of the instructions.
__device__ __forceinline__ float inv_len(float ss) {
    if (ss < 1e-24f) return 1e12f;   // F.normalize eps = 1e-12 on the norm
    const float y = rsqrtf(ss);
    return y * fmaf(-0.5f * ss * y, y, 1.5f);
}

// Raw per-sample inputs of one lane: loaded one (surfel, pass) ahead of their use so that the HBM
// latency of the next 32 samples hides behind the arithmetic of the current ones.
struct RawSample { float wx, wy, wz, r0, r1, r2, vis, area, t0, t1, t2; };

__device__ __forceinline__ void fetch_raw(const ShadeArgs& a, int n, int s0, int lane, RawSample& r) {
    const int s = s0 + lane;
    const size_t is = (size_t)n * a.Ns + (s < a.Ns ? s : a.Ns - 1);
    const float* d = a.dirs + is * 3;
    r.wx = __ldg(d); r.wy = __ldg(d + 1); r.wz = __ldg(d + 2);
    const float* rad = a.radiance + is * 3;
    r.r0 = __ldg(rad); r.r1 = __ldg(rad + 1); r.r2 = __ldg(rad + 2);
    r.vis = __ldg(a.visibility + is); r.area = __ldg(a.areas + is);
    if (a.taps) { const float* tp = a.taps + is * 3; r.t0 = __ldg(tp); r.t1 = __ldg(tp + 1); r.t2 = __ldg(tp + 2); }
    // nothing here may consume the loaded values: the point is to leave them in flight
}
// a padded lane (sample index >= Ns) contributes nothing
__device__ __forceinline__ RawSample mask_raw(const RawSample& r, bool ok) {
    RawSample o = r;
    if (!ok) { o.r0 = o.r1 = o.r2 = 0.f; o.vis = 0.f; o.area = 0.f; }
    return o;
}

// Raw per-surfel inputs: view direction (all lanes) and, on lanes 0..3, that vertex's normal/roughness.
struct RawSurfel { float vx, vy, vz, nx, ny, nz, rough, met, base; };

template <bool MET>
__device__ __forceinline__ void fetch_surfel(const ShadeArgs& a, int n, int lane, RawSurfel& r) {
    if (a.viewdirs) {
        const float* vd = a.viewdirs + (size_t)n * 3;
        r.vx = __ldg(vd); r.vy = __ldg(vd + 1); r.vz = __ldg(vd + 2);
    } else {   // svgss.py:95: viewdirs = normalize(camera_center - means3D); the kernels normalise
        const float* m = a.means3D + (size_t)n * 3;
        r.vx = __ldg(a.campos) - __ldg(m); r.vy = __ldg(a.campos + 1) - __ldg(m + 1); r.vz = __ldg(a.campos + 2) - __ldg(m + 2);
    }
    const int v = lane & 3;
    const float* nn = a.normals + ((size_t)n * 4 + v) * 3;
    r.nx = __ldg(nn); r.ny = __ldg(nn + 1); r.nz = __ldg(nn + 2);
    r.rough = __ldg(a.roughness + (size_t)n * 4 + v);
    r.met = MET ? __ldg(a.metallic + (size_t)n * 4 + v) : 0.f;
    r.base = __ldg(a.base_color + (size_t)n * 12 + (lane < 12 ? lane : 0));   // lane = 4*ch + v
}

// Per-sample, vertex-independent quantities (one lane = one light sample of the warp's surfel).
struct Sample {
    float wx, wy, wz, il;   // raw incident direction, 1/|w|
    float hx, hy, hz, ih;   // half vector (normalised), 1/|(L+V)/2|
    float voh_raw, p;       // V.H before the clamp, 2^((a1 VoH + a0) VoH)
    float ag[3], al[3];     // area * clamp(env)*vis , area * radiance
    float lg[3];            // clamp(env*scale) (without visibility)
    float raw[3];           // bilinear env value before scale/clamp
    EnvTap tap;
};

template <bool ENV4>   // ENV4: `env` is the padded float4 copy in shared memory
__device__ __forceinline__ void make_sample(const ShadeArgs& a, const float* env, const RawSample& r, float Vx,
                                            float Vy, float Vz, Sample& o) {
    o.wx = r.wx; o.wy = r.wy; o.wz = r.wz;
    o.il = inv_len(o.wx * o.wx + o.wy * o.wy + o.wz * o.wz);
    const float lx = o.wx * o.il, ly = o.wy * o.il, lz = o.wz * o.il;
    const float hx = (lx + Vx) * 0.5f, hy = (ly + Vy) * 0.5f, hz = (lz + Vz) * 0.5f;
    const float ih = inv_len(hx * hx + hy * hy + hz * hz);
    o.ih = ih;
    o.hx = hx * ih; o.hy = hy * ih; o.hz = hz * ih;
    o.voh_raw = Vx * o.hx + Vy * o.hy + Vz * o.hz;
    const float voh = fminf(fmaxf(o.voh_raw, 1e-6f), 1.f);
    o.p = exp2f((-5.55473f * voh - 6.98316f) * voh);
    if (a.taps) {
        o.tap = env_tap_unpack(r.t0, r.t1, r.t2);   // cached: the same env_coords result, computed once per run
    } else {
        float qx = o.wx, qy = o.wy, qz = o.wz;  // direct_light uses the raw direction
        if (a.transform) {
            const float* t = a.transform;  // dirs @ transform.T
            const float tx = qx * t[0] + qy * t[1] + qz * t[2];
            const float ty = qx * t[3] + qy * t[4] + qz * t[5];
            const float tz = qx * t[6] + qy * t[7] + qz * t[8];
            qx = tx; qy = ty; qz = tz;
        }
        o.tap = env_coords(qx, qy, qz, a.He, a.We);
    }
    if (ENV4) env_fetch4(reinterpret_cast<const float4*>(env), a.He, a.We, o.tap, o.raw);
    else env_fetch(env, a.He, a.We, o.tap, o.raw);
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        o.lg[ch] = fminf(fmaxf(o.raw[ch] * a.env_scale, 0.f), 64.f);
        o.ag[ch] = r.area * (o.lg[ch] * r.vis);
    }
    o.al[0] = r.area * r.r0; o.al[1] = r.area * r.r1; o.al[2] = r.area * r.r2;
}

// Per-(surfel, vertex) constants. c = sign(V.N^)/|N| so that N~.x = c (N.x) for the normalised,
// view-facing normal N~ of GGX_specular4 (svgss.py:603-607); N stays raw for n.w (svgss.py:548).
struct VertexConst {
    float Nx, Ny, Nz, c;
    float a2, k, nom1, NoV;
};

__device__ __forceinline__ void vertex_consts(const RawSurfel& s, float Vx, float Vy, float Vz, VertexConst& c) {
    c.Nx = s.nx; c.Ny = s.ny; c.Nz = s.nz;
    const float inv_nlen = inv_len(c.Nx * c.Nx + c.Ny * c.Ny + c.Nz * c.Nz);
    const float d = (c.Nx * inv_nlen) * Vx + (c.Ny * inv_nlen) * Vy + (c.Nz * inv_nlen) * Vz;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    c.c = sgn * inv_nlen;
    const float nv_raw = (c.Nx * Vx + c.Ny * Vy + c.Nz * Vz) * c.c;   // N~.V before the clamp
    c.NoV = fminf(fmaxf(nv_raw, 1e-6f), 1.f);
    const float al = s.rough * s.rough;
    c.a2 = al * al;
    c.k = (al + 2.f * s.rough + 1.0f) * 0.125f;
    c.nom1 = c.NoV * (1.f - c.k) + c.k;
}

struct ShadeOutK {
    float* pbr; float* diffuse; float* specular; float* direct; float* indirect;  // [N,row_stride] or null
    float* mean_vis;       // rows of mean_*stride floats, or null
    float* mean_local;
    float* mean_incident;
    float* mean_global;
    float* pack;           // optional: base12 | view-space normals 12 | roughness 4 (row_stride)
    float* sum_direct;     // [N,12] mean_s n.w*A_env (or of the total light when sum_indirect is null): saved for backward
    float* sum_indirect;   // [N,12] mean_s n.w*A_radiance, or null
    int row_stride, mean_vis_stride, mean_stride;
};

// One warp per surfel, one lane per light sample: every per-sample buffer is read with fully
// coalesced loads exactly once; per-vertex sums are combined with one transposed warp reduction.
template <bool SPLIT, bool MET, bool ENV_SMEM>
__global__ void __launch_bounds__(SH_THREADS) shade_fwd_kernel(const ShadeArgs a, const ShadeOutK out) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int WPC = SH_THREADS / 32;
    float* env_s = smem_f + WPC * 4 * VUF_FLOATS;
    const float* env = a.env_act;
    if (ENV_SMEM) {
        for (int i = threadIdx.x; i < a.He * a.We * 4; i += SH_THREADS) env_s[i] = (i & 3) < 3 ? a.env_act[(i >> 2) * 3 + (i & 3)] : 0.f;
        __syncthreads();
        env = env_s;
    }
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float* vu = smem_f + (threadIdx.x >> 5) * 4 * VUF_FLOATS;
    const int Ns = a.Ns;
    const float inv = 1.f / (float)Ns;
    const int stride = gridDim.x * WPC;
    const int count = a.list_count ? min(__ldg(a.list_count), a.N) : a.N;
    int slot = blockIdx.x * WPC + (threadIdx.x >> 5);
    int n = surfel_at(a, slot, count);
    if (n >= a.N) return;
    int n_next = surfel_at(a, slot + stride, count);

    RawSample raw;
    RawSurfel rs;
    fetch_surfel<MET>(a, n, lane, rs);
    fetch_raw(a, n, 0, lane, raw);
    while (n < a.N) {
        // ---- surfel prologue --------------------------------------------------------------------
        const float inv_vlen = inv_len(rs.vx * rs.vx + rs.vy * rs.vy + rs.vz * rs.vz);
        const float Vx = rs.vx * inv_vlen, Vy = rs.vy * inv_vlen, Vz = rs.vz * inv_vlen;
        // per-vertex uniforms, computed by lanes 0..3 and read back as broadcast LDS.128:
        //   [0..3] N, c | [4..7] a2, k, nom1, F0.r | [8..11] F0.g, F0.b, -, -
        const float base_g = MET ? __shfl_down_sync(full, rs.base, 4) : 0.f;   // lane v: base colour g, b of vertex v
        const float base_b = MET ? __shfl_down_sync(full, rs.base, 8) : 0.f;
        __syncwarp();
        if (lane < 4) {
            VertexConst c;
            vertex_consts(rs, Vx, Vy, Vz, c);
            float* u = vu + lane * VUF_FLOATS;
            u[0] = c.Nx; u[1] = c.Ny; u[2] = c.Nz; u[3] = c.c;
            u[4] = c.a2; u[5] = c.k; u[6] = c.nom1;
            if (MET) {
                const float m = rs.met;
                u[7] = 0.04f * (1.f - m) + rs.base * m;
                u[8] = 0.04f * (1.f - m) + base_g * m;
                u[9] = 0.04f * (1.f - m) + base_b * m;
            }
        }
        __syncwarp();
        const RawSurfel sc = rs;   // this surfel's per-lane raw values (normal/roughness of vertex lane&3, base of lane)
        const int n_next2 = surfel_at(a, slot + 2 * stride, count);   // work-list entry two ahead (the one-ahead surfel is fetched now)
        if (n_next < a.N) fetch_surfel<MET>(a, n_next, lane, rs);

        // D = sum ndi*A, S = sum f_s*ndi*A; [0]: env ("direct") light or the total when !SPLIT, [1]: cached radiance
        float D[SPLIT ? 2 : 1][4][3], S[SPLIT ? 2 : 1][4][3];
#pragma unroll
        for (int q = 0; q < (SPLIT ? 2 : 1); q++)
#pragma unroll
            for (int v = 0; v < 4; v++)
#pragma unroll
                for (int ch = 0; ch < 3; ch++) { D[q][v][ch] = 0.f; S[q][v][ch] = 0.f; }
        float m_vis = 0.f, m_g[3] = {0, 0, 0}, m_l[3] = {0, 0, 0};

        for (int s0 = 0; s0 < Ns; s0 += 32) {
            const RawSample cur = mask_raw(raw, s0 + lane < Ns);
            if (s0 + 32 < Ns) fetch_raw(a, n, s0 + 32, lane, raw);
            else if (n_next < a.N) fetch_raw(a, n_next, 0, lane, raw);
            Sample sm;
            make_sample<ENV_SMEM>(a, env, cur, Vx, Vy, Vz, sm);
            m_vis += cur.vis;
            m_l[0] += cur.r0; m_l[1] += cur.r1; m_l[2] += cur.r2;
            float A0[3], A1[3];
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                m_g[ch] = fmaf(sm.lg[ch], cur.vis, m_g[ch]);
                A0[ch] = SPLIT ? sm.ag[ch] : sm.ag[ch] + sm.al[ch];
                A1[ch] = sm.al[ch];
            }
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const float4 u0 = *reinterpret_cast<const float4*>(vu + v * VUF_FLOATS);
                const float4 u1 = *reinterpret_cast<const float4*>(vu + v * VUF_FLOATS + 4);
                float F0[3] = {0.04f, 0.04f, 0.04f};
                if (MET) {
                    const float2 u2 = *reinterpret_cast<const float2*>(vu + v * VUF_FLOATS + 8);
                    F0[0] = u1.w; F0[1] = u2.x; F0[2] = u2.y;
                }
                const float ndi_raw = u0.x * sm.wx + u0.y * sm.wy + u0.z * sm.wz;
                const float ndi = fmaxf(ndi_raw, 0.f);
                const float nh = u0.x * sm.hx + u0.y * sm.hy + u0.z * sm.hz;
                const float NoL = fminf(fmaxf(u0.w * sm.il * ndi_raw, 1e-6f), 1.f);
                const float NoH = fminf(fmaxf(u0.w * nh, 1e-6f), 1.f);
                const float nom0 = NoH * NoH * (u1.x - 1.f) + 1.f;
                const float nom2 = NoL * (1.f - u1.y) + u1.y;
                const float nom = fminf(fmaxf(4.f * PI_F * nom0 * nom0 * u1.z * nom2, 1e-6f), 4.f * PI_F);
                const float Dt = __fdividef(u1.x, nom);
                const float dn = Dt * ndi;
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    const float f0 = F0[ch];
                    const float fsn = (f0 + (1.f - f0) * sm.p) * dn;
                    D[0][v][ch] = fmaf(ndi, A0[ch], D[0][v][ch]);
                    S[0][v][ch] = fmaf(fsn, A0[ch], S[0][v][ch]);
                    if (SPLIT) {
                        D[1][v][ch] = fmaf(ndi, A1[ch], D[1][v][ch]);
                        S[1][v][ch] = fmaf(fsn, A1[ch], S[1][v][ch]);
                    }
                }
            }
        }
        // slots: 0..11 D (channel-major: 4*ch+v), 12..23 S, 24 vis, 25..27 local, 28..30 global
        float r0[32];
#pragma unroll
        for (int v = 0; v < 4; v++)
#pragma unroll
            for (int ch = 0; ch < 3; ch++) { r0[4 * ch + v] = D[0][v][ch]; r0[12 + 4 * ch + v] = S[0][v][ch]; }
        r0[24] = m_vis;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) { r0[25 + ch] = m_l[ch]; r0[28 + ch] = m_g[ch]; }
        r0[31] = 0.f;
        const float t0 = warp_reduce_slots<32>(r0, lane) * inv;
        float t1 = 0.f;
        if (SPLIT) {
            float r1[32];
#pragma unroll
            for (int i = 0; i < 32; i++) r1[i] = 0.f;
#pragma unroll
            for (int v = 0; v < 4; v++)
#pragma unroll
                for (int ch = 0; ch < 3; ch++) { r1[4 * ch + v] = D[1][v][ch]; r1[12 + 4 * ch + v] = S[1][v][ch]; }
            t1 = warp_reduce_slots<32>(r1, lane) * inv;
        }
        const float s0v = __shfl_down_sync(full, t0, 12);   // lanes 0..11: S of the same (ch,v)
        const float s1v = __shfl_down_sync(full, t1, 12);
        const float gl = __shfl_down_sync(full, t0, 3);     // lanes 25..27: global light of the same channel
        if (lane < 12) {
            (void)0;
            const size_t o = (size_t)n * out.row_stride + lane;
            const float base = sc.base;
            const float met = MET ? sc.met : 0.f;
            const float fd = (1.f - met) * base * (1.f / PI_F);
            if (out.sum_direct) out.sum_direct[(size_t)n * 12 + lane] = t0;
            if (!SPLIT) {
                if (out.diffuse) out.diffuse[o] = t0;
                if (out.specular) out.specular[o] = s0v;
                if (out.pbr) out.pbr[o] = fd * t0 + s0v;
            } else {
                const float dir = fd * t0 + s0v, ind = fd * t1 + s1v;
                if (out.sum_indirect) out.sum_indirect[(size_t)n * 12 + lane] = t1;
                if (out.diffuse) out.diffuse[o] = t0 + t1;
                if (out.specular) out.specular[o] = s0v + s1v;
                if (out.pbr) out.pbr[o] = fd * (t0 + t1) + (s0v + s1v);
                if (out.direct) out.direct[o] = dir;
                if (out.indirect) out.indirect[o] = ind;
            }
            if (out.pack) {  // render_view's pass-through columns (svgss.py:157-166)
                float* pk = out.pack + (size_t)n * out.row_stride;
                pk[lane] = base;
                const int j = lane >> 2;  // view-space axis; n_view[v][j] = sum_i N[v][i] R[i][j], R row-major [3,3]
                const int vs = a.view_stride;
                pk[12 + lane] = sc.nx * a.view3x3[j] + sc.ny * a.view3x3[vs + j] + sc.nz * a.view3x3[2 * vs + j];
                if (lane < 4) pk[24 + lane] = sc.rough;
            }
        } else if (lane == 24) {
            if (out.mean_vis) out.mean_vis[(size_t)n * out.mean_vis_stride] = t0;
        } else if (lane >= 25 && lane < 28) {
            const int ch = lane - 25;
            if (out.mean_local) out.mean_local[(size_t)n * out.mean_stride + ch] = t0;
            if (out.mean_incident) out.mean_incident[(size_t)n * out.mean_stride + ch] = t0 + gl;
        } else if (lane >= 28 && lane < 31) {
            if (out.mean_global) out.mean_global[(size_t)n * out.mean_stride + (lane - 28)] = t0;
        }
        n = n_next; n_next = n_next2; slot += stride;
    }
}

// ---------------------------------------------------------------------------------------------
struct ShadeGradsK {
    const float* g_pbr; const float* g_diffuse; const float* g_specular; const float* g_direct;
    const float* g_indirect;                     // rows of g_row_stride floats, or null
    const float* g_mean_vis;                     // rows of g_mean_*stride floats, or null
    const float* g_mean_local;
    const float* g_mean_incident;
    const float* g_mean_global;
    const float* g_pack;                         // optional grads of the pass-through columns (base|normal|roughness)
    const float* sum_direct;                     // [N,12] saved by the forward kernel
    const float* sum_indirect;                   // [N,12] or null (then sum_direct is the total)
    int g_row_stride, g_mean_vis_stride, g_mean_stride;
    const float* env_param;                      // raw parameter (learnable mode) for softplus'
    float* d_base_color; float* d_roughness; float* d_metallic; float* d_normals; float* d_viewdirs;
    float* d_radiance;                           // [N,Ns,3] or null
    float* d_visibility;                         // [N,Ns] or null
    float* d_env_acc;                            // [SVGIR_SHADE_ENV_COPIES][He,We,4] zeroed accumulators of the env gradient, or null
    float* d_means3D;                            // fused view directions: [N,3] += -(gradient of the raw view vector)
    int accumulate;                              // += into d_base_color / d_roughness / d_metallic / d_normals
};

__global__ void env_grad_finalize_kernel(int ntex, int copies, int env_mode, const float* __restrict__ acc,
                                         const float* __restrict__ env_param, float* __restrict__ d_env) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntex * 3) return;
    float val = 0.f;
    for (int c = 0; c < copies; c++) val += acc[((size_t)c * ntex + i / 3) * 4 + (i % 3)];
    if (env_mode == 0) val = val / (1.f + expf(-env_param[i]));  // softplus'
    d_env[i] += val;
}

// Upstream gradients and saved sums of one surfel, on lanes 0..11 (lane = 4*ch + v) plus the uniform
// mean gradients; fetched one surfel ahead like the other inputs.
struct RawGrad { float gp, gd, gs, gdi, gin, pk_base, sd, si, gmv, gml[3], gmg[3]; };

__device__ __forceinline__ void fetch_grad(const ShadeGradsK& g, int n, int lane, RawGrad& r) {
    const int l12 = lane < 12 ? lane : 0;
    const size_t og = (size_t)n * g.g_row_stride + l12;
    r.gp = g.g_pbr ? __ldg(g.g_pbr + og) : 0.f;
    r.gd = g.g_diffuse ? __ldg(g.g_diffuse + og) : 0.f;
    r.gs = g.g_specular ? __ldg(g.g_specular + og) : 0.f;
    r.gdi = g.g_direct ? __ldg(g.g_direct + og) : 0.f;
    r.gin = g.g_indirect ? __ldg(g.g_indirect + og) : 0.f;
    r.pk_base = g.g_pack ? __ldg(g.g_pack + og) : 0.f;
    r.sd = __ldg(g.sum_direct + (size_t)n * 12 + l12);
    r.si = g.sum_indirect ? __ldg(g.sum_indirect + (size_t)n * 12 + l12) : 0.f;
    const size_t om = (size_t)n * g.g_mean_stride;
    r.gmv = g.g_mean_vis ? __ldg(g.g_mean_vis + (size_t)n * g.g_mean_vis_stride) : 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float gi = g.g_mean_incident ? __ldg(g.g_mean_incident + om + ch) : 0.f;
        r.gml[ch] = (g.g_mean_local ? __ldg(g.g_mean_local + om + ch) : 0.f) + gi;
        r.gmg[ch] = (g.g_mean_global ? __ldg(g.g_mean_global + om + ch) : 0.f) + gi;
    }
}

template <bool MET, bool ENV_SMEM>
__global__ void __launch_bounds__(SHB_THREADS, SHB_MIN_CTAS) shade_bwd_kernel(const ShadeArgs a, const ShadeGradsK g) {
    extern __shared__ __align__(16) float smem_b[];
    constexpr int WPC = SHB_THREADS / 32;
    constexpr int NACC = MET ? 10 : 7;   // per-vertex partial sums kept by every lane
    if (a.skip_flag && __ldg(a.skip_flag) != 0) return;   // the step's binning overflowed: contribute nothing
    const int nenv = a.He * a.We * 3;
    // same-address reductions serialise in L2: each CTA scatters into one of SVGIR_SHADE_ENV_COPIES replicas of the accumulator
    float* env_copy = g.d_env_acc ? g.d_env_acc + (size_t)(blockIdx.x % SVGIR_SHADE_ENV_COPIES) * a.He * a.We * 4 : nullptr;
    float* vu_all = smem_b;                                         // [WPC][4][VU_FLOATS]
    float* env_s = smem_b + WPC * 4 * VU_FLOATS;                    // activated env (if it fits)
    const float* env = a.env_act;
    if (ENV_SMEM) {
        for (int i = threadIdx.x; i < nenv; i += SHB_THREADS) env_s[i] = a.env_act[i];   // scalar layout: the padded float4 lookup costs this kernel registers it does not have (0.508 -> 0.526 ms)
        __syncthreads();
        env = env_s;
    }
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float* vu = vu_all + (threadIdx.x >> 5) * 4 * VU_FLOATS;
    const int Ns = a.Ns;
    const float inv = 1.f / (float)Ns;
    const bool want_rad = g.d_radiance != nullptr;
    const int stride = gridDim.x * WPC;
    const int count = a.list_count ? min(__ldg(a.list_count), a.N) : a.N;
    int slot = blockIdx.x * WPC + (threadIdx.x >> 5);
    int n = surfel_at(a, slot, count);
    int n_next = surfel_at(a, slot + stride, count);

    RawSample raw;
    RawSurfel rs;
    RawGrad rg;
    if (n < a.N) {
        fetch_surfel<MET>(a, n, lane, rs);
        fetch_grad(g, n, lane, rg);
        fetch_raw(a, n, 0, lane, raw);
    }
    while (n < a.N) {
        const float inv_vlen = inv_len(rs.vx * rs.vx + rs.vy * rs.vy + rs.vz * rs.vz);
        const float Vx = rs.vx * inv_vlen, Vy = rs.vy * inv_vlen, Vz = rs.vz * inv_vlen;
        // ---- per-vertex uniforms -> shared: [0..3] N,c | [4..7] a2,k,nom1,NoV | [8..10] gDg,[11] F0r |
        //      [12..14] gDl,[15] F0g | [16..18] gSg,[19] F0b | [20..22] gSl,[23] roughness
        __syncwarp();
        if (lane < 4) {
            VertexConst c;
            vertex_consts(rs, Vx, Vy, Vz, c);
            float* u = vu + lane * VU_FLOATS;
            u[0] = c.Nx; u[1] = c.Ny; u[2] = c.Nz; u[3] = c.c;
            u[4] = c.a2; u[5] = c.k; u[6] = c.nom1; u[7] = c.NoV;
            u[23] = rs.rough;
        }
        const RawSurfel sc = rs;
        const RawGrad gc = rg;
        if (lane < 12) {
            const int v = lane & 3, ch = lane >> 2;
            const float met_l = MET ? sc.met : 0.f;
            const float fd = (1.f - met_l) * sc.base * (1.f / PI_F);
            float* u = vu + v * VU_FLOATS;
            u[8 + ch] = (gc.gd + (gc.gp + gc.gdi) * fd) * inv;
            u[12 + ch] = (gc.gd + (gc.gp + gc.gin) * fd) * inv;
            u[16 + ch] = (gc.gs + gc.gp + gc.gdi) * inv;
            u[20 + ch] = (gc.gs + gc.gp + gc.gin) * inv;
            u[11 + 4 * ch] = MET ? 0.04f * (1.f - met_l) + sc.base * met_l : 0.04f;
        }
        __syncwarp();
        const int n_next2 = surfel_at(a, slot + 2 * stride, count);
        if (n_next < a.N) {
            fetch_surfel<MET>(a, n_next, lane, rs);
            fetch_grad(g, n_next, lane, rg);
        }
        const float gmv = gc.gmv * inv;
        const float gml[3] = {gc.gml[0] * inv, gc.gml[1] * inv, gc.gml[2] * inv};
        const float gmg[3] = {gc.gmg[0] * inv, gc.gmg[1] * inv, gc.gmg[2] * inv};

        // per-vertex partial sums of this lane: 0..2 dN, 3 d_c, 4 d_a2, 5 d_k, 6 d_nom1, 7..9 dF0 (metallic only)
        float acc[4][NACC];
#pragma unroll
        for (int v = 0; v < 4; v++)
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[v][i] = 0.f;
        float dV[3] = {0, 0, 0};

        for (int s0 = 0; s0 < Ns; s0 += 32) {
            const RawSample cur = mask_raw(raw, s0 + lane < Ns);
            if (s0 + 32 < Ns) fetch_raw(a, n, s0 + 32, lane, raw);
            else if (n_next < a.N) fetch_raw(a, n_next, 0, lane, raw);
            const int s = s0 + lane;
            const bool ok = s < Ns;
            const size_t is = (size_t)n * Ns + (ok ? s : Ns - 1);
            Sample sm;
            make_sample<false>(a, env, cur, Vx, Vy, Vz, sm);
            float dAg[3] = {0, 0, 0}, dAl[3] = {0, 0, 0}, dp_acc = 0.f, dhN[3] = {0, 0, 0};
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const float4 u0 = *reinterpret_cast<const float4*>(vu + v * VU_FLOATS);
                const float4 u1 = *reinterpret_cast<const float4*>(vu + v * VU_FLOATS + 4);
                const float4 uDg = *reinterpret_cast<const float4*>(vu + v * VU_FLOATS + 8);
                const float4 uDl = *reinterpret_cast<const float4*>(vu + v * VU_FLOATS + 12);
                const float4 uSg = *reinterpret_cast<const float4*>(vu + v * VU_FLOATS + 16);
                const float4 uSl = *reinterpret_cast<const float4*>(vu + v * VU_FLOATS + 20);
                const float Nx = u0.x, Ny = u0.y, Nz = u0.z, cc = u0.w;
                const float a2 = u1.x, k = u1.y, nom1 = u1.z, NoV = u1.w;
                const float gDg[3] = {uDg.x, uDg.y, uDg.z}, gDl[3] = {uDl.x, uDl.y, uDl.z};
                const float gSg[3] = {uSg.x, uSg.y, uSg.z}, gSl[3] = {uSl.x, uSl.y, uSl.z};
                const float F0[3] = {uDg.w, uDl.w, uSg.w};

                const float ndi_raw = Nx * sm.wx + Ny * sm.wy + Nz * sm.wz;
                const float ndi = fmaxf(ndi_raw, 0.f);
                const float nh = Nx * sm.hx + Ny * sm.hy + Nz * sm.hz;
                const float nol_raw = cc * sm.il * ndi_raw, noh_raw = cc * nh;
                const float NoL = fminf(fmaxf(nol_raw, 1e-6f), 1.f);
                const float NoH = fminf(fmaxf(noh_raw, 1e-6f), 1.f);
                const float nom0 = NoH * NoH * (a2 - 1.f) + 1.f;
                const float nom2 = NoL * (1.f - k) + k;
                const float n02 = nom0 * nom0;
                const float nom_raw = 4.f * PI_F * n02 * nom1 * nom2;
                const float nom = fminf(fmaxf(nom_raw, 1e-6f), 4.f * PI_F);
                const float rn = __fdividef(1.f, nom);
                const float Dt = a2 * rn;
                const float dn = Dt * ndi;

                float qD = 0.f, Fq = 0.f, Gq = 0.f;   // sum gD*A ; sum F[ch]*qs[ch] ; sum (1-F0[ch])*qs[ch]
                float* ac = acc[v];
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    const float qs = gSg[ch] * sm.ag[ch] + gSl[ch] * sm.al[ch];
                    qD = fmaf(gDg[ch], sm.ag[ch], fmaf(gDl[ch], sm.al[ch], qD));
                    float F;
                    if (MET) {
                        F = F0[ch] + (1.f - F0[ch]) * sm.p;
                        Fq = fmaf(F, qs, Fq);
                        Gq = fmaf(1.f - F0[ch], qs, Gq);
                        ac[7 + ch] = fmaf(dn * (1.f - sm.p), qs, ac[7 + ch]);
                    } else {
                        F = 0.04f + 0.96f * sm.p;
                        Fq += qs;     // scaled by F / 0.96 after the loop
                    }
                    const float fsn = F * dn;
                    dAg[ch] += ndi * gDg[ch] + fsn * gSg[ch];
                    dAl[ch] += ndi * gDl[ch] + fsn * gSl[ch];
                }
                if (!MET) { Gq = 0.96f * Fq; Fq *= 0.04f + 0.96f * sm.p; }
                const float d_ndi = qD + Dt * Fq;
                const float d_Dt = ndi * Fq;
                dp_acc = fmaf(dn, Gq, dp_acc);
                // Dt = a2 / clamp(nom_raw)
                const bool nom_in = nom_raw >= 1e-6f && nom_raw <= 4.f * PI_F;
                float d_a2 = d_Dt * rn;
                const float c4 = nom_in ? -4.f * PI_F * d_Dt * Dt * rn : 0.f;
                const float d_nom0 = c4 * 2.f * nom0 * nom1 * nom2;
                const float g1 = c4 * n02 * nom2;
                const float d_nom2 = c4 * n02 * nom1;
                const bool noh_in = noh_raw >= 1e-6f && noh_raw <= 1.f;
                const bool nol_in = nol_raw >= 1e-6f && nol_raw <= 1.f;
                const float d_noh = noh_in ? d_nom0 * 2.f * NoH * (a2 - 1.f) : 0.f;
                d_a2 = fmaf(d_nom0, NoH * NoH, d_a2);
                const float d_nol = nol_in ? d_nom2 * (1.f - k) : 0.f;
                const float g_ndi_raw = (ndi_raw >= 0.f ? d_ndi : 0.f) + d_nol * cc * sm.il;
                const float g_nh = d_noh * cc;
                ac[0] += g_ndi_raw * sm.wx + g_nh * sm.hx;
                ac[1] += g_ndi_raw * sm.wy + g_nh * sm.hy;
                ac[2] += g_ndi_raw * sm.wz + g_nh * sm.hz;
                ac[3] += d_nol * sm.il * ndi_raw + d_noh * nh;
                ac[4] += d_a2;
                ac[5] += g1 * (1.f - NoV) + d_nom2 * (1.f - NoL);
                ac[6] += g1;
                dhN[0] = fmaf(g_nh, Nx, dhN[0]);
                dhN[1] = fmaf(g_nh, Ny, dhN[1]);
                dhN[2] = fmaf(g_nh, Nz, dhN[2]);
            }
            // ---- per-sample tail: V.H and H -> view direction; light gradients ------------------
            {
                const float voh = fminf(fmaxf(sm.voh_raw, 1e-6f), 1.f);
                const bool voh_in = sm.voh_raw >= 1e-6f && sm.voh_raw <= 1.f;
                const float d_voh = voh_in ? dp_acc * sm.p * 0.6931471805599453f * (2.f * -5.55473f * voh - 6.98316f) : 0.f;
                const float dH[3] = {dhN[0] + d_voh * Vx, dhN[1] + d_voh * Vy, dhN[2] + d_voh * Vz};
                const float hd = sm.hx * dH[0] + sm.hy * dH[1] + sm.hz * dH[2];
                const float ih = 0.5f * sm.ih;
                dV[0] += d_voh * sm.hx + (dH[0] - sm.hx * hd) * ih;
                dV[1] += d_voh * sm.hy + (dH[1] - sm.hy * hd) * ih;
                dV[2] += d_voh * sm.hz + (dH[2] - sm.hz * hd) * ih;
            }
            if (want_rad && ok) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++) g.d_radiance[is * 3 + ch] = dAl[ch] * cur.area + gml[ch];
            }
            float dvis = gmv, draw[3];
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                const float x = sm.raw[ch] * a.env_scale;
                const float dLg = dAg[ch] * cur.area + gmg[ch];
                dvis = fmaf(dLg, sm.lg[ch], dvis);
                draw[ch] = (ok && x >= 0.f && x <= 64.f) ? dLg * cur.vis * a.env_scale : 0.f;
            }
            if (g.d_visibility && ok) g.d_visibility[is] = dvis;
            if (g.d_env_acc && (draw[0] != 0.f || draw[1] != 0.f || draw[2] != 0.f)) {
                const float wx0 = 1.f - sm.tap.wx1, wy0 = 1.f - sm.tap.wy1;
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    const int x = sm.tap.x0 + (kk & 1), y = sm.tap.y0 + (kk >> 1);
                    if (x < 0 || x > a.We - 1 || y < 0 || y > a.He - 1) continue;
                    const float w = ((kk & 1) ? sm.tap.wx1 : wx0) * ((kk >> 1) ? sm.tap.wy1 : wy0);
                    red_add_v4(env_copy + (size_t)(y * a.We + x) * 4, draw[0] * w, draw[1] * w, draw[2] * w);
                }
            }
        }

        // ---- per-vertex epilogue ----------------------------------------------------------------------
        // One 32-slot transposed reduction for all four vertices: afterwards lane L holds the total of
        // slot (L & 7) of vertex (L >> 3); the metallic-only dF0 sums take a second, 16-slot reduction.
        float r[32];
#pragma unroll
        for (int v = 0; v < 4; v++) {
#pragma unroll
            for (int i = 0; i < 7; i++) r[8 * v + i] = acc[v][i];
            r[8 * v + 7] = 0.f;
        }
        const float tot = warp_reduce_slots<32>(r, lane);
        float totF = 0.f;   // lanes 2*(4*ch+v), +1: dF0[v][ch]
        if (MET) {
            float rf[16];
#pragma unroll
            for (int i = 0; i < 16; i++) rf[i] = 0.f;
#pragma unroll
            for (int v = 0; v < 4; v++)
#pragma unroll
                for (int ch = 0; ch < 3; ch++) rf[4 * ch + v] = acc[v][MET ? 7 + ch : 0];
            totF = warp_reduce_slots<16>(rf, lane);
        }
        {
            const int hb = lane & 24, slot = lane & 7, v = lane >> 3;
            const float* u = vu + v * VU_FLOATS;
            const float4 u0 = *reinterpret_cast<const float4*>(u);
            const float Nx = u0.x, Ny = u0.y, Nz = u0.z, cc = u0.w;
            const float k = u[5], rgh = u[23];
            const float d_c = __shfl_sync(full, tot, hb + 3), d_a2 = __shfl_sync(full, tot, hb + 4);
            const float d_k = __shfl_sync(full, tot, hb + 5), d_nom1 = __shfl_sync(full, tot, hb + 6);
            const float nvr = Nx * Vx + Ny * Vy + Nz * Vz;
            const float nv_raw = nvr * cc;
            const float d_nov = (nv_raw >= 1e-6f && nv_raw <= 1.f) ? d_nom1 * (1.f - k) : 0.f;
            float dvu = 0.f;
            if (slot < 3) {
                const float Nk = slot == 0 ? Nx : (slot == 1 ? Ny : Nz);
                const float Vk = slot == 0 ? Vx : (slot == 1 ? Vy : Vz);
                const float nl2 = fmaxf(Nx * Nx + Ny * Ny + Nz * Nz, 1e-24f);
                // c = sgn/|N|  ->  dc/dN = -c N/|N|^2
                float val = tot + d_nov * cc * Vk - (d_c + d_nov * nvr) * cc * Nk / nl2;
                if (g.g_pack) {  // packed view-space normals: n_view[v][j] = sum_i N[v][i] R[i][j]
                    const float* gp = g.g_pack + (size_t)n * g.g_row_stride + 12;
                    const float* vr = a.view3x3 + a.view_stride * slot;
                    val += gp[v] * vr[0] + gp[4 + v] * vr[1] + gp[8 + v] * vr[2];
                }
                float* dst = g.d_normals + ((size_t)n * 4 + v) * 3 + slot;
                *dst = g.accumulate ? *dst + val : val;
                dvu = d_nov * cc * Nk;     // vertex-level N~.V term of dV component `slot`
            } else if (slot == 4) {
                float val = d_a2 * 4.f * rgh * rgh * rgh + d_k * (2.f * rgh + 2.f) * 0.125f;
                if (g.g_pack) val += g.g_pack[(size_t)n * g.g_row_stride + 24 + v];
                float* dst = g.d_roughness + (size_t)n * 4 + v;
                *dst = g.accumulate ? *dst + val : val;
            }
            // fold the vertex-level view terms into the per-sample ones: component = slot (0..2)
            dV[0] += slot == 0 ? dvu : 0.f;
            dV[1] += slot == 1 ? dvu : 0.f;
            dV[2] += slot == 2 ? dvu : 0.f;
        }
        // base colour / metallic through f_d = (1-m) base/pi and F0 = 0.04(1-m) + base m; lane = 4*ch + v
        {
            const float dF0 = __shfl_sync(full, totF, (2 * lane) & 31);
            float dm = 0.f;
            if (lane < 12) {
                const float base = sc.base;
                const float met = MET ? sc.met : 0.f;
                const float dfd = g.sum_indirect ? gc.gp * (gc.sd + gc.si) + gc.gdi * gc.sd + gc.gin * gc.si : gc.gp * gc.sd;
                float db = dfd * (1.f - met) * (1.f / PI_F);
                dm = dfd * (-base * (1.f / PI_F));
                if (MET) { db += dF0 * met; dm += dF0 * (base - 0.04f); }   // dF0 already carries the 1/Ns of the upstream
                db += gc.pk_base;
                float* dst = g.d_base_color + (size_t)n * 12 + lane;
                *dst = g.accumulate ? *dst + db : db;
            }
            if (MET && g.d_metallic) {  // uniform branch: all lanes shuffle; lanes v, 4+v, 8+v hold the 3 channels
                const float t = dm + __shfl_down_sync(full, dm, 4) + __shfl_down_sync(full, dm, 8);
                if (lane < 4) {
                    float* dst = g.d_metallic + (size_t)n * 4 + lane;
                    *dst = g.accumulate ? *dst + t : t;
                }
            }
        }
        // view direction: summed over the warp, then through the normalisation V = vd/|vd|
#pragma unroll
        for (int ch = 0; ch < 3; ch++) dV[ch] = warp_sum(dV[ch]);
        if (lane < 3) {
            const float vdot = Vx * dV[0] + Vy * dV[1] + Vz * dV[2];
            const float Vk = lane == 0 ? Vx : (lane == 1 ? Vy : Vz);
            const float dk = lane == 0 ? dV[0] : (lane == 1 ? dV[1] : dV[2]);
            const float val = (dk - Vk * vdot) * inv_vlen;   // gradient of the un-normalised view vector
            if (a.viewdirs) g.d_viewdirs[(size_t)n * 3 + lane] = val;
            else atomicAdd(g.d_means3D + (size_t)n * 3 + lane, -val);   // view vector = campos - means3D
        }
        n = n_next; n_next = n_next2; slot += stride;
    }
}

// ---- env tap cache --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) env_taps_kernel(long long n, int He, int We, const float* __restrict__ transform,
                                                       const float* __restrict__ dirs, float* __restrict__ taps) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
    if (transform) {
        const float tx = x * transform[0] + y * transform[1] + z * transform[2];
        const float ty = x * transform[3] + y * transform[4] + z * transform[5];
        const float tz = x * transform[6] + y * transform[7] + z * transform[8];
        x = tx; y = ty; z = tz;
    }
    const EnvTap t = env_coords(x, y, z, He, We);
    taps[3 * i] = env_tap_pack(t); taps[3 * i + 1] = t.wx1; taps[3 * i + 2] = t.wy1;
}

// ---- stand-alone env lookup (DirectLightMap.direct_light / EnvLight.direct_light) ---------------
__global__ void __launch_bounds__(256) direct_light_fwd_kernel(int n, int He, int We, float scale,
                                                               const float* __restrict__ env_act,
                                                               const float* __restrict__ transform,
                                                               const float* __restrict__ dirs,
                                                               float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = dirs[3 * (size_t)i], y = dirs[3 * (size_t)i + 1], z = dirs[3 * (size_t)i + 2];
    if (transform) {
        const float tx = x * transform[0] + y * transform[1] + z * transform[2];
        const float ty = x * transform[3] + y * transform[4] + z * transform[5];
        const float tz = x * transform[6] + y * transform[7] + z * transform[8];
        x = tx; y = ty; z = tz;
    }
    const EnvTap t = env_coords(x, y, z, He, We);
    float o[3];
    env_fetch(env_act, He, We, t, o);
    out[3 * (size_t)i] = o[0] * scale; out[3 * (size_t)i + 1] = o[1] * scale; out[3 * (size_t)i + 2] = o[2] * scale;
}

__global__ void __launch_bounds__(256) direct_light_bwd_kernel(int n, int He, int We, float scale, int env_mode,
                                                               const float* __restrict__ env_param,
                                                               const float* __restrict__ transform,
                                                               const float* __restrict__ dirs,
                                                               const float* __restrict__ g_out,
                                                               float* __restrict__ d_env) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = dirs[3 * (size_t)i], y = dirs[3 * (size_t)i + 1], z = dirs[3 * (size_t)i + 2];
    if (transform) {
        const float tx = x * transform[0] + y * transform[1] + z * transform[2];
        const float ty = x * transform[3] + y * transform[4] + z * transform[5];
        const float tz = x * transform[6] + y * transform[7] + z * transform[8];
        x = tx; y = ty; z = tz;
    }
    const EnvTap t = env_coords(x, y, z, He, We);
    const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
        if (xx < 0 || xx > We - 1 || yy < 0 || yy > He - 1) continue;
        const float w = ((k & 1) ? t.wx1 : wx0) * ((k >> 1) ? t.wy1 : wy0) * scale;
        const int b = (yy * We + xx) * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            float val = g_out[3 * (size_t)i + ch] * w;
            if (env_mode == 0) val = val / (1.f + expf(-env_param[b + ch]));
            if (val != 0.f) atomicAdd(&d_env[b + ch], val);
        }
    }
}

void launch_env_activate(int nenv, const float* param, float* act, int env_mode, cudaStream_t s) {
    env_activate_kernel<<<(nenv + 255) / 256, 256, 0, s>>>(nenv, param, act, env_mode);
}
void launch_env_grad_finalize(int ntex, int copies, int env_mode, const float* acc, const float* env_param, float* d_env, cudaStream_t s) {
    env_grad_finalize_kernel<<<(ntex * 3 + 255) / 256, 256, 0, s>>>(ntex, copies, env_mode, acc, env_param, d_env);
}

}  // namespace svgir

using namespace svgir;

// persistent grid: enough CTAs to fill the 148 SMs a few times over, never more than the work
// svgir_shade_reserve_sms(n): the persistent grids leave n SMs free for a kernel running beside them (the
// peer-memory gradient all-reduce of the rasteriser-side segment, which is resident before the shading backward
// starts); with a full-width grid the displaced CTAs would run as a second wave and double the kernel's time.
static int g_reserved_sms = 0;
static int shade_grid(int N, int threads, int ctas_per_sm) {
    const int wpc = threads / 32;
    const int need = (N + wpc - 1) / wpc;
    const int sms = 148 - g_reserved_sms;
    return need < sms * ctas_per_sm ? need : sms * ctas_per_sm;
}

template <bool SPLIT, bool MET>
static void launch_shade_fwd(const ShadeArgs& a, const ShadeOutK& so, cudaStream_t s) {
    const size_t env_bytes = (size_t)a.He * a.We * 4 * sizeof(float);   // padded float4 texels
    const size_t vu_bytes = (size_t)(SH_THREADS / 32) * 4 * VUF_FLOATS * sizeof(float);
    const int grid = shade_grid(a.N, SH_THREADS, 2);
    if (env_bytes + vu_bytes <= 96 * 1024) {
        if (env_bytes + vu_bytes > 48 * 1024)
            cudaFuncSetAttribute(shade_fwd_kernel<SPLIT, MET, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(env_bytes + vu_bytes));
        shade_fwd_kernel<SPLIT, MET, true><<<grid, SH_THREADS, env_bytes + vu_bytes, s>>>(a, so);
    } else {
        shade_fwd_kernel<SPLIT, MET, false><<<grid, SH_THREADS, vu_bytes, s>>>(a, so);
    }
}

template <bool MET>
static void launch_shade_bwd(const ShadeArgs& a, const ShadeGradsK& g, cudaStream_t s) {
    const size_t env_bytes = (size_t)a.He * a.We * 3 * sizeof(float);
    const size_t vu_bytes = (size_t)(SHB_THREADS / 32) * 4 * VU_FLOATS * sizeof(float);
    const int grid = shade_grid(a.N, SHB_THREADS, SHB_MIN_CTAS);
    if (env_bytes + vu_bytes <= 48 * 1024) {
        shade_bwd_kernel<MET, true><<<grid, SHB_THREADS, env_bytes + vu_bytes, s>>>(a, g);
    } else {
        shade_bwd_kernel<MET, false><<<grid, SHB_THREADS, vu_bytes, s>>>(a, g);
    }
}

extern "C" {

void svgir_shade_reserve_sms(int n) { g_reserved_sms = n < 0 ? 0 : (n > 100 ? 100 : n); }

static int shade_prepare(const svgir_shade_cfg* c, const svgir_shade_in* in, ShadeArgs& a, cudaStream_t s) {
    if (!c || !in || c->N < 0 || c->Ns <= 0 || c->env_h <= 0 || c->env_w <= 0) { set_error("shade: bad cfg"); return SVGIR_ERR_INVALID; }
    if (!in->base_color || !in->roughness || !in->normals || !in->radiance || !in->visibility ||
        !in->incident_dirs || !in->incident_areas || !in->env || !in->env_act_scratch) {
        set_error("shade: missing input");
        return SVGIR_ERR_INVALID;
    }
    if (!in->viewdirs && !(in->means3D && in->campos)) { set_error("shade: provide viewdirs, or means3D + campos"); return SVGIR_ERR_INVALID; }
    const int nenv = c->env_h * c->env_w * 3;
    if (!(c->flags & SVGIR_SHADE_ENV_READY))
        { TimedScope ts_("env_activate", s); env_activate_kernel<<<(nenv + 255) / 256, 256, 0, s>>>(nenv, in->env, in->env_act_scratch, c->env_mode); }
    a.N = c->N; a.Ns = c->Ns; a.He = c->env_h; a.We = c->env_w;
    a.env_scale = c->env_mode == 0 ? 2.0f : 1.0f;
    a.env_act = in->env_act_scratch; a.transform = in->env_transform; a.view3x3 = in->view3x3;
    a.base_color = in->base_color; a.roughness = in->roughness; a.metallic = in->metallic;
    a.normals = in->normals; a.viewdirs = in->viewdirs; a.radiance = in->radiance;
    a.visibility = in->visibility; a.dirs = in->incident_dirs; a.areas = in->incident_areas;
    a.list = in->surfel_list; a.list_count = in->surfel_list ? in->surfel_count : nullptr;
    a.means3D = in->means3D; a.campos = in->campos; a.skip_flag = in->skip_flag; a.taps = in->env_taps;
    a.view_stride = (c->flags & SVGIR_SHADE_VIEW_4X4) ? 4 : 3;
    if (in->surfel_list && !in->surfel_count) { set_error("shade: surfel_list needs surfel_count"); return SVGIR_ERR_INVALID; }
    return SVGIR_OK;
}

int svgir_shade_forward(const svgir_shade_cfg* c, const svgir_shade_in* in, const svgir_shade_out* o, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    ShadeArgs a;
    int rc = shade_prepare(c, in, a, s);
    if (rc) return rc;
    if (!o || !(o->pbr || o->diffuse_light || o->specular || o->direct || o->indirect)) { set_error("shade: no output requested"); return SVGIR_ERR_INVALID; }
    if (o->pack && !in->view3x3) { set_error("shade: packed pass-through columns need view3x3"); return SVGIR_ERR_INVALID; }
    if (c->N == 0) return SVGIR_OK;
    ShadeOutK so{o->pbr, o->diffuse_light, o->specular, o->direct, o->indirect, o->mean_visibility,
                 o->mean_local, o->mean_incident, o->mean_global, o->pack, o->sum_direct, o->sum_indirect,
                 o->row_stride > 0 ? o->row_stride : 12, o->mean_vis_stride > 0 ? o->mean_vis_stride : 1,
                 o->mean_stride > 0 ? o->mean_stride : 3};
    const bool split = o->direct || o->indirect || o->sum_indirect;
    const bool met = in->metallic != nullptr;
    {
        TimedScope ts_("shade_fwd", s);
        if (split) { if (met) launch_shade_fwd<true, true>(a, so, s); else launch_shade_fwd<true, false>(a, so, s); }
        else { if (met) launch_shade_fwd<false, true>(a, so, s); else launch_shade_fwd<false, false>(a, so, s); }
    }
    return check_launch("shade_forward", c->debug, s);
}

int svgir_shade_backward(const svgir_shade_cfg* c, const svgir_shade_in* in, const svgir_shade_grads* gr, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    ShadeArgs a;
    int rc = shade_prepare(c, in, a, s);
    if (rc) return rc;
    if (!gr || !gr->d_base_color || !gr->d_roughness || !gr->d_normals || !(in->viewdirs ? gr->d_viewdirs : gr->d_means3D)) {
        set_error("shade_backward: missing buffers (d_viewdirs with viewdirs, d_means3D with means3D + campos)");
        return SVGIR_ERR_INVALID;
    }
    if (gr->g_pack && !in->view3x3) { set_error("shade_backward: packed pass-through columns need view3x3"); return SVGIR_ERR_INVALID; }
    if (!gr->sum_direct) { set_error("shade_backward: sum_direct (saved by svgir_shade_forward) is required"); return SVGIR_ERR_INVALID; }
    if ((gr->g_direct || gr->g_indirect) && !gr->sum_indirect) { set_error("shade_backward: g_direct/g_indirect need the split sums (sum_indirect)"); return SVGIR_ERR_INVALID; }
    if (c->N == 0) return SVGIR_OK;
    ShadeGradsK g{gr->g_pbr, gr->g_diffuse_light, gr->g_specular, gr->g_direct, gr->g_indirect, gr->g_mean_visibility,
                  gr->g_mean_local, gr->g_mean_incident, gr->g_mean_global, gr->g_pack, gr->sum_direct, gr->sum_indirect,
                  gr->g_row_stride > 0 ? gr->g_row_stride : 12, gr->g_mean_vis_stride > 0 ? gr->g_mean_vis_stride : 1,
                  gr->g_mean_stride > 0 ? gr->g_mean_stride : 3, in->env,
                  gr->d_base_color, gr->d_roughness, gr->d_metallic, gr->d_normals, gr->d_viewdirs,
                  gr->d_radiance, gr->d_visibility, gr->d_env ? gr->d_env_scratch : nullptr,
                  gr->d_means3D, (c->flags & SVGIR_SHADE_ACCUMULATE) ? 1 : 0};
    const int ntex = a.He * a.We;
    if (gr->d_env) {
        if (!gr->d_env_scratch) { set_error("shade_backward: d_env needs d_env_scratch [SVGIR_SHADE_ENV_COPIES*env_h*env_w*4]"); return SVGIR_ERR_INVALID; }
        if (cudaMemsetAsync(gr->d_env_scratch, 0, (size_t)SVGIR_SHADE_ENV_COPIES * ntex * 4 * sizeof(float), s) != cudaSuccess) { set_error("memset failed"); return SVGIR_ERR_CUDA; }
    }
    {
        TimedScope ts_("shade_bwd", s);
        if (in->metallic) launch_shade_bwd<true>(a, g, s); else launch_shade_bwd<false>(a, g, s);
    }
    if (gr->d_env) {
        TimedScope ts_("env_grad_finalize", s);
        env_grad_finalize_kernel<<<(ntex * 3 + 255) / 256, 256, 0, s>>>(ntex, SVGIR_SHADE_ENV_COPIES, c->env_mode, gr->d_env_scratch, in->env, gr->d_env);
    }
    return check_launch("shade_backward", c->debug, s);
}

int svgir_env_taps(long long n, int env_h, int env_w, const float* transform, const float* dirs, float* taps, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n < 0 || env_h <= 0 || env_w <= 0 || env_h > 32767 || env_w > 32767 || (n > 0 && (!dirs || !taps))) { set_error("env_taps: bad args"); return SVGIR_ERR_INVALID; }
    if (n == 0) return SVGIR_OK;
    const long long blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffLL) { set_error("env_taps: too many directions for one launch"); return SVGIR_ERR_INVALID; }
    { TimedScope ts_("env_taps", s); env_taps_kernel<<<(unsigned)blocks, 256, 0, s>>>(n, env_h, env_w, transform, dirs, taps); }
    return check_launch("env_taps", false, s);
}

int svgir_direct_light_forward(int n, int env_h, int env_w, int env_mode, const float* env, float* env_act_scratch,
                               const float* transform, const float* dirs, float* out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n < 0 || env_h <= 0 || env_w <= 0 || !env || !env_act_scratch || (n > 0 && (!dirs || !out))) { set_error("direct_light: bad args"); return SVGIR_ERR_INVALID; }
    const int nenv = env_h * env_w * 3;
    env_activate_kernel<<<(nenv + 255) / 256, 256, 0, s>>>(nenv, env, env_act_scratch, env_mode);
    if (n > 0)
        { TimedScope ts_("direct_light_fwd", s); direct_light_fwd_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, env_h, env_w, env_mode == 0 ? 2.0f : 1.0f,
                                                               env_act_scratch, transform, dirs, out); }
    return check_launch("direct_light_forward", false, s);
}

int svgir_direct_light_backward(int n, int env_h, int env_w, int env_mode, const float* env, const float* transform,
                                const float* dirs, const float* g_out, float* d_env, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n < 0 || env_h <= 0 || env_w <= 0 || !env || !d_env || (n > 0 && (!dirs || !g_out))) { set_error("direct_light_backward: bad args"); return SVGIR_ERR_INVALID; }
    if (n > 0)
        { TimedScope ts_("direct_light_bwd", s); direct_light_bwd_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, env_h, env_w, env_mode == 0 ? 2.0f : 1.0f, env_mode,
                                                               env, transform, dirs, g_out, d_env); }
    return check_launch("direct_light_backward", false, s);
}

}  // extern "C"
