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
// Mapping: a lane owns one (surfel, vertex) pair -- 8 surfels x 4 vertices per warp -- and loops
// over all Ns samples, so every per-vertex sum lives in registers and needs no cross-lane
// reduction.  Sample-level, vertex-independent work (env lookup, half vector, Fresnel power) is
// split across the 4 lanes of a surfel (one sample each) and exchanged with shuffles.
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

namespace svgir {

#define SH_THREADS 256
#define PI_F 3.14159265358979323846f

struct SampleShared {  // vertex-independent per-sample quantities, produced by one lane of the quad
    float wx, wy, wz;      // raw incident direction
    float lx, ly, lz;      // normalised
    float hx, hy, hz;      // half vector
    float hlen;            // |(L+V)/2|
    float voh_raw, p;      // V.H before clamp, 2^FMi
    float gr, gg, gb;      // area * clamp(env)*vis
    float lr, lg, lb;      // area * radiance
};

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// lat-long bilinear lookup (grid_sample, align_corners=True, zero padding). env [He][We][3].
// Returns texel corner (x0,y0) and weights so the backward can scatter.
struct EnvTap { int x0, y0; float wx1, wy1; };

__device__ __forceinline__ EnvTap env_coords(float dx, float dy, float dz, int He, int We) {
    const float phi = acosf(dz) - 1e-6f;
    const float theta = atan2f(dy, dx);
    const float qy = (phi / PI_F) * 2.f - 1.f;
    const float qx = -theta / PI_F;
    const float ix = (qx + 1.f) / 2.f * (float)(We - 1);
    const float iy = (qy + 1.f) / 2.f * (float)(He - 1);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    EnvTap t;
    t.x0 = (int)fx0; t.y0 = (int)fy0;
    t.wx1 = ix - fx0; t.wy1 = iy - fy0;
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

__global__ void env_activate_kernel(int n, const float* __restrict__ param, float* __restrict__ act, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) act[i] = mode == 0 ? softplus_f(param[i]) : param[i];
}

// ---------------------------------------------------------------------------------------------
struct ShadeArgs {
    int N, Ns, He, We;
    float env_scale;          // 2.0 for the learnable map (direct_light_map.py:83), 1.0 for HDR maps
    int env_in_smem;
    const float* env_act;     // [He,We,3] activated env
    const float* transform;   // optional [3,3] applied to dirs before the lookup (envmap.py:58-61)
    const float* base_color;  // [N,12] channel-major
    const float* roughness;   // [N,4]
    const float* metallic;    // [N,4] or null
    const float* normals;     // [N,4,3]
    const float* viewdirs;    // [N,3]
    const float* radiance;    // [N,Ns,3]
    const float* visibility;  // [N,Ns]
    const float* dirs;        // [N,Ns,3]
    const float* areas;       // [N,Ns]
};

struct VertexConst {  // per-lane constants
    float Nx, Ny, Nz;        // raw shading normal
    float inv_nlen;
    float tx, ty, tz;        // normalised, sign-flipped normal  N~
    float sgn;
    float Vx, Vy, Vz;        // normalised view dir
    float inv_vlen;
    float nov_raw, NoV;
    float r, a2, k;
};

__device__ __forceinline__ void load_vertex(const ShadeArgs& a, int n, int v, VertexConst& c) {
    const float* nn = a.normals + ((size_t)n * 4 + v) * 3;
    c.Nx = nn[0]; c.Ny = nn[1]; c.Nz = nn[2];
    const float nl = fmaxf(sqrtf(c.Nx * c.Nx + c.Ny * c.Ny + c.Nz * c.Nz), 1e-12f);
    c.inv_nlen = 1.f / nl;
    float hx = c.Nx * c.inv_nlen, hy = c.Ny * c.inv_nlen, hz = c.Nz * c.inv_nlen;
    const float* vd = a.viewdirs + (size_t)n * 3;
    const float vl = fmaxf(sqrtf(vd[0] * vd[0] + vd[1] * vd[1] + vd[2] * vd[2]), 1e-12f);
    c.inv_vlen = 1.f / vl;
    c.Vx = vd[0] * c.inv_vlen; c.Vy = vd[1] * c.inv_vlen; c.Vz = vd[2] * c.inv_vlen;
    const float d = c.Vx * hx + c.Vy * hy + c.Vz * hz;
    c.sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    c.tx = hx * c.sgn; c.ty = hy * c.sgn; c.tz = hz * c.sgn;
    c.nov_raw = c.tx * c.Vx + c.ty * c.Vy + c.tz * c.Vz;
    c.NoV = fminf(fmaxf(c.nov_raw, 1e-6f), 1.f);
    c.r = a.roughness[(size_t)n * 4 + v];
    const float al = c.r * c.r;
    c.a2 = al * al;
    c.k = (al + 2.f * c.r + 1.0f) / 8.0f;
}

// One lane prepares the vertex-independent part of sample s of surfel n.
__device__ __forceinline__ void prepare_sample(const ShadeArgs& a, const float* env, int n, int s, float Vx, float Vy,
                                               float Vz, SampleShared& o, float& vis, float raw_env[3],
                                               EnvTap& tap, float Lg[3], float Ll[3]) {
    const size_t is = (size_t)n * a.Ns + s;
    const float* d = a.dirs + is * 3;
    o.wx = d[0]; o.wy = d[1]; o.wz = d[2];
    const float il = 1.f / fmaxf(sqrtf(o.wx * o.wx + o.wy * o.wy + o.wz * o.wz), 1e-12f);
    o.lx = o.wx * il; o.ly = o.wy * il; o.lz = o.wz * il;
    float hx = (o.lx + Vx) * 0.5f, hy = (o.ly + Vy) * 0.5f, hz = (o.lz + Vz) * 0.5f;
    o.hlen = fmaxf(sqrtf(hx * hx + hy * hy + hz * hz), 1e-12f);
    const float ih = 1.f / o.hlen;
    o.hx = hx * ih; o.hy = hy * ih; o.hz = hz * ih;
    o.voh_raw = Vx * o.hx + Vy * o.hy + Vz * o.hz;
    const float voh = fminf(fmaxf(o.voh_raw, 1e-6f), 1.f);
    o.p = exp2f((-5.55473f * voh - 6.98316f) * voh);
    float qx = o.wx, qy = o.wy, qz = o.wz;  // direct_light uses the raw direction
    if (a.transform) {
        const float* t = a.transform;  // dirs @ transform.T
        const float tx = qx * t[0] + qy * t[1] + qz * t[2];
        const float ty = qx * t[3] + qy * t[4] + qz * t[5];
        const float tz = qx * t[6] + qy * t[7] + qz * t[8];
        qx = tx; qy = ty; qz = tz;
    }
    tap = env_coords(qx, qy, qz, a.He, a.We);
    env_fetch(env, a.He, a.We, tap, raw_env);
    vis = a.visibility[is];
    const float area = a.areas[is];
    const float* rad = a.radiance + is * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        Lg[ch] = fminf(fmaxf(raw_env[ch] * a.env_scale, 0.f), 64.f) * vis;
        Ll[ch] = rad[ch];
    }
    o.gr = area * Lg[0]; o.gg = area * Lg[1]; o.gb = area * Lg[2];
    o.lr = area * Ll[0]; o.lg = area * Ll[1]; o.lb = area * Ll[2];
}

__device__ __forceinline__ SampleShared quad_bcast(const SampleShared& m, int src_lane) {
    SampleShared o;
    const unsigned full = 0xffffffffu;
#define BC(f) o.f = __shfl_sync(full, m.f, src_lane)
    BC(wx); BC(wy); BC(wz); BC(lx); BC(ly); BC(lz); BC(hx); BC(hy); BC(hz); BC(hlen); BC(voh_raw); BC(p);
    BC(gr); BC(gg); BC(gb); BC(lr); BC(lg); BC(lb);
#undef BC
    return o;
}

struct VertexSample {  // vertex-level BRDF terms of one sample
    float ndi, ndi_raw, Dterm;
    float nol_raw, noh_raw, NoL, NoH, nom0, nom1, nom2, nom_raw;
};

__device__ __forceinline__ void eval_vertex(const VertexConst& c, const SampleShared& s, VertexSample& o) {
    o.ndi_raw = c.Nx * s.wx + c.Ny * s.wy + c.Nz * s.wz;
    o.ndi = fmaxf(o.ndi_raw, 0.f);
    o.nol_raw = c.tx * s.lx + c.ty * s.ly + c.tz * s.lz;
    o.noh_raw = c.tx * s.hx + c.ty * s.hy + c.tz * s.hz;
    o.NoL = fminf(fmaxf(o.nol_raw, 1e-6f), 1.f);
    o.NoH = fminf(fmaxf(o.noh_raw, 1e-6f), 1.f);
    o.nom0 = o.NoH * o.NoH * (c.a2 - 1.f) + 1.f;
    o.nom1 = c.NoV * (1.f - c.k) + c.k;
    o.nom2 = o.NoL * (1.f - c.k) + c.k;
    o.nom_raw = 4.f * PI_F * o.nom0 * o.nom0 * o.nom1 * o.nom2;
    const float nom = fminf(fmaxf(o.nom_raw, 1e-6f), 4.f * PI_F);
    o.Dterm = c.a2 / nom;
}

struct ShadeOut {
    float* pbr; float* diffuse; float* specular; float* direct; float* indirect;  // [N,12]
    float* mean_vis;       // [N,1]
    float* mean_local;     // [N,3]
    float* mean_incident;  // [N,3]
    float* mean_global;    // [N,3]
};

template <bool ENV_SMEM>
__global__ void __launch_bounds__(SH_THREADS) shade_fwd_kernel(const ShadeArgs a, const ShadeOut out) {
    extern __shared__ __align__(16) float env_s[];
    const float* env = a.env_act;
    if (ENV_SMEM) {
        for (int i = threadIdx.x; i < a.He * a.We * 3; i += SH_THREADS) env_s[i] = a.env_act[i];
        __syncthreads();
        env = env_s;
    }
    const int lane = threadIdx.x & 31;
    const int quad_base = lane & ~3, v = lane & 3;
    const int n = (blockIdx.x * SH_THREADS + threadIdx.x) >> 2;
    const bool valid = n < a.N;
    const int nc = valid ? n : a.N - 1;  // clamp so every lane runs the shuffles
    VertexConst c;
    load_vertex(a, nc, v, c);
    float Dg[3] = {0, 0, 0}, Dl[3] = {0, 0, 0}, SAg[3] = {0, 0, 0}, SBg[3] = {0, 0, 0}, SAl[3] = {0, 0, 0}, SBl[3] = {0, 0, 0};
    float m_vis = 0.f, m_g[3] = {0, 0, 0}, m_l[3] = {0, 0, 0};
    const int Ns = a.Ns;
    for (int s0 = 0; s0 < Ns; s0 += 4) {
        SampleShared mine;
        float vis = 0.f, raw_env[3], Lg_s[3], Ll_s[3];
        EnvTap tap;
        const int s = s0 + v;
        const bool sv = s < Ns;
        prepare_sample(a, env, nc, sv ? s : Ns - 1, c.Vx, c.Vy, c.Vz, mine, vis, raw_env, tap, Lg_s, Ll_s);
        if (sv) {
            m_vis += vis;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) { m_g[ch] += Lg_s[ch]; m_l[ch] += Ll_s[ch]; }
        }
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const SampleShared sh = quad_bcast(mine, quad_base + t);
            if (s0 + t >= Ns) continue;
            VertexSample vs;
            eval_vertex(c, sh, vs);
            const float tg[3] = {sh.gr * vs.ndi, sh.gg * vs.ndi, sh.gb * vs.ndi};
            const float tl[3] = {sh.lr * vs.ndi, sh.lg * vs.ndi, sh.lb * vs.ndi};
            const float pd = sh.p * vs.Dterm;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                Dg[ch] += tg[ch]; Dl[ch] += tl[ch];
                SAg[ch] = fmaf(vs.Dterm, tg[ch], SAg[ch]); SBg[ch] = fmaf(pd, tg[ch], SBg[ch]);
                SAl[ch] = fmaf(vs.Dterm, tl[ch], SAl[ch]); SBl[ch] = fmaf(pd, tl[ch], SBl[ch]);
            }
        }
    }
    // quad-reduce the sample means (each lane saw a quarter of the samples)
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        m_vis += __shfl_xor_sync(full, m_vis, o);
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            m_g[ch] += __shfl_xor_sync(full, m_g[ch], o);
            m_l[ch] += __shfl_xor_sync(full, m_l[ch], o);
        }
    }
    if (!valid) return;
    const float inv = 1.f / (float)Ns;
    const float met = a.metallic ? a.metallic[(size_t)n * 4 + v] : 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const size_t o = (size_t)n * 12 + 4 * ch + v;
        const float base = a.base_color[o];
        const float fd = (1.f - met) * base / PI_F;
        const float F0 = 0.04f * (1.f - met) + base * met;
        const float specG = (F0 * SAg[ch] + (1.f - F0) * SBg[ch]) * inv;
        const float specL = (F0 * SAl[ch] + (1.f - F0) * SBl[ch]) * inv;
        const float dg = Dg[ch] * inv, dl = Dl[ch] * inv;
        const float diff = dg + dl;
        out.diffuse[o] = diff;
        out.specular[o] = specG + specL;
        out.pbr[o] = fd * diff + (specG + specL);
        out.direct[o] = fd * dg + specG;
        out.indirect[o] = fd * dl + specL;
    }
    if (v == 0) {
        if (out.mean_vis) out.mean_vis[n] = m_vis * inv;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            if (out.mean_local) out.mean_local[(size_t)n * 3 + ch] = m_l[ch] * inv;
            if (out.mean_global) out.mean_global[(size_t)n * 3 + ch] = m_g[ch] * inv;
            if (out.mean_incident) out.mean_incident[(size_t)n * 3 + ch] = (m_l[ch] + m_g[ch]) * inv;
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct ShadeGrads {
    const float* g_pbr; const float* g_diffuse; const float* g_specular; const float* g_direct;
    const float* g_indirect;                     // [N,12] or null
    const float* g_mean_vis;                     // [N,1] or null
    const float* g_mean_local;                   // [N,3] or null
    const float* g_mean_incident;                // [N,3] or null
    const float* g_mean_global;                  // [N,3] or null
    const float* env_param;                      // raw parameter (learnable mode) for softplus'
    float* d_base_color; float* d_roughness; float* d_metallic; float* d_normals; float* d_viewdirs;
    float* d_radiance;                           // [N,Ns,3] or null
    float* d_visibility;                         // [N,Ns] or null
    float* d_env;                                // [He,We,3] accumulated with atomics, or null
};

template <bool ENV_SMEM>
__global__ void __launch_bounds__(SH_THREADS) shade_bwd_kernel(const ShadeArgs a, const ShadeGrads g, int env_mode) {
    extern __shared__ __align__(16) float smem_b[];
    const int nenv = a.He * a.We * 3;
    float* env_s = smem_b;                               // activated env (if it fits)
    float* denv_s = ENV_SMEM ? smem_b + nenv : nullptr;  // per-CTA env gradient accumulator
    const float* env = a.env_act;
    if (ENV_SMEM) {
        for (int i = threadIdx.x; i < nenv; i += SH_THREADS) { env_s[i] = a.env_act[i]; denv_s[i] = 0.f; }
        __syncthreads();
        env = env_s;
    }
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int quad_base = lane & ~3, v = lane & 3;
    const int n = (blockIdx.x * SH_THREADS + threadIdx.x) >> 2;
    const bool valid = n < a.N;
    const int nc = valid ? n : a.N - 1;
    const int Ns = a.Ns;
    const float inv = 1.f / (float)Ns;
    VertexConst c;
    load_vertex(a, nc, v, c);
    const float met = a.metallic ? a.metallic[(size_t)nc * 4 + v] : 0.f;

    // upstream gradients of the per-(surfel,vertex) sums
    float gDg[3], gDl[3], gSAg[3], gSBg[3], gSAl[3], gSBl[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const size_t o = (size_t)nc * 12 + 4 * ch + v;
        const float base = a.base_color[o];
        const float fd = (1.f - met) * base / PI_F;
        const float F0 = 0.04f * (1.f - met) + base * met;
        const float Gp = (valid && g.g_pbr) ? g.g_pbr[o] : 0.f;
        const float Gd = (valid && g.g_diffuse) ? g.g_diffuse[o] : 0.f;
        const float Gs = (valid && g.g_specular) ? g.g_specular[o] : 0.f;
        const float Gdi = (valid && g.g_direct) ? g.g_direct[o] : 0.f;
        const float Gin = (valid && g.g_indirect) ? g.g_indirect[o] : 0.f;
        gDg[ch] = (Gd + (Gp + Gdi) * fd) * inv;
        gDl[ch] = (Gd + (Gp + Gin) * fd) * inv;
        const float sg = (Gs + Gp + Gdi) * inv, sl = (Gs + Gp + Gin) * inv;
        gSAg[ch] = F0 * sg; gSBg[ch] = (1.f - F0) * sg;
        gSAl[ch] = F0 * sl; gSBl[ch] = (1.f - F0) * sl;
    }
    float gmv = 0.f, gml[3] = {0, 0, 0}, gmg[3] = {0, 0, 0};
    if (valid) {
        if (g.g_mean_vis) gmv = g.g_mean_vis[nc] * inv;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            float gi = g.g_mean_incident ? g.g_mean_incident[(size_t)nc * 3 + ch] : 0.f;
            gml[ch] = ((g.g_mean_local ? g.g_mean_local[(size_t)nc * 3 + ch] : 0.f) + gi) * inv;
            gmg[ch] = ((g.g_mean_global ? g.g_mean_global[(size_t)nc * 3 + ch] : 0.f) + gi) * inv;
        }
    }

    float dN[3] = {0, 0, 0};      // raw-normal gradient through n.w
    float dNt[3] = {0, 0, 0};     // gradient w.r.t. N~ (normalised, flipped)
    float dV[3] = {0, 0, 0};      // gradient w.r.t. normalised V
    float d_a2 = 0.f, d_k = 0.f, d_nov = 0.f;
    float Dl_sum[3] = {0, 0, 0}, Dg_sum[3] = {0, 0, 0};  // sum_s T (for d f_d)
    float SA_g[3] = {0, 0, 0}, SB_g[3] = {0, 0, 0}, SA_l[3] = {0, 0, 0}, SB_l[3] = {0, 0, 0};  // for d F0 (metallic)
    const bool has_met = a.metallic != nullptr;

    for (int s0 = 0; s0 < Ns; s0 += 4) {
        SampleShared mine;
        float vis = 0.f, raw_env[3];
        EnvTap tap;
        const int s = s0 + v;
        const bool sv = s < Ns;
        const int sc = sv ? s : Ns - 1;
        float Lg_s[3], Ll_s[3];
        prepare_sample(a, env, nc, sc, c.Vx, c.Vy, c.Vz, mine, vis, raw_env, tap, Lg_s, Ll_s);
        // per-sample gradient slots owned by this lane (its sample s): d(area*Lg), d(area*Ll)
        float dLg_mine[3] = {0, 0, 0}, dLl_mine[3] = {0, 0, 0};
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const SampleShared sh = quad_bcast(mine, quad_base + t);
            float dlg[3] = {0, 0, 0}, dll[3] = {0, 0, 0};
            if (s0 + t < Ns) {
                VertexSample vs;
                eval_vertex(c, sh, vs);
                const float Lg[3] = {sh.gr, sh.gg, sh.gb}, Ll[3] = {sh.lr, sh.lg, sh.lb};
                float d_ndi = 0.f, d_D = 0.f, d_p = 0.f;
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    const float tg = Lg[ch] * vs.ndi, tl = Ll[ch] * vs.ndi;
                    Dg_sum[ch] += tg; Dl_sum[ch] += tl;
                    if (has_met) {
                        SA_g[ch] = fmaf(vs.Dterm, tg, SA_g[ch]); SB_g[ch] = fmaf(sh.p * vs.Dterm, tg, SB_g[ch]);
                        SA_l[ch] = fmaf(vs.Dterm, tl, SA_l[ch]); SB_l[ch] = fmaf(sh.p * vs.Dterm, tl, SB_l[ch]);
                    }
                    const float gTg = gDg[ch] + vs.Dterm * (gSAg[ch] + sh.p * gSBg[ch]);
                    const float gTl = gDl[ch] + vs.Dterm * (gSAl[ch] + sh.p * gSBl[ch]);
                    dlg[ch] = gTg * vs.ndi;
                    dll[ch] = gTl * vs.ndi;
                    d_ndi += gTg * Lg[ch] + gTl * Ll[ch];
                    d_D += tg * (gSAg[ch] + sh.p * gSBg[ch]) + tl * (gSAl[ch] + sh.p * gSBl[ch]);
                    d_p += vs.Dterm * (tg * gSBg[ch] + tl * gSBl[ch]);
                }
                // ndi = max(N.w, 0)
                if (vs.ndi_raw >= 0.f) { dN[0] += d_ndi * sh.wx; dN[1] += d_ndi * sh.wy; dN[2] += d_ndi * sh.wz; }
                // Dterm = a2 / clamp(nom_raw)
                const bool nom_in = vs.nom_raw >= 1e-6f && vs.nom_raw <= 4.f * PI_F;
                const float nom = fminf(fmaxf(vs.nom_raw, 1e-6f), 4.f * PI_F);
                d_a2 += d_D / nom;
                const float d_nom = nom_in ? -d_D * c.a2 / (nom * nom) : 0.f;
                const float c4 = 4.f * PI_F * d_nom;
                const float d_nom0 = c4 * 2.f * vs.nom0 * vs.nom1 * vs.nom2;
                const float d_nom1 = c4 * vs.nom0 * vs.nom0 * vs.nom2;
                const float d_nom2 = c4 * vs.nom0 * vs.nom0 * vs.nom1;
                float d_noh = d_nom0 * 2.f * vs.NoH * (c.a2 - 1.f);
                d_a2 += d_nom0 * vs.NoH * vs.NoH;
                d_nov += d_nom1 * (1.f - c.k);
                d_k += d_nom1 * (1.f - c.NoV) + d_nom2 * (1.f - vs.NoL);
                float d_nol = d_nom2 * (1.f - c.k);
                if (!(vs.noh_raw >= 1e-6f && vs.noh_raw <= 1.f)) d_noh = 0.f;
                if (!(vs.nol_raw >= 1e-6f && vs.nol_raw <= 1.f)) d_nol = 0.f;
                // p = 2^((a1 voh + a0) voh)
                const float voh = fminf(fmaxf(sh.voh_raw, 1e-6f), 1.f);
                float d_voh = d_p * sh.p * 0.6931471805599453f * (2.f * -5.55473f * voh - 6.98316f);
                if (!(sh.voh_raw >= 1e-6f && sh.voh_raw <= 1.f)) d_voh = 0.f;
                // N~.L , N~.H , V.H
                dNt[0] += d_nol * sh.lx + d_noh * sh.hx;
                dNt[1] += d_nol * sh.ly + d_noh * sh.hy;
                dNt[2] += d_nol * sh.lz + d_noh * sh.hz;
                float dH[3] = {d_noh * c.tx + d_voh * c.Vx, d_noh * c.ty + d_voh * c.Vy, d_noh * c.tz + d_voh * c.Vz};
                dV[0] += d_voh * sh.hx; dV[1] += d_voh * sh.hy; dV[2] += d_voh * sh.hz;
                // H = h/|h|, h = (L + V)/2
                const float hd = sh.hx * dH[0] + sh.hy * dH[1] + sh.hz * dH[2];
                const float ih = 0.5f / sh.hlen;
                dV[0] += (dH[0] - sh.hx * hd) * ih;
                dV[1] += (dH[1] - sh.hy * hd) * ih;
                dV[2] += (dH[2] - sh.hz * hd) * ih;
            }
            // sum the 4 vertices' contributions to this sample's light gradients; owner lane keeps them
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                float x = dlg[ch], y = dll[ch];
                x += __shfl_xor_sync(full, x, 1); x += __shfl_xor_sync(full, x, 2);
                y += __shfl_xor_sync(full, y, 1); y += __shfl_xor_sync(full, y, 2);
                if (t == v) { dLg_mine[ch] = x; dLl_mine[ch] = y; }
            }
        }
        if (sv && valid) {
            const size_t is = (size_t)nc * Ns + s;
            const float area = a.areas[is];
            if (g.d_radiance) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++) g.d_radiance[is * 3 + ch] = dLl_mine[ch] * area + gml[ch];
            }
            // Lg = clamp(env_scale*raw,0,64)*vis
            float dvis = gmv, draw[3];
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                const float x = raw_env[ch] * a.env_scale;
                const float cl = fminf(fmaxf(x, 0.f), 64.f);
                const float dLg = dLg_mine[ch] * area + gmg[ch];
                dvis += dLg * cl;
                draw[ch] = (x >= 0.f && x <= 64.f) ? dLg * vis * a.env_scale : 0.f;
            }
            if (g.d_visibility) g.d_visibility[is] = dvis;
            if (g.d_env) {
                const float wx0 = 1.f - tap.wx1, wy0 = 1.f - tap.wy1;
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    const int x = tap.x0 + (kk & 1), y = tap.y0 + (kk >> 1);
                    if (x < 0 || x > a.We - 1 || y < 0 || y > a.He - 1) continue;
                    const float w = ((kk & 1) ? tap.wx1 : wx0) * ((kk >> 1) ? tap.wy1 : wy0);
                    const int base_i = (y * a.We + x) * 3;
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        const float val = draw[ch] * w;
                        if (val != 0.f) {
                            if (ENV_SMEM) atomicAdd(&denv_s[base_i + ch], val);
                            else atomicAdd(&g.d_env[base_i + ch], env_mode == 0 ? val / (1.f + expf(-g.env_param[base_i + ch])) : val);
                        }
                    }
                }
            }
        }
    }

    if (valid) {
        // roughness: a2 = r^4, k = (r^2 + 2r + 1)/8
        g.d_roughness[(size_t)n * 4 + v] = d_a2 * 4.f * c.r * c.r * c.r + d_k * (2.f * c.r + 2.f) / 8.0f;
        // base colour / metallic through f_d = (1-m) base/pi
        float dm = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const size_t o = (size_t)n * 12 + 4 * ch + v;
            const float base = a.base_color[o];
            const float Gp = g.g_pbr ? g.g_pbr[o] : 0.f;
            const float Gs = g.g_specular ? g.g_specular[o] : 0.f;
            const float Gdi = g.g_direct ? g.g_direct[o] : 0.f;
            const float Gin = g.g_indirect ? g.g_indirect[o] : 0.f;
            const float dgm = Dg_sum[ch] * inv, dlm = Dl_sum[ch] * inv;
            const float dfd = Gp * (dgm + dlm) + Gdi * dgm + Gin * dlm;
            float db = dfd * (1.f - met) / PI_F;
            dm += dfd * (-base / PI_F);
            if (has_met) {  // F0 = 0.04(1-m) + base*m
                const float dF0 = ((Gs + Gp) * ((SA_g[ch] + SA_l[ch]) - (SB_g[ch] + SB_l[ch])) +
                                   Gdi * (SA_g[ch] - SB_g[ch]) + Gin * (SA_l[ch] - SB_l[ch])) * inv;
                db += dF0 * met;
                dm += dF0 * (base - 0.04f);
            }
            g.d_base_color[o] = db;
        }
        if (g.d_metallic) g.d_metallic[(size_t)n * 4 + v] = dm;
        // NoV = clamp(N~.V): contributes to N~ and V
        if (!(c.nov_raw >= 1e-6f && c.nov_raw <= 1.f)) d_nov = 0.f;
        dNt[0] += d_nov * c.Vx; dNt[1] += d_nov * c.Vy; dNt[2] += d_nov * c.Vz;
        dV[0] += d_nov * c.tx; dV[1] += d_nov * c.ty; dV[2] += d_nov * c.tz;
        // N~ = sgn * N/|N|
        const float hx = c.tx * c.sgn, hy = c.ty * c.sgn, hz = c.tz * c.sgn;  // N^ (sgn^2 = 1 unless 0)
        const float dh[3] = {dNt[0] * c.sgn, dNt[1] * c.sgn, dNt[2] * c.sgn};
        const float nd = hx * dh[0] + hy * dh[1] + hz * dh[2];
        float* dn = g.d_normals + ((size_t)n * 4 + v) * 3;
        dn[0] = dN[0] + (dh[0] - hx * nd) * c.inv_nlen;
        dn[1] = dN[1] + (dh[1] - hy * nd) * c.inv_nlen;
        dn[2] = dN[2] + (dh[2] - hz * nd) * c.inv_nlen;
    } else {
        dV[0] = dV[1] = dV[2] = 0.f;
    }
    // view direction: sum over the 4 vertices, then through the normalisation
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        dV[ch] += __shfl_xor_sync(full, dV[ch], 1);
        dV[ch] += __shfl_xor_sync(full, dV[ch], 2);
    }
    if (valid && v == 0) {
        const float vd = c.Vx * dV[0] + c.Vy * dV[1] + c.Vz * dV[2];
        float* o = g.d_viewdirs + (size_t)n * 3;
        o[0] = (dV[0] - c.Vx * vd) * c.inv_vlen;
        o[1] = (dV[1] - c.Vy * vd) * c.inv_vlen;
        o[2] = (dV[2] - c.Vz * vd) * c.inv_vlen;
    }
    if (ENV_SMEM && g.d_env) {
        __syncthreads();
        for (int i = threadIdx.x; i < nenv; i += SH_THREADS) {
            float val = denv_s[i];
            if (val != 0.f) {
                if (env_mode == 0) val = val / (1.f + expf(-g.env_param[i]));  // softplus'
                atomicAdd(&g.d_env[i], val);
            }
        }
    }
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

}  // namespace svgir

using namespace svgir;

extern "C" {

static int shade_prepare(const svgir_shade_cfg* c, const svgir_shade_in* in, ShadeArgs& a, cudaStream_t s) {
    if (!c || !in || c->N < 0 || c->Ns <= 0 || c->env_h <= 0 || c->env_w <= 0) { set_error("shade: bad cfg"); return SVGIR_ERR_INVALID; }
    if (!in->base_color || !in->roughness || !in->normals || !in->viewdirs || !in->radiance || !in->visibility ||
        !in->incident_dirs || !in->incident_areas || !in->env || !in->env_act_scratch) {
        set_error("shade: missing input");
        return SVGIR_ERR_INVALID;
    }
    const int nenv = c->env_h * c->env_w * 3;
    { TimedScope ts_("env_activate", s); env_activate_kernel<<<(nenv + 255) / 256, 256, 0, s>>>(nenv, in->env, in->env_act_scratch, c->env_mode); }
    a.N = c->N; a.Ns = c->Ns; a.He = c->env_h; a.We = c->env_w;
    a.env_scale = c->env_mode == 0 ? 2.0f : 1.0f;
    a.env_act = in->env_act_scratch; a.transform = in->env_transform;
    a.base_color = in->base_color; a.roughness = in->roughness; a.metallic = in->metallic;
    a.normals = in->normals; a.viewdirs = in->viewdirs; a.radiance = in->radiance;
    a.visibility = in->visibility; a.dirs = in->incident_dirs; a.areas = in->incident_areas;
    return SVGIR_OK;
}

int svgir_shade_forward(const svgir_shade_cfg* c, const svgir_shade_in* in, const svgir_shade_out* o, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    ShadeArgs a;
    int rc = shade_prepare(c, in, a, s);
    if (rc) return rc;
    if (!o || !o->pbr || !o->diffuse_light || !o->specular || !o->direct || !o->indirect) { set_error("shade: missing output"); return SVGIR_ERR_INVALID; }
    if (c->N == 0) return SVGIR_OK;
    ShadeOut so{o->pbr, o->diffuse_light, o->specular, o->direct, o->indirect, o->mean_visibility,
                o->mean_local, o->mean_incident, o->mean_global};
    const size_t env_bytes = (size_t)a.He * a.We * 3 * sizeof(float);
    const int grid = (int)(((size_t)c->N * 4 + SH_THREADS - 1) / SH_THREADS);
    if (env_bytes <= 96 * 1024) {
        if (env_bytes > 48 * 1024)
            cudaFuncSetAttribute(shade_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)env_bytes);
        { TimedScope ts_("shade_fwd", s); shade_fwd_kernel<true><<<grid, SH_THREADS, env_bytes, s>>>(a, so); }
    } else {
        { TimedScope ts_("shade_fwd", s); shade_fwd_kernel<false><<<grid, SH_THREADS, 0, s>>>(a, so); }
    }
    return check_launch("shade_forward", c->debug, s);
}

int svgir_shade_backward(const svgir_shade_cfg* c, const svgir_shade_in* in, const svgir_shade_grads* gr, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    ShadeArgs a;
    int rc = shade_prepare(c, in, a, s);
    if (rc) return rc;
    if (!gr || !gr->d_base_color || !gr->d_roughness || !gr->d_normals || !gr->d_viewdirs) { set_error("shade_backward: missing buffers"); return SVGIR_ERR_INVALID; }
    if (c->N == 0) return SVGIR_OK;
    ShadeGrads g{gr->g_pbr, gr->g_diffuse_light, gr->g_specular, gr->g_direct, gr->g_indirect, gr->g_mean_visibility,
                 gr->g_mean_local, gr->g_mean_incident, gr->g_mean_global, in->env,
                 gr->d_base_color, gr->d_roughness, gr->d_metallic, gr->d_normals, gr->d_viewdirs,
                 gr->d_radiance, gr->d_visibility, gr->d_env};
    const size_t env_bytes = (size_t)a.He * a.We * 3 * sizeof(float);
    const int grid = (int)(((size_t)c->N * 4 + SH_THREADS - 1) / SH_THREADS);
    if (2 * env_bytes <= 96 * 1024) {
        if (2 * env_bytes > 48 * 1024)
            cudaFuncSetAttribute(shade_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * env_bytes));
        { TimedScope ts_("shade_bwd", s); shade_bwd_kernel<true><<<grid, SH_THREADS, 2 * env_bytes, s>>>(a, g, c->env_mode); }
    } else {
        { TimedScope ts_("shade_bwd", s); shade_bwd_kernel<false><<<grid, SH_THREADS, 0, s>>>(a, g, c->env_mode); }
    }
    return check_launch("shade_backward", c->debug, s);
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
