// SH-lit render_equation (per-surfel scalar material, metallic workflow, lighting stored as degree-3
// SH). B200 implementation of the reference's R3DG-style CUDA operators
//   RenderEquationForwardCUDA / _complex / RenderEquationBackwardCUDA
//   (rgss-rasterization/render_equation.cu:555-729, 55-277, 280-550; decls render_equation.h:7-46),
// which the reference ships but never builds.
//
// The reference runs one THREAD per surfel looping over the samples, so its per-sample stores
// (`incident_dirs[idx*Ns+ray]`, 12 B at a stride of Ns*12 B across the warp) touch one 32-B sector
// per lane. Here one WARP owns a surfel and its lanes own samples: per-sample stores of a warp are
// contiguous (32 x 12 B = 384 B), the surfel's SH rows (48+16 floats) are warp-broadcast loads, and
// the per-surfel sums are butterfly-reduced in registers.
#include "common.cuh"

namespace svgir {
namespace {

constexpr float PI_R = 3.14159f;  // the reference's literal (render_equation.cu:92)
constexpr float SH0 = 0.28209479177387814f, SH1 = 0.4886025119029199f;
__constant__ float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                             -1.0925484305920792f, 0.5462742152960396f};
__constant__ float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                             -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

// computeSHcoef (render_equation.cu:19-53), degree 3
__device__ __forceinline__ void sh_basis(float x, float y, float z, float* c) {
    c[0] = SH0;
    c[1] = -SH1 * y; c[2] = SH1 * z; c[3] = -SH1 * x;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    c[4] = kC2[0] * xy; c[5] = kC2[1] * yz; c[6] = kC2[2] * (2.0f * zz - xx - yy); c[7] = kC2[3] * xz;
    c[8] = kC2[4] * (xx - yy);
    c[9] = kC3[0] * y * (3.0f * xx - yy); c[10] = kC3[1] * xy * z; c[11] = kC3[2] * y * (4.0f * zz - xx - yy);
    c[12] = kC3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy); c[13] = kC3[4] * x * (4.0f * zz - xx - yy);
    c[14] = kC3[5] * z * (xx - yy); c[15] = kC3[6] * x * (xx - 3.0f * yy);
}

// Fibonacci direction of sample `ray` rotated from +z to `n` (render_equation.cu:90-117 / 590-618)
__device__ __forceinline__ float3 fib_dir(int ray, int Ns, float rnd, bool use_rnd, float3 n) {
    const float delta = PI_R * (3.0f - sqrtf(5.0f));
    float z = 1 - 2 * (float)ray / (2 * (float)Ns - 1);
    float rad = sqrtf(1 - z * z);
    float theta = delta * ray;
    if (use_rnd) theta = rnd * 2 * PI_R + theta;
    float y = cosf(theta) * rad, x = sinf(theta) * rad;
    float v1 = -n.y, v2 = n.x;
    float c = fmaxf(n.z + 1, 0.0000001f);
    float v12 = v1 * v2;
    float ox = (1 + (-v2 * v2) / c) * x + (v12 / c) * y + v2 * z;
    float oy = (v12 / c) * x + (1 + (-v1 * v1) / c) * y + (-v1) * z;
    float oz = (-v2) * x + v1 * y + (1 + (-v2 * v2 - v1 * v1) / c) * z;
    float nn = sqrtf(fmaxf(0.0000001f, ox * ox + oy * oy + oz * oz));
    return make_float3(ox / nn, oy / nn, oz / nn);
}

struct Lights {
    float3 local, global_raw, global;  // global_raw: before the visibility factor
    float vis;
};

// SH lighting of one direction (render_equation.cu:122-142)
__device__ __forceinline__ Lights eval_lights(const float* coef, const float* __restrict__ inc, int S_inc,
                                              const float* __restrict__ dsh, int S_dir,
                                              const float* __restrict__ vsh, int S_vis) {
    Lights L;
    float3 a = make_float3(0.f, 0.f, 0.f);
    _Pragma("unroll") for (int i = 0; i < 16; i++) if (i < S_inc) { a.x += inc[3 * i] * coef[i]; a.y += inc[3 * i + 1] * coef[i]; a.z += inc[3 * i + 2] * coef[i]; }
    L.local = make_float3(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f));
    float3 g = make_float3(0.5f, 0.5f, 0.5f);
    _Pragma("unroll") for (int i = 0; i < 16; i++) if (i < S_dir) { g.x += dsh[3 * i] * coef[i]; g.y += dsh[3 * i + 1] * coef[i]; g.z += dsh[3 * i + 2] * coef[i]; }
    L.global_raw = make_float3(fmaxf(g.x, 0.f), fmaxf(g.y, 0.f), fmaxf(g.z, 0.f));
    float v = 0.5f;
    _Pragma("unroll") for (int i = 0; i < 16; i++) if (i < S_vis) v += vsh[i] * coef[i];
    L.vis = fmaxf(0.0f, fminf(v, 1.0f));
    L.global = make_float3(L.vis * L.global_raw.x, L.vis * L.global_raw.y, L.vis * L.global_raw.z);
    return L;
}

__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct Brdf {
    float3 half_n;  // normalised half vector
    float half_norm, h_d_n, h_d_o, n_d_i, n_d_o;
    float r2, amp, sharp, e, D, r2v, den1, den2, g1, g2, V, pw5;
    float3 F0, F;
};

// SG-NDF / Schlick / Schlick-GGX terms (render_equation.cu:147-168)
__device__ __forceinline__ Brdf eval_brdf(float3 d, float3 view, float3 n, float3 base, float metal, float rough,
                                          bool exact_pow) {
    Brdf b;
    float3 h = make_float3(d.x + view.x, d.y + view.y, d.z + view.z);
    b.half_norm = fmaxf(sqrtf(dot3(h, h)), 0.0000001f);
    b.half_n = make_float3(h.x / b.half_norm, h.y / b.half_norm, h.z / b.half_norm);
    b.h_d_n = fmaxf(dot3(b.half_n, n), 0.0f);
    b.h_d_o = fmaxf(dot3(b.half_n, view), 0.0f);
    b.n_d_i = fmaxf(dot3(n, d), 0.0f);
    b.n_d_o = fmaxf(dot3(n, view), 0.0f);
    b.r2 = fmaxf(rough * rough, 0.0000001f);
    b.amp = 1.0f / (b.r2 * PI_R);
    b.sharp = 2.0f / b.r2;
    b.e = expf(b.sharp * (b.h_d_n - 1.0f));
    b.D = b.amp * b.e;
    b.F0 = make_float3(0.04f * (1.0f - metal) + base.x * metal, 0.04f * (1.0f - metal) + base.y * metal,
                       0.04f * (1.0f - metal) + base.z * metal);
    b.pw5 = powf(1.0f - b.h_d_o, 5.0f);
    b.F = make_float3(b.F0.x + (1.0f - b.F0.x) * b.pw5, b.F0.y + (1.0f - b.F0.y) * b.pw5, b.F0.z + (1.0f - b.F0.z) * b.pw5);
    // forward uses __powf (render_equation.cu:166), backward powf (:369)
    b.r2v = (exact_pow ? powf(1.0f + rough, 2.0f) : __powf(1.0f + rough, 2.0f)) / 8.0f;
    b.den1 = fmaxf(b.n_d_i * (1 - b.r2v) + b.r2v, 0.0000001f);
    b.den2 = fmaxf(b.n_d_o * (1 - b.r2v) + b.r2v, 0.0000001f);
    b.g1 = 0.5f / b.den1;
    b.g2 = 0.5f / b.den2;
    b.V = b.g1 * b.g2;
    return b;
}

struct FwdPtrs {
    const float *base_color, *roughness, *metallic, *normals, *viewdirs, *incidents_shs, *direct_shs, *visibility_shs,
        *rand_float;
    float *pbr, *incident_dirs, *diffuse_light;
    float *incident_lights, *local_incident_lights, *global_incident_lights, *incident_visibility, *local_diffuse_light,
        *accum, *rgb_d, *rgb_s;
};

template <bool COMPLEX>
__global__ void __launch_bounds__(256) req_sh_forward_kernel(int P, int S_inc, int S_dir, int S_vis, int Ns, int is_training,
                                                             FwdPtrs p) {
    __shared__ float s_dsh[48];
    if (threadIdx.x < 3 * S_dir) s_dsh[threadIdx.x] = p.direct_shs[threadIdx.x];
    __syncthreads();
    int lane = threadIdx.x & 31;
    int warps = (gridDim.x * blockDim.x) >> 5;
    for (int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < P; idx += warps) {
        float3 n = make_float3(p.normals[3 * idx], p.normals[3 * idx + 1], p.normals[3 * idx + 2]);
        float3 view = make_float3(p.viewdirs[3 * idx], p.viewdirs[3 * idx + 1], p.viewdirs[3 * idx + 2]);
        float3 base = make_float3(p.base_color[3 * idx], p.base_color[3 * idx + 1], p.base_color[3 * idx + 2]);
        float metal = p.metallic[idx], rough = p.roughness[idx];
        const float* inc = p.incidents_shs + (size_t)idx * S_inc * 3;
        const float* vsh = p.visibility_shs + (size_t)idx * S_vis;
        float3 f_d = make_float3((1 - metal) * base.x / PI_R, (1 - metal) * base.y / PI_R, (1 - metal) * base.z / PI_R);
        float3 a_d = make_float3(0, 0, 0), a_s = a_d, a_diff = a_d, a_ldiff = a_d;
        for (int ray = lane; ray < Ns; ray += 32) {
            size_t w = (size_t)idx * Ns + ray;
            bool rnd = (!COMPLEX) && is_training && p.rand_float;
            float3 d = fib_dir(ray, Ns, rnd ? p.rand_float[w] : 0.f, rnd, n);
            float coef[16];
            sh_basis(d.x, d.y, d.z, coef);
            Lights L = eval_lights(coef, inc, S_inc, s_dsh, S_dir, vsh, S_vis);
            float3 li = make_float3(L.global.x + L.local.x, L.global.y + L.local.y, L.global.z + L.local.z);
            Brdf b = eval_brdf(d, view, n, base, metal, rough, false);
            float dv = b.D * b.V;
            float3 f_s = make_float3(dv * b.F.x, dv * b.F.y, dv * b.F.z);
            float t = 2.0f * PI_R * b.n_d_i / (float)Ns;
            float3 tr = make_float3(li.x * t, li.y * t, li.z * t);
            a_diff.x += tr.x; a_diff.y += tr.y; a_diff.z += tr.z;
            a_d.x += f_d.x * tr.x; a_d.y += f_d.y * tr.y; a_d.z += f_d.z * tr.z;
            a_s.x += f_s.x * tr.x; a_s.y += f_s.y * tr.y; a_s.z += f_s.z * tr.z;
            p.incident_dirs[3 * w] = d.x; p.incident_dirs[3 * w + 1] = d.y; p.incident_dirs[3 * w + 2] = d.z;
            if (COMPLEX) {
                a_ldiff.x += L.local.x * t; a_ldiff.y += L.local.y * t; a_ldiff.z += L.local.z * t;
                p.incident_lights[3 * w] = li.x; p.incident_lights[3 * w + 1] = li.y; p.incident_lights[3 * w + 2] = li.z;
                p.local_incident_lights[3 * w] = L.local.x; p.local_incident_lights[3 * w + 1] = L.local.y;
                p.local_incident_lights[3 * w + 2] = L.local.z;
                p.global_incident_lights[3 * w] = L.global.x; p.global_incident_lights[3 * w + 1] = L.global.y;
                p.global_incident_lights[3 * w + 2] = L.global.z;
                p.incident_visibility[w] = L.vis;
            }
        }
        a_d.x = warp_sum(a_d.x); a_d.y = warp_sum(a_d.y); a_d.z = warp_sum(a_d.z);
        a_s.x = warp_sum(a_s.x); a_s.y = warp_sum(a_s.y); a_s.z = warp_sum(a_s.z);
        a_diff.x = warp_sum(a_diff.x); a_diff.y = warp_sum(a_diff.y); a_diff.z = warp_sum(a_diff.z);
        if (COMPLEX) { a_ldiff.x = warp_sum(a_ldiff.x); a_ldiff.y = warp_sum(a_ldiff.y); a_ldiff.z = warp_sum(a_ldiff.z); }
        if (lane == 0) {
            p.pbr[3 * idx] = a_d.x + a_s.x; p.pbr[3 * idx + 1] = a_d.y + a_s.y; p.pbr[3 * idx + 2] = a_d.z + a_s.z;
            p.diffuse_light[3 * idx] = a_diff.x; p.diffuse_light[3 * idx + 1] = a_diff.y; p.diffuse_light[3 * idx + 2] = a_diff.z;
            if (COMPLEX) {
                float ax = a_diff.x / PI_R + a_s.x, ay = a_diff.y / PI_R + a_s.y, az = a_diff.z / PI_R + a_s.z;
                p.accum[idx] = (ax + ay + az) / 3;
                p.rgb_d[3 * idx] = a_d.x; p.rgb_d[3 * idx + 1] = a_d.y; p.rgb_d[3 * idx + 2] = a_d.z;
                p.rgb_s[3 * idx] = a_s.x; p.rgb_s[3 * idx + 1] = a_s.y; p.rgb_s[3 * idx + 2] = a_s.z;
                p.local_diffuse_light[3 * idx] = a_ldiff.x; p.local_diffuse_light[3 * idx + 1] = a_ldiff.y;
                p.local_diffuse_light[3 * idx + 2] = a_ldiff.z;
            }
        }
    }
}

struct BwdPtrs {
    const float *base_color, *roughness, *metallic, *normals, *viewdirs, *incidents_shs, *direct_shs, *visibility_shs;
    const float *incident_dirs, *dL_dpbr, *dL_ddiffuse_light;
    float *dL_dbase_color, *dL_droughness, *dL_dmetallic, *dL_dnormals, *dL_dviewdirs, *dL_dincidents_shs, *dL_ddirect_shs,
        *dL_dvisibility_shs;
};

// Backward (render_equation.cu:280-470). `legacy` reproduces the reference arithmetic exactly, including
//   * dL_dn_d_i being OVERWRITTEN by the geometry-term gradient (:406), which drops the transport term,
//   * the incident-SH gradient loop running over S_direct coefficients (:453),
//   * the clamp masks testing the already-clamped values (:438,444-446,450-452), i.e. no masking;
// legacy == 0 gives the analytic gradient of the forward instead. The reference accumulates
// dL_ddirect_shs with unsynchronised `+=` from every thread (:447-449, a data race); here it is the
// exact sum (block-level reduction + one atomic per coefficient per block).
__global__ void __launch_bounds__(128) req_sh_backward_kernel(int P, int S_inc, int S_dir, int S_vis, int Ns, int legacy,
                                                              BwdPtrs p) {
    __shared__ float s_dsh[48];
    __shared__ float s_gd[48];
    if (threadIdx.x < 48) { s_dsh[threadIdx.x] = threadIdx.x < 3 * S_dir ? p.direct_shs[threadIdx.x] : 0.f; s_gd[threadIdx.x] = 0.f; }
    __syncthreads();
    int lane = threadIdx.x & 31;
    int warps = (gridDim.x * blockDim.x) >> 5;
    int n_inc_g = legacy ? min(S_dir, S_inc) : S_inc;
    for (int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < P; idx += warps) {
        float3 n = make_float3(p.normals[3 * idx], p.normals[3 * idx + 1], p.normals[3 * idx + 2]);
        float3 view = make_float3(p.viewdirs[3 * idx], p.viewdirs[3 * idx + 1], p.viewdirs[3 * idx + 2]);
        float3 base = make_float3(p.base_color[3 * idx], p.base_color[3 * idx + 1], p.base_color[3 * idx + 2]);
        float3 g_pbr = make_float3(p.dL_dpbr[3 * idx], p.dL_dpbr[3 * idx + 1], p.dL_dpbr[3 * idx + 2]);
        float3 g_dl = make_float3(p.dL_ddiffuse_light[3 * idx], p.dL_ddiffuse_light[3 * idx + 1], p.dL_ddiffuse_light[3 * idx + 2]);
        float metal = p.metallic[idx], rough = p.roughness[idx];
        const float* inc = p.incidents_shs + (size_t)idx * S_inc * 3;
        const float* vsh = p.visibility_shs + (size_t)idx * S_vis;
        float3 f_d = make_float3((1 - metal) * base.x / PI_R, (1 - metal) * base.y / PI_R, (1 - metal) * base.z / PI_R);
        float a_base[3] = {0, 0, 0}, a_n[3] = {0, 0, 0}, a_v[3] = {0, 0, 0}, a_metal = 0.f, a_rough = 0.f;
        float a_inc[48], a_vis[16], a_dir[48];
#pragma unroll
        for (int i = 0; i < 48; i++) { a_inc[i] = 0.f; a_dir[i] = 0.f; }
#pragma unroll
        for (int i = 0; i < 16; i++) a_vis[i] = 0.f;
        for (int ray = lane; ray < Ns; ray += 32) {
            size_t w = (size_t)idx * Ns + ray;
            float3 d = make_float3(p.incident_dirs[3 * w], p.incident_dirs[3 * w + 1], p.incident_dirs[3 * w + 2]);
            float coef[16];
            sh_basis(d.x, d.y, d.z, coef);
            // un-clamped sums are needed for the analytic masks
            float3 lraw = make_float3(0, 0, 0), graw = make_float3(0.5f, 0.5f, 0.5f);
            float vraw = 0.5f;
            _Pragma("unroll") for (int i = 0; i < 16; i++) if (i < S_inc) { lraw.x += inc[3 * i] * coef[i]; lraw.y += inc[3 * i + 1] * coef[i]; lraw.z += inc[3 * i + 2] * coef[i]; }
            _Pragma("unroll") for (int i = 0; i < 16; i++) if (i < S_dir) { graw.x += s_dsh[3 * i] * coef[i]; graw.y += s_dsh[3 * i + 1] * coef[i]; graw.z += s_dsh[3 * i + 2] * coef[i]; }
            _Pragma("unroll") for (int i = 0; i < 16; i++) if (i < S_vis) vraw += vsh[i] * coef[i];
            float3 local = make_float3(fmaxf(lraw.x, 0.f), fmaxf(lraw.y, 0.f), fmaxf(lraw.z, 0.f));
            float3 glob = make_float3(fmaxf(graw.x, 0.f), fmaxf(graw.y, 0.f), fmaxf(graw.z, 0.f));
            float vis = fmaxf(0.0f, fminf(vraw, 1.0f));
            float3 li = make_float3(vis * glob.x + local.x, vis * glob.y + local.y, vis * glob.z + local.z);
            Brdf b = eval_brdf(d, view, n, base, metal, rough, true);
            float dv = b.D * b.V;
            float3 f_s = make_float3(dv * b.F.x, dv * b.F.y, dv * b.F.z);
            float t = 2.0f * PI_R * b.n_d_i / (float)Ns;
            float t0 = 2.0f * PI_R / (float)Ns;
            float3 fds = make_float3(f_d.x + f_s.x, f_d.y + f_s.y, f_d.z + f_s.z);
            float3 g_f = make_float3(g_pbr.x * li.x * t, g_pbr.y * li.y * t, g_pbr.z * li.z * t);  // dL_dfd == dL_dfs
            float3 g_li = make_float3(g_pbr.x * fds.x * t + g_dl.x * t, g_pbr.y * fds.y * t + g_dl.y * t,
                                      g_pbr.z * fds.z * t + g_dl.z * t);
            float g_ndi = (g_pbr.x * fds.x * li.x + g_pbr.y * fds.y * li.y + g_pbr.z * fds.z * li.z) * t0 +
                          (g_dl.x * li.x + g_dl.y * li.y + g_dl.z * li.z) * t0;
            // diffuse lobe
            float3 g_base = make_float3(g_f.x * (1 - metal) / PI_R, g_f.y * (1 - metal) / PI_R, g_f.z * (1 - metal) / PI_R);
            float g_metal = -dot3(g_f, base) / PI_R;
            // specular lobe
            float g_D = b.V * dot3(g_f, b.F);
            float3 g_F = make_float3(g_f.x * dv, g_f.y * dv, g_f.z * dv);
            float g_V = b.D * dot3(g_f, b.F);
            float g_amp = g_D * b.e, g_e = g_D * b.amp;
            float g_sharp = (b.h_d_n - 1.0f) * b.e * g_e;
            float g_hdn = b.sharp * b.e * g_e;
            float g_r2 = -2.0f / (b.r2 * b.r2) * g_sharp - 1.0f / (b.r2 * b.r2 * PI_R) * g_amp;
            float g_rough = g_r2 * 2.0f * rough;
            float3 g_F0 = make_float3((1.0f - b.pw5) * g_F.x, (1.0f - b.pw5) * g_F.y, (1.0f - b.pw5) * g_F.z);
            float g_hdo = ((1.0f - b.F0.x) * g_F.x + (1.0f - b.F0.y) * g_F.y + (1.0f - b.F0.z) * g_F.z) * -5.0f *
                          powf(1.0f - b.h_d_o, 4.0f);
            g_base.x += metal * g_F0.x; g_base.y += metal * g_F0.y; g_base.z += metal * g_F0.z;
            g_metal += (base.x - 0.04f) * g_F0.x + (base.y - 0.04f) * g_F0.y + (base.z - 0.04f) * g_F0.z;
            float g_g1 = g_V * b.g2, g_g2 = g_V * b.g1;
            float g_den1 = -0.5f / (b.den1 * b.den1) * g_g1;
            float g_den2 = -0.5f / (b.den2 * b.den2) * g_g2;
            if (legacy) g_ndi = g_den1 * (1 - b.r2v);       // render_equation.cu:406 (overwrite)
            else g_ndi += g_den1 * (1 - b.r2v);
            float g_ndo = g_den2 * (1 - b.r2v);
            float g_r2v = (1.0f - b.n_d_i) * g_den1 + (1.0f - b.n_d_o) * g_den2;
            g_rough += (1.0f + rough) / 4.0f * g_r2v;
            float3 g_h = make_float3(0, 0, 0), g_n = g_h, g_v = g_h;
            if (b.h_d_n > 0.0f) {
                g_h.x += n.x * g_hdn; g_h.y += n.y * g_hdn; g_h.z += n.z * g_hdn;
                g_n.x += b.half_n.x * g_hdn; g_n.y += b.half_n.y * g_hdn; g_n.z += b.half_n.z * g_hdn;
            }
            if (b.h_d_o > 0.0f) {
                g_h.x += view.x * g_hdo; g_h.y += view.y * g_hdo; g_h.z += view.z * g_hdo;
                g_v.x += b.half_n.x * g_hdo; g_v.y += b.half_n.y * g_hdo; g_v.z += b.half_n.z * g_hdo;
            }
            if (b.n_d_i > 0.0f) { g_n.x += d.x * g_ndi; g_n.y += d.y * g_ndi; g_n.z += d.z * g_ndi; }
            if (b.n_d_o > 0.0f) {
                g_n.x += view.x * g_ndo; g_n.y += view.y * g_ndo; g_n.z += view.z * g_ndo;
                g_v.x += n.x * g_ndo; g_v.y += n.y * g_ndo; g_v.z += n.z * g_ndo;
            }
            if (legacy) {  // "TODO:consider norm" (:432-433): the normalisation Jacobian is ignored
                g_v.x += g_h.x / b.half_norm; g_v.y += g_h.y / b.half_norm; g_v.z += g_h.z / b.half_norm;
            } else {
                float hp = dot3(g_h, b.half_n);
                g_v.x += (g_h.x - b.half_n.x * hp) / b.half_norm; g_v.y += (g_h.y - b.half_n.y * hp) / b.half_norm;
                g_v.z += (g_h.z - b.half_n.z * hp) / b.half_norm;
            }
            // lighting
            float3 g_loc = g_li;
            float3 g_glob = make_float3(g_li.x * vis, g_li.y * vis, g_li.z * vis);
            float g_vis = g_li.x * glob.x + g_li.y * glob.y + g_li.z * glob.z;
            if (!legacy) {
                if (vraw < 0.f || vraw > 1.f) g_vis = 0.f;
                if (graw.x < 0.f) g_glob.x = 0.f;
                if (graw.y < 0.f) g_glob.y = 0.f;
                if (graw.z < 0.f) g_glob.z = 0.f;
                if (lraw.x < 0.f) g_loc.x = 0.f;
                if (lraw.y < 0.f) g_loc.y = 0.f;
                if (lraw.z < 0.f) g_loc.z = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 16; i++) {
                a_vis[i] += g_vis * coef[i];
                a_dir[3 * i] += g_glob.x * coef[i]; a_dir[3 * i + 1] += g_glob.y * coef[i]; a_dir[3 * i + 2] += g_glob.z * coef[i];
                a_inc[3 * i] += g_loc.x * coef[i]; a_inc[3 * i + 1] += g_loc.y * coef[i]; a_inc[3 * i + 2] += g_loc.z * coef[i];
            }
            a_base[0] += g_base.x; a_base[1] += g_base.y; a_base[2] += g_base.z;
            a_n[0] += g_n.x; a_n[1] += g_n.y; a_n[2] += g_n.z;
            a_v[0] += g_v.x; a_v[1] += g_v.y; a_v[2] += g_v.z;
            a_metal += g_metal; a_rough += g_rough;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) { a_base[k] = warp_sum(a_base[k]); a_n[k] = warp_sum(a_n[k]); a_v[k] = warp_sum(a_v[k]); }
        a_metal = warp_sum(a_metal); a_rough = warp_sum(a_rough);
#pragma unroll
        for (int i = 0; i < 48; i++) { a_inc[i] = warp_sum(a_inc[i]); a_dir[i] = warp_sum(a_dir[i]); }
#pragma unroll
        for (int i = 0; i < 16; i++) a_vis[i] = warp_sum(a_vis[i]);
        if (lane == 0) {
            for (int k = 0; k < 3; k++) {
                p.dL_dbase_color[3 * idx + k] = a_base[k]; p.dL_dnormals[3 * idx + k] = a_n[k]; p.dL_dviewdirs[3 * idx + k] = a_v[k];
            }
            p.dL_dmetallic[idx] = a_metal; p.dL_droughness[idx] = a_rough;
        }
        // per-surfel SH gradients: lane i writes coefficient i (rows are [S,3] / [S])
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (lane == i) {
                if (i < S_inc) {
                    bool on = i < n_inc_g;
                    float* o = p.dL_dincidents_shs + ((size_t)idx * S_inc + i) * 3;
                    o[0] = on ? a_inc[3 * i] : 0.f; o[1] = on ? a_inc[3 * i + 1] : 0.f; o[2] = on ? a_inc[3 * i + 2] : 0.f;
                }
                if (i < S_vis) p.dL_dvisibility_shs[(size_t)idx * S_vis + i] = a_vis[i];
                if (i < S_dir) {
                    atomicAdd(&s_gd[3 * i], a_dir[3 * i]); atomicAdd(&s_gd[3 * i + 1], a_dir[3 * i + 1]);
                    atomicAdd(&s_gd[3 * i + 2], a_dir[3 * i + 2]);
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 3 * S_dir) atomicAdd(&p.dL_ddirect_shs[threadIdx.x], s_gd[threadIdx.x]);
}

}  // namespace
}  // namespace svgir

using namespace svgir;

static int req_validate(const svgir_req_sh_cfg* c) {
    if (!c) { set_error("render_equation_sh: null cfg"); return SVGIR_ERR_INVALID; }
    if (c->P < 0 || c->sample_num <= 0) { set_error("render_equation_sh: bad P=%d sample_num=%d", c->P, c->sample_num); return SVGIR_ERR_INVALID; }
    if (c->S_incident < 0 || c->S_incident > 16 || c->S_direct < 0 || c->S_direct > 16 || c->S_vis < 0 || c->S_vis > 16) {
        set_error("render_equation_sh: SH sizes (%d,%d,%d) must be <= 16 (degree 3)", c->S_incident, c->S_direct, c->S_vis);
        return SVGIR_ERR_INVALID;
    }
    return SVGIR_OK;
}

extern "C" int svgir_render_equation_sh_forward(const svgir_req_sh_cfg* c, const svgir_req_sh_in* in,
                                                const svgir_req_sh_out* out, void* stream) {
    int rc = req_validate(c);
    if (rc) return rc;
    if (c->P == 0) return SVGIR_OK;
    if (!in || !out || !in->base_color || !in->roughness || !in->metallic || !in->normals || !in->viewdirs ||
        (c->S_incident && !in->incidents_shs) || (c->S_direct && !in->direct_shs) || (c->S_vis && !in->visibility_shs) ||
        !out->pbr || !out->incident_dirs || !out->diffuse_light) {
        set_error("render_equation_sh_forward: null pointer"); return SVGIR_ERR_INVALID;
    }
    bool cx = out->incident_lights || out->local_incident_lights || out->global_incident_lights || out->incident_visibility ||
              out->local_diffuse_light || out->accum || out->rgb_d || out->rgb_s;
    if (cx && !(out->incident_lights && out->local_incident_lights && out->global_incident_lights && out->incident_visibility &&
                out->local_diffuse_light && out->accum && out->rgb_d && out->rgb_s)) {
        set_error("render_equation_sh_forward: the _complex outputs must be given together"); return SVGIR_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    FwdPtrs p{in->base_color, in->roughness, in->metallic, in->normals, in->viewdirs, in->incidents_shs, in->direct_shs,
              in->visibility_shs, in->rand_float, out->pbr, out->incident_dirs, out->diffuse_light, out->incident_lights,
              out->local_incident_lights, out->global_incident_lights, out->incident_visibility, out->local_diffuse_light,
              out->accum, out->rgb_d, out->rgb_s};
    int blocks = (int)min((long long)(c->P + 7) / 8, (long long)148 * 16);
    {
        TimedScope ts_("req_sh_fwd", s);
        if (cx) req_sh_forward_kernel<true><<<blocks, 256, 0, s>>>(c->P, c->S_incident, c->S_direct, c->S_vis, c->sample_num, 0, p);
        else req_sh_forward_kernel<false><<<blocks, 256, 0, s>>>(c->P, c->S_incident, c->S_direct, c->S_vis, c->sample_num,
                                                                  c->is_training, p);
    }
    return check_launch("req_sh_fwd", c->debug, s);
}

extern "C" int svgir_render_equation_sh_backward(const svgir_req_sh_cfg* c, const svgir_req_sh_in* in,
                                                 const svgir_req_sh_grads* g, void* stream) {
    int rc = req_validate(c);
    if (rc) return rc;
    if (c->P == 0) return SVGIR_OK;
    if (!in || !g || !in->base_color || !in->roughness || !in->metallic || !in->normals || !in->viewdirs ||
        !g->incident_dirs || !g->dL_dpbr || !g->dL_ddiffuse_light || !g->dL_dbase_color || !g->dL_droughness ||
        !g->dL_dmetallic || !g->dL_dnormals || !g->dL_dviewdirs || (c->S_incident && !g->dL_dincidents_shs) ||
        (c->S_direct && !g->dL_ddirect_shs) || (c->S_vis && !g->dL_dvisibility_shs)) {
        set_error("render_equation_sh_backward: null pointer"); return SVGIR_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    BwdPtrs p{in->base_color, in->roughness, in->metallic, in->normals, in->viewdirs, in->incidents_shs, in->direct_shs,
              in->visibility_shs, g->incident_dirs, g->dL_dpbr, g->dL_ddiffuse_light, g->dL_dbase_color, g->dL_droughness,
              g->dL_dmetallic, g->dL_dnormals, g->dL_dviewdirs, g->dL_dincidents_shs, g->dL_ddirect_shs, g->dL_dvisibility_shs};
    if (c->S_direct && cudaMemsetAsync(g->dL_ddirect_shs, 0, sizeof(float) * 3 * c->S_direct, s) != cudaSuccess) {
        set_error("render_equation_sh_backward: memset failed"); return SVGIR_ERR_CUDA;
    }
    int blocks = (int)min((long long)(c->P + 3) / 4, (long long)148 * 16);
    {
        TimedScope ts_("req_sh_bwd", s);
        req_sh_backward_kernel<<<blocks, 128, 0, s>>>(c->P, c->S_incident, c->S_direct, c->S_vis, c->sample_num,
                                                      c->legacy_exact, p);
    }
    return check_launch("req_sh_bwd", c->debug, s);
}
