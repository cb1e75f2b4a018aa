// Radiance cache and radiance-consistency loss: the B200-native equivalent of the reference's
// Slang kernels render_radiance_with_sampling_SH / gs_bvh_hit / ellipse_hit and render_irradiance_sample
// (pbgi/bvhworkers/intersect_test.slang) with shading_brdf_simple (pbgi/bvhworkers/pbr.slang:283-330).
//
// How it differs in HOW:
//  * no second tree: the closest-hit query walks the packed traversal records svgir_bvh_build wrote for the
//    visibility trace (two child boxes per 64-byte record, near child first, stack in shared memory);
//  * one 128-byte record per surfel holds what a leaf test needs (the reference gathers centre, scale, quaternion,
//    normal, opacity and Sigma^-1 from six tensors and rebuilds the rotation matrix, a covariance it never uses and
//    an orthonormal basis it never uses, per leaf visit);
//  * rays of one surfel are adjacent lanes (same origin, 64 directions): they share the upper tree and the leaves;
//  * the loss is ONE forward and ONE backward kernel, a warp per surfel: the arg-max over the incident samples, the
//    env lookup (the reference materialises direct_light(dirs) * areas for all P*S directions every iteration and
//    back-propagates through all of it), the four-vertex BRDF blend, the sum over the secondary samples and the L1
//    term are fused; only the surfels that are hit touch the env map.
#include <cfloat>
#include "common.cuh"
#include "env.cuh"

namespace svgir {

// ---- surfel records ----------------------------------------------------------------------------------------
// rec[0] = c.xyz, opacity | rec[1] = R[:,0], sx | rec[2] = R[:,1], sy | rec[3] = R[:,2], 0 | rec[4] = normalize(n), 0
// rec[5] = S0..S3 | rec[6] = S4, S5, 0, 0 | rec[7] = 0
__global__ void radiance_pack_kernel(int P, const float* __restrict__ means, const float* __restrict__ scales, int sstride,
                                     const float* __restrict__ rot, const float* __restrict__ normals,
                                     const float* __restrict__ opac, const float* __restrict__ cinv,
                                     float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float r0 = rot[4 * i], r1 = rot[4 * i + 1], r2 = rot[4 * i + 2], r3 = rot[4 * i + 3];
    // matrixFromRotationQuaternions (intersect_test.slang:224-249)
    const float norm = sqrtf(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3 + 0.00000001f);
    const float r = r0 / norm, x = r1 / norm, y = r2 / norm, z = r3 / norm;
    const float m00 = 1 - 2 * (y * y + z * z), m01 = 2 * (x * y - r * z), m02 = 2 * (x * z + r * y);
    const float m10 = 2 * (x * y + r * z), m11 = 1 - 2 * (x * x + z * z), m12 = 2 * (y * z - r * x);
    const float m20 = 2 * (x * z - r * y), m21 = 2 * (y * z + r * x), m22 = 1 - 2 * (x * x + y * y);
    float nx = normals[3 * i], ny = normals[3 * i + 1], nz = normals[3 * i + 2];
    const float nl = sqrtf(nx * nx + ny * ny + nz * nz);
    nx /= nl; ny /= nl; nz /= nl;
    const float* c = cinv + (size_t)i * 6;
    float4* o = out + (size_t)i * 8;
    o[0] = make_float4(means[3 * i], means[3 * i + 1], means[3 * i + 2], opac[i]);
    o[1] = make_float4(m00, m10, m20, scales[(size_t)i * sstride]);
    o[2] = make_float4(m01, m11, m21, scales[(size_t)i * sstride + 1]);
    o[3] = make_float4(m02, m12, m22, 0.f);
    o[4] = make_float4(nx, ny, nz, 0.f);
    o[5] = make_float4(c[0], c[1], c[2], c[3]);
    o[6] = make_float4(c[4], c[5], 0.f, 0.f);
    o[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- closest hit -------------------------------------------------------------------------------------------
static constexpr int RC_THREADS = 128;
static constexpr int RC_STACK = 64;

struct Hit { float t; int index; float u, v, alpha; };

// aabb_hit (intersect_test.slang:20-43): slab test against [t_min, t_max]; returns the entry distance or -1. The boxes
// here are the tight 8-corner boxes of the visibility tree (a surfel with a vanishing third scale has a box of
// vanishing thickness), so the rejection keeps a 1e-6 margin; a conservative box test cannot change a result, every
// candidate goes through the exact leaf test.
__device__ __forceinline__ float slab_entry(float lx, float ly, float lz, float ux, float uy, float uz, float ox, float oy,
                                            float oz, float ix, float iy, float iz, float t_min, float t_max) {
    float t0 = (lx - ox) * ix, t1 = (ux - ox) * ix;
    if (ix < 0.f) { const float t = t0; t0 = t1; t1 = t; }
    t_min = t0 > t_min ? t0 : t_min; t_max = t1 < t_max ? t1 : t_max;
    if (t_max + 1e-6f < t_min) return -1.f;
    t0 = (ly - oy) * iy; t1 = (uy - oy) * iy;
    if (iy < 0.f) { const float t = t0; t0 = t1; t1 = t; }
    t_min = t0 > t_min ? t0 : t_min; t_max = t1 < t_max ? t1 : t_max;
    if (t_max + 1e-6f < t_min) return -1.f;
    t0 = (lz - oz) * iz; t1 = (uz - oz) * iz;
    if (iz < 0.f) { const float t = t0; t0 = t1; t1 = t; }
    t_min = t0 > t_min ? t0 : t_min; t_max = t1 < t_max ? t1 : t_max;
    if (t_max + 1e-6f < t_min) return -1.f;
    return t_min;
}

// The leaf test of gs_bvh_hit (intersect_test.slang:307-424) = ellipse_hit (:94-149) followed by the power, alpha and
// facing rejections, in the reference's order. A hit replaces the current one only when it is strictly closer (:409).
__device__ __forceinline__ void surfel_test(const float4* __restrict__ rec, int obj, float ox, float oy, float oz, float dx,
                                            float dy, float dz, float t_min, Hit& h) {
    const float4* r = rec + (size_t)obj * 8;
    const float4 c = __ldg(r), nw = __ldg(r + 3);
    const float denom = nw.x * dx + nw.y * dy + nw.z * dz;
    if (fabsf(denom) < 1e-6f) return;
    const float t = ((c.x - ox) * nw.x + (c.y - oy) * nw.y + (c.z - oz) * nw.z) / denom;
    if (!(t >= t_min) || !(t < h.t)) return;
    const float px = ox + t * dx, py = oy + t * dy, pz = oz + t * dz;
    const float4 ua = __ldg(r + 1), va = __ldg(r + 2);
    const float ex = px - c.x, ey = py - c.y, ez = pz - c.z;
    const float a = ua.x * ex + ua.y * ey + ua.z * ez;       // posM = R^T (pos - centre)
    const float b = va.x * ex + va.y * ey + va.z * ez;
    const float disM = (a * a) / (ua.w * ua.w) + (b * b) / (va.w * va.w);
    if (!(disM <= 9.0f)) return;
    const float4 s0 = __ldg(r + 5), s1 = __ldg(r + 6);
    // gaussian_fn (:189-196) with d = centre - pos
    const float power = -0.5f * (ex * ex * s0.x + ey * ey * s0.w + ez * ez * s1.y + 2 * ex * ey * s0.y + 2 * ex * ez * s0.z +
                                 2 * ey * ez * s1.x);
    if (power > 0.0f) return;
    const float alpha = fminf(0.99f, c.w * expf(power));
    if (alpha < 1.0f / 255.0f) return;
    const float4 n = __ldg(r + 4);
    if (!(dx * n.x + dy * n.y + dz * n.z < -0.0f)) return;
    float u = a / ua.w, v = b / va.w;
    if (u < v) { const float tmp = u; u = v; v = tmp; }       // :126-129
    h.t = t; h.index = obj; h.alpha = alpha;
    h.u = fminf(fmaxf(u * 0.5f + 0.5f, 0.001f), 0.999f);
    h.v = fminf(fmaxf(v * 0.5f + 0.5f, 0.001f), 0.999f);
}

// eval_sh (pbgi/bvhworkers/sh_utils.slang): degree 3, +0.5, no clamp. sh [16][3] of one surfel.
__device__ __forceinline__ void eval_sh3(const float* __restrict__ sh, float x, float y, float z, float out[3]) {
    const float il = 1.f / sqrtf(x * x + y * y + z * z);
    x *= il; y *= il; z *= il;
    const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
    const float xx = x * x, yy = y * y, zz = z * z;
    float w[16];
    w[0] = C0; w[1] = -C1 * y; w[2] = C1 * z; w[3] = -C1 * x;
    w[4] = 1.0925484305920792f * x * y; w[5] = -1.0925484305920792f * y * z;
    w[6] = 0.31539156525252005f * (2.0f * zz - xx - yy); w[7] = -1.0925484305920792f * x * z;
    w[8] = 0.5462742152960396f * (xx - yy);
    w[9] = -0.5900435899266435f * y * (3.0f * xx - yy); w[10] = 2.890611442640554f * x * y * z;
    w[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
    w[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    w[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy); w[14] = 1.445305721320277f * z * (xx - yy);
    w[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
    float r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        r = fmaf(w[k], __ldg(sh + 3 * k), r); g = fmaf(w[k], __ldg(sh + 3 * k + 1), g); b = fmaf(w[k], __ldg(sh + 3 * k + 2), b);
    }
    out[0] = r + 0.5f; out[1] = g + 0.5f; out[2] = b + 0.5f;
}

// One thread per ray, ray = n * S + s.
__global__ void __launch_bounds__(RC_THREADS) radiance_cache_kernel(
    long long n_rays, int S, int P, int first_index, int self_mod, const float4* __restrict__ packed,
    const float4* __restrict__ rec, const float* __restrict__ shs, const float* __restrict__ origins,
    const float* __restrict__ dirs, float* __restrict__ radiance, float* __restrict__ visibility,
    int32_t* __restrict__ hit_index, float* __restrict__ uv) {
    __shared__ int stack[RC_STACK][RC_THREADS];
    const long long ray = (long long)blockIdx.x * RC_THREADS + threadIdx.x;
    if (ray >= n_rays) return;
    const int tid = threadIdx.x;
    const int n = (int)(ray / S);
    float dx = dirs[3 * ray], dy = dirs[3 * ray + 1], dz = dirs[3 * ray + 2];
    { const float il = 1.f / sqrtf(dx * dx + dy * dy + dz * dz); dx *= il; dy *= il; dz *= il; }
    float ox = origins[3 * (size_t)n], oy = origins[3 * (size_t)n + 1], oz = origins[3 * (size_t)n + 2];
    const int self = self_mod > 0 ? (first_index + n) % self_mod : first_index + n;
    // aabb_hit replaces a zero component by 1e-6 before inverting (:26)
    const float ix = 1.0f / (dx == 0.f ? 0.000001f : dx), iy = 1.0f / (dy == 0.f ? 0.000001f : dy),
                iz = 1.0f / (dz == 0.f ? 0.000001f : dz);
    float T = 1.0f, t_min = 0.042f;
    const float t_max = 0.2f;
    float sr = 0.f, sg = 0.f, sb = 0.f;
    bool visible = true;
    int first_hit = -1;
    float fu = 0.f, fv = 0.f;
    for (int iter = 0; iter < 4096 && T > 0.001f; iter++) {
        Hit h;
        h.t = t_max; h.index = -1; h.u = h.v = 0.f; h.alpha = 0.f;
        int sp = 0;
        if (P == 1) surfel_test(rec, 0, ox, oy, oz, dx, dy, dz, t_min, h);
        else stack[sp++][tid] = 0;
        while (sp > 0) {
            const int id = stack[--sp][tid];
            const float4 a = __ldg(packed + (size_t)id * 4), b = __ldg(packed + (size_t)id * 4 + 1);
            const float4 c = __ldg(packed + (size_t)id * 4 + 2), d = __ldg(packed + (size_t)id * 4 + 3);
            const int l = __float_as_int(a.x), rr = __float_as_int(a.y);
            const float lt = slab_entry(a.z, a.w, b.x, b.y, b.z, b.w, ox, oy, oz, ix, iy, iz, t_min, h.t);
            const float rt = slab_entry(c.x, c.y, c.z, c.w, d.x, d.y, ox, oy, oz, ix, iy, iz, t_min, h.t);
            // near child first so that the far one is usually culled by the shrunken interval when it is popped
            const bool l_first = !(lt > rt);
            const int first = l_first ? l : rr, second = l_first ? rr : l;
            const float ft = l_first ? lt : rt, st = l_first ? rt : lt;
            if (ft >= 0.f && first < 0) surfel_test(rec, ~first, ox, oy, oz, dx, dy, dz, t_min, h);
            if (st >= 0.f) {
                if (second < 0) surfel_test(rec, ~second, ox, oy, oz, dx, dy, dz, t_min, h);
                else if (sp < RC_STACK) stack[sp++][tid] = second;
            }
            if (ft >= 0.f && first >= 0 && sp < RC_STACK) stack[sp++][tid] = first;
        }
        // an internal node popped later is re-tested against the current closest hit by the slab test of its
        // children only; that is conservative (never drops a closer hit)
        if (h.index < 0 || h.index == self) break;          // :1932, :1970-1973
        if (first_hit == -1) { first_hit = h.index; fu = h.u; fv = h.v; t_min = 0.01f; }   // :1944-1949
        const float4 hc = __ldg(rec + (size_t)h.index * 8);
        float col[3];
        eval_sh3(shs + (size_t)h.index * 48, hc.x - ox, hc.y - oy, hc.z - oz, col);
        ox += dx * h.t; oy += dy * h.t; oz += dz * h.t;
        const float w = h.alpha * T;                        // (1 - debug_res.x) * test_T, :1961
        sr += col[0] * w; sg += col[1] * w; sb += col[2] * w;
        T *= 1.f - h.alpha;
        if (T < 0.2f) visible = false;
    }
    visibility[ray] = visible ? T : 0.0f;
    radiance[3 * ray] = fminf(fmaxf(sr, 0.0f), 10.0f);
    radiance[3 * ray + 1] = fminf(fmaxf(sg, 0.0f), 10.0f);
    radiance[3 * ray + 2] = fminf(fmaxf(sb, 0.0f), 10.0f);
    hit_index[ray] = first_hit;
    uv[2 * ray] = fu; uv[2 * ray + 1] = fv;
}

// ---- radiance-consistency loss ---------------------------------------------------------------------------------
static constexpr int RL_THREADS = 256;
static constexpr int RL_WPC = RL_THREADS / 32;
// The env-map gradient of ~10^7 samples lands on a few hundred texels: same-address reductions serialise in L2, so the
// accumulator is replicated and each CTA scatters into copy blockIdx % ENV_COPIES; the finalize kernel folds the copies.
static constexpr int ENV_COPIES = SVGIR_RADIANCE_ENV_COPIES;

struct RLArgs {
    int P, S, He, We, env_mode, rough_stride, ref_grid, nrm_vmajor;
    float env_scale;
    const float *means3D, *campos, *geo_normal, *dirs, *areas, *vis, *uv, *radiances, *ratio, *normals, *albedo,
        *roughness, *env_act, *env_param;
    const int32_t* hit;
    const int32_t* skip_flag;
    const float* taps;
};

// shading_brdf_simple (pbr.slang:283-330): specular part and, optionally, its derivative w.r.t. roughness.
// V, L and the vertex normal (nx,ny,nz) normalised.
template <bool GRAD>
__device__ __forceinline__ float brdf_specular(float Vx, float Vy, float Vz, float Lx, float Ly, float Lz, float Hx, float Hy,
                                               float Hz, float VoH, float nx, float ny, float nz, float rough,
                                               float& dspec_dr) {
    const float NoL = fminf(fmaxf(nx * Lx + ny * Ly + nz * Lz, 1e-6f), 1.f);
    const float NoV = fminf(fmaxf(nx * Vx + ny * Vy + nz * Vz, 1e-6f), 1.f);
    const float NoH = fminf(fmaxf(nx * Hx + ny * Hy + nz * Hz, 1e-6f), 1.f);
    const float alpha = rough * rough, alpha2 = alpha * alpha;
    const float k = (alpha + 2.0f * rough + 1.0f) / 8.0f;
    const float FMi = (-5.55473f * VoH - 6.98316f) * VoH;
    const float F = 0.04f + (1 - 0.04f) * exp2f(FMi);
    const float frac = F * alpha2;
    const float nom0 = NoH * NoH * (alpha2 - 1.0f) + 1.0f;
    const float nom1 = NoV * (1.0f - k) + k, nom2 = NoL * (1.0f - k) + k;
    const float FOUR_PI = 4.f * PI_F;
    const float nom_raw = FOUR_PI * nom0 * nom0 * nom1 * nom2;
    const float nom = fminf(fmaxf(nom_raw, 1e-6f), FOUR_PI);
    if (GRAD) {
        const float r3 = 4.f * rough * rough * rough;           // d alpha2 / d r
        const float dk = (rough + 1.0f) * 0.25f;                // d k / d r
        const float dfrac = F * r3;
        float dnom = 0.f;
        if (nom_raw >= 1e-6f && nom_raw <= FOUR_PI)
            dnom = FOUR_PI * (2.f * nom0 * (NoH * NoH * r3) * nom1 * nom2 + nom0 * nom0 * ((1.f - NoV) * dk) * nom2 +
                              nom0 * nom0 * nom1 * ((1.f - NoL) * dk));
        dspec_dr = dfrac / nom - frac / (nom * nom) * dnom;
    }
    return frac / nom;
}

__device__ __forceinline__ float nan_to_num_f(float x) {
    if (x != x) return 0.f;
    if (x == INFINITY) return FLT_MAX;
    if (x == -INFINITY) return -FLT_MAX;
    return x;
}

// Selection of the primary sample (gaussian_model.py:549-565), a warp per surfel: first index of the maximum of
// dot(dirs[n,s], view_reflect) * (1 - visibility[n,s]). Everything the irradiance kernels need about surfel n goes into
// two float4: saved[2n] = { V = normalize(-dirs[n,sel]), bits(hit[n,sel]) }, saved[2n+1] = { target rgb, bits(sel) },
// target = nan_to_num(radiances[n,sel] * ratio) (:323-324, :573).
__global__ void __launch_bounds__(RL_THREADS) radiance_select_kernel(const RLArgs a, int32_t* __restrict__ sample_index,
                                                                     float4* __restrict__ saved) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * RL_WPC + (threadIdx.x >> 5);
    if (n >= a.P) return;
    const float cx = __ldg(a.campos), cy = __ldg(a.campos + 1), cz = __ldg(a.campos + 2);
    float vx = __ldg(a.means3D + 3 * (size_t)n) - cx, vy = __ldg(a.means3D + 3 * (size_t)n + 1) - cy,
          vz = __ldg(a.means3D + 3 * (size_t)n + 2) - cz;
    const float vl = fmaxf(sqrtf(vx * vx + vy * vy + vz * vz), 1e-12f);   // F.normalize
    vx /= vl; vy /= vl; vz /= vl;
    const float gx = __ldg(a.geo_normal + 3 * (size_t)n), gy = __ldg(a.geo_normal + 3 * (size_t)n + 1),
                gz = __ldg(a.geo_normal + 3 * (size_t)n + 2);
    const float d2 = 2.f * (gx * vx + gy * vy + gz * vz);
    const float rx = d2 * gx + vx, ry = d2 * gy + vy, rz = d2 * gz + vz;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int s = lane; s < a.S; s += 32) {
        const float* d = a.dirs + ((size_t)n * a.S + s) * 3;
        const float v = (__ldg(d) * rx + __ldg(d + 1) * ry + __ldg(d + 2) * rz) * (1.f - __ldg(a.vis + (size_t)n * a.S + s));
        if (v > best || bi == 0x7fffffff) { best = v; bi = s; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
    }
    if (lane == 0) {
        const int sel = bi;
        const int h = __ldg(a.hit + (size_t)n * a.S + sel);
        const float* pd = a.dirs + ((size_t)n * a.S + sel) * 3;
        float Vx = -__ldg(pd), Vy = -__ldg(pd + 1), Vz = -__ldg(pd + 2);
        const float il = 1.f / sqrtf(Vx * Vx + Vy * Vy + Vz * Vz);
        const float ratio = a.ratio ? __ldg(a.ratio) : 1.f;
        const float* rp = a.radiances + ((size_t)n * a.S + sel) * 3;
        saved[2 * (size_t)n] = make_float4(Vx * il, Vy * il, Vz * il, __int_as_float(h));
        saved[2 * (size_t)n + 1] = make_float4(nan_to_num_f(__ldg(rp) * ratio), nan_to_num_f(__ldg(rp + 1) * ratio),
                                               nan_to_num_f(__ldg(rp + 2) * ratio), __int_as_float(sel));
        if (sample_index) sample_index[n] = sel;
    }
}

struct HitSurfel { float nrm[12], alb[12], rough; };
__device__ __forceinline__ void load_hit_surfel(const RLArgs& a, int h, HitSurfel& hs) {
    const float4* np = reinterpret_cast<const float4*>(a.normals + (size_t)h * 12);
    const float4* ap = reinterpret_cast<const float4*>(a.albedo + (size_t)h * 12);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float4 nv = __ldg(np + k), av = __ldg(ap + k);
        if (a.nrm_vmajor) {   // flat element e = 4k..4k+3 of [v][c] goes to slot 4*(e % 3) + e / 3
            const float e4[4] = {nv.x, nv.y, nv.z, nv.w};
#pragma unroll
            for (int j = 0; j < 4; j++) hs.nrm[4 * ((4 * k + j) % 3) + (4 * k + j) / 3] = e4[j];
        } else {
            hs.nrm[4 * k] = nv.x; hs.nrm[4 * k + 1] = nv.y; hs.nrm[4 * k + 2] = nv.z; hs.nrm[4 * k + 3] = nv.w;
        }
        hs.alb[4 * k] = av.x; hs.alb[4 * k + 1] = av.y; hs.alb[4 * k + 2] = av.z; hs.alb[4 * k + 3] = av.w;
    }
    hs.rough = __ldg(a.roughness + (size_t)h * a.rough_stride);
#pragma unroll
    for (int v = 0; v < 4; v++) {   // normalize(normal) of shading_brdf_simple, once per surfel instead of once per sample
        const float il = rsqrtf(hs.nrm[v] * hs.nrm[v] + hs.nrm[4 + v] * hs.nrm[4 + v] + hs.nrm[8 + v] * hs.nrm[8 + v]);
        hs.nrm[v] *= il; hs.nrm[4 + v] *= il; hs.nrm[8 + v] *= il;
    }
}

// Raw inputs of one secondary sample of render_irradiance_sample (:1212-1303): loaded unconditionally so that all the
// gathers of a surfel are in flight together.
struct SecRaw { int hit2; float rx, ry, rz, u, v, area, t0, t1, t2; };
__device__ __forceinline__ void load_secondary_raw(const RLArgs& a, int h, int s2, SecRaw& r) {
    const size_t i = (size_t)h * a.S + (s2 < a.S ? s2 : 0);
    r.hit2 = s2 < a.S ? __ldg(a.hit + i) : 0;           // anything but -1 closes the sample
    const float* d = a.dirs + i * 3;
    r.rx = __ldg(d); r.ry = __ldg(d + 1); r.rz = __ldg(d + 2);
    const float2 uv = __ldg(reinterpret_cast<const float2*>(a.uv) + i);
    r.u = uv.x; r.v = uv.y;
    r.area = __ldg(a.areas + i);
    if (a.taps) { const float* tp = a.taps + i * 3; r.t0 = __ldg(tp); r.t1 = __ldg(tp + 1); r.t2 = __ldg(tp + 2); }
}

struct SecSample { float Lx, Ly, Lz, Hx, Hy, Hz, VoH, w[4], E[3]; EnvTap tap; float escale; };
__device__ __forceinline__ void make_secondary(const RLArgs& a, const SecRaw& r, float Vx, float Vy, float Vz, SecSample& q) {
    const float il = rsqrtf(r.rx * r.rx + r.ry * r.ry + r.rz * r.rz);
    q.Lx = r.rx * il; q.Ly = r.ry * il; q.Lz = r.rz * il;
    float hx = Vx + q.Lx, hy = Vy + q.Ly, hz = Vz + q.Lz;
    const float hl = rsqrtf(hx * hx + hy * hy + hz * hz);
    q.Hx = hx * hl; q.Hy = hy * hl; q.Hz = hz * hl;
    q.VoH = fminf(fmaxf(Vx * q.Hx + Vy * q.Hy + Vz * q.Hz, 1e-6f), 1.f);
    q.w[0] = (1 - r.u) * (1 - r.v); q.w[1] = r.u * (1 - r.v); q.w[2] = (1 - r.u) * r.v; q.w[3] = r.u * r.v;
    // envmap[h,s2] = direct_light(dirs[h,s2]) * areas[h,s2] (gaussian_model.py:547), on the raw direction
    q.tap = a.taps ? env_tap_unpack(r.t0, r.t1, r.t2) : env_coords(r.rx, r.ry, r.rz, a.He, a.We);
    q.escale = a.env_scale * r.area;
    float e[3];
    env_fetch(a.env_act, a.He, a.We, q.tap, e);
    q.E[0] = e[0] * q.escale; q.E[1] = e[1] * q.escale; q.E[2] = e[2] * q.escale;
}

// Gradient of one surfel's irradiance (G = d loss / d irradiance[n] / S) w.r.t. the hit surfel's albedo and roughness
// and the env map: the secondary samples are evaluated again (their rows are in L1/L2 from the forward pass).
__device__ __forceinline__ void radiance_surfel_grads(const RLArgs& a, int h, float Vx, float Vy, float Vz, const float (&G)[3],
                                                      const HitSurfel& hs, int lane, float* __restrict__ d_albedo,
                                                      float* __restrict__ d_roughness, float* __restrict__ env_copy) {
    float acc[13];                                   // 12 albedo entries (4*c + v) + roughness
#pragma unroll
    for (int k = 0; k < 13; k++) acc[k] = 0.f;
    // reference-grid mode: every one of the S backward threads differentiates secondary sample 0
    const int s_end = a.ref_grid ? 1 : a.S;
    const float rep = a.ref_grid ? (float)a.S : 1.f;
#pragma unroll 2
    for (int s2 = lane; s2 < s_end; s2 += 32) {
        SecRaw raw;
        load_secondary_raw(a, h, s2, raw);
        if (raw.hit2 != -1) continue;
        SecSample q;
        make_secondary(a, raw, Vx, Vy, Vz, q);
        float spec = 0.f, dspec = 0.f, mix[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int v = 0; v < 4; v++) {
            float ds;
            const float sp = brdf_specular<true>(Vx, Vy, Vz, q.Lx, q.Ly, q.Lz, q.Hx, q.Hy, q.Hz, q.VoH, hs.nrm[v],
                                                 hs.nrm[4 + v], hs.nrm[8 + v], hs.rough, ds);
            spec += q.w[v] * sp;
            dspec += q.w[v] * ds;
#pragma unroll
            for (int c = 0; c < 3; c++) mix[c] += q.w[v] * (hs.alb[4 * c + v] * (1.f / PI_F));
        }
        float ge[3], gsum = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float g = G[c] * rep;
            const float gE = g * q.E[c];
#pragma unroll
            for (int v = 0; v < 4; v++) acc[4 * c + v] += gE * q.w[v] * (1.f / PI_F);
            gsum += gE;
            ge[c] = g * (spec + mix[c]) * q.escale;    // d loss / d (bilinear env value)
        }
        acc[12] += gsum * dspec;
        if (env_copy) {
            const float wx0 = 1.f - q.tap.wx1, wy0 = 1.f - q.tap.wy1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int x = q.tap.x0 + (k & 1), y = q.tap.y0 + (k >> 1);
                if (x < 0 || x > a.We - 1 || y < 0 || y > a.He - 1) continue;
                const float w = ((k & 1) ? q.tap.wx1 : wx0) * ((k >> 1) ? q.tap.wy1 : wy0);
                red_add_v4(env_copy + ((size_t)y * a.We + x) * 4, ge[0] * w, ge[1] * w, ge[2] * w);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 13; k++) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    float mine = 0.f;   // lane k adds entry k
#pragma unroll
    for (int k = 0; k < 13; k++) if (lane == k) mine = acc[k];
    if (lane < 12) { if (d_albedo && mine != 0.f) atomicAdd(d_albedo + (size_t)h * 12 + lane, mine); }
    else if (lane == 12) { if (d_roughness && mine != 0.f) atomicAdd(d_roughness + (size_t)h * a.rough_stride, mine); }
}

// sign(irradiance - target) * grad / numel, and the 1/S of :1301
__device__ __forceinline__ bool radiance_upstream(const float (&irr)[3], const float4& B, float gscale, float inv_S, float (&G)[3]) {
    const float d0 = irr[0] - B.x, d1 = irr[1] - B.y, d2 = irr[2] - B.z;
    G[0] = (d0 > 0.f ? gscale : (d0 < 0.f ? -gscale : 0.f)) * inv_S;
    G[1] = (d1 > 0.f ? gscale : (d1 < 0.f ? -gscale : 0.f)) * inv_S;
    G[2] = (d2 > 0.f ? gscale : (d2 < 0.f ? -gscale : 0.f)) * inv_S;
    return G[0] != 0.f || G[1] != 0.f || G[2] != 0.f;
}

// WITH_GRAD: forward and backward in one pass (the step knows the upstream gradient beforehand: lambda_radiance).
template <bool WITH_GRAD>
__global__ void __launch_bounds__(RL_THREADS, WITH_GRAD ? 2 : 3) radiance_loss_fwd_kernel(
    const RLArgs a, const float4* __restrict__ saved, float* __restrict__ irradiance, float* __restrict__ partials,
    const float* __restrict__ grad_loss, float* __restrict__ d_albedo, float* __restrict__ d_roughness,
    float* __restrict__ d_env_acc) {
    __shared__ float wsum[RL_WPC];
    const bool want_grad = WITH_GRAD && !(a.skip_flag && __ldg(a.skip_flag) != 0);
    const float gscale = WITH_GRAD ? (grad_loss ? __ldg(grad_loss) : 1.f) / (3.f * (float)a.P) : 0.f;
    float* env_copy = (WITH_GRAD && d_env_acc) ? d_env_acc + (size_t)(blockIdx.x % ENV_COPIES) * a.He * a.We * 4 : nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_S = 1.f / (float)a.S;
    const int stride = gridDim.x * RL_WPC;
    float loss_acc = 0.f;
    int n = blockIdx.x * RL_WPC + warp;
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A;
    if (n < a.P) { A = __ldg(saved + 2 * (size_t)n); B = __ldg(saved + 2 * (size_t)n + 1); }
    while (n < a.P) {
        const int n_next = n + stride;
        float4 An = A, Bn = B;
        if (n_next < a.P) { An = __ldg(saved + 2 * (size_t)n_next); Bn = __ldg(saved + 2 * (size_t)n_next + 1); }
        const int h = __float_as_int(A.w);
        float irr[3] = {0.f, 0.f, 0.f};
        if (h != -1) {
            const float Vx = A.x, Vy = A.y, Vz = A.z;
            HitSurfel hs;
            load_hit_surfel(a, h, hs);
#pragma unroll 2
            for (int s2 = lane; s2 < a.S; s2 += 32) {
                SecRaw raw;
                load_secondary_raw(a, h, s2, raw);
                if (raw.hit2 != -1) continue;
                SecSample q;
                make_secondary(a, raw, Vx, Vy, Vz, q);
                float spec = 0.f, mix[3] = {0.f, 0.f, 0.f}, dummy;
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    const float sp = brdf_specular<false>(Vx, Vy, Vz, q.Lx, q.Ly, q.Lz, q.Hx, q.Hy, q.Hz, q.VoH, hs.nrm[v],
                                                          hs.nrm[4 + v], hs.nrm[8 + v], hs.rough, dummy);
                    spec += q.w[v] * sp;
#pragma unroll
                    for (int c = 0; c < 3; c++) mix[c] += q.w[v] * (hs.alb[4 * c + v] * (1.f / PI_F));
                }
#pragma unroll
                for (int c = 0; c < 3; c++) irr[c] += (spec + mix[c]) * q.E[c] * inv_S;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int c = 0; c < 3; c++) irr[c] += __shfl_xor_sync(0xffffffffu, irr[c], o);
            }
            if (WITH_GRAD) {
                float G[3];
                if (radiance_upstream(irr, B, gscale, inv_S, G) && want_grad)
                    radiance_surfel_grads(a, h, Vx, Vy, Vz, G, hs, lane, d_albedo, d_roughness, env_copy);
            }
        }
        if (lane == 0) {
            irradiance[3 * (size_t)n] = irr[0]; irradiance[3 * (size_t)n + 1] = irr[1]; irradiance[3 * (size_t)n + 2] = irr[2];
            loss_acc += fabsf(irr[0] - B.x) + fabsf(irr[1] - B.y) + fabsf(irr[2] - B.z);
        }
        n = n_next; A = An; B = Bn;
    }
    if (lane == 0) wsum[warp] = loss_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < RL_WPC; w++) s += wsum[w];
        partials[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) radiance_loss_reduce_kernel(int nblocks, float scale, const float* __restrict__ partials,
                                                                   float* __restrict__ loss) {
    __shared__ float sh[256];
    float s = 0.f;
    for (int i = threadIdx.x; i < nblocks; i += 256) s += partials[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = sh[0] * scale;
}

__global__ void __launch_bounds__(RL_THREADS, 2) radiance_loss_bwd_kernel(const RLArgs a, const float* __restrict__ grad_loss,
                                                                          const float4* __restrict__ saved,
                                                                          const float* __restrict__ irradiance,
                                                                          float* __restrict__ d_albedo,
                                                                          float* __restrict__ d_roughness,
                                                                          float* __restrict__ d_env_acc) {
    if (a.skip_flag && __ldg(a.skip_flag) != 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_S = 1.f / (float)a.S;
    const float gscale = (grad_loss ? __ldg(grad_loss) : 1.f) / (3.f * (float)a.P);
    const int stride = gridDim.x * RL_WPC;
    float* env_copy = d_env_acc ? d_env_acc + (size_t)(blockIdx.x % ENV_COPIES) * a.He * a.We * 4 : nullptr;
    int n = blockIdx.x * RL_WPC + warp;
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A;
    float I[3] = {0.f, 0.f, 0.f};
    if (n < a.P) {
        A = __ldg(saved + 2 * (size_t)n); B = __ldg(saved + 2 * (size_t)n + 1);
        I[0] = __ldg(irradiance + 3 * (size_t)n); I[1] = __ldg(irradiance + 3 * (size_t)n + 1); I[2] = __ldg(irradiance + 3 * (size_t)n + 2);
    }
    while (n < a.P) {
        const int n_next = n + stride;
        float4 An = A, Bn = B;
        float J0 = 0.f, J1 = 0.f, J2 = 0.f;
        if (n_next < a.P) {
            An = __ldg(saved + 2 * (size_t)n_next); Bn = __ldg(saved + 2 * (size_t)n_next + 1);
            J0 = __ldg(irradiance + 3 * (size_t)n_next); J1 = __ldg(irradiance + 3 * (size_t)n_next + 1);
            J2 = __ldg(irradiance + 3 * (size_t)n_next + 2);
        }
        const int h = __float_as_int(A.w);
        float G[3];
        if (radiance_upstream(I, B, gscale, inv_S, G) && h != -1) {
            HitSurfel hs;
            load_hit_surfel(a, h, hs);
            radiance_surfel_grads(a, h, A.x, A.y, A.z, G, hs, lane, d_albedo, d_roughness, env_copy);
        }
        n = n_next; A = An; B = Bn; I[0] = J0; I[1] = J1; I[2] = J2;
    }
}

static int rl_prepare(const svgir_radiance_loss_cfg* c, const svgir_radiance_loss_in* in, RLArgs& a, cudaStream_t s) {
    if (!c || !in || c->P < 0 || c->S <= 0 || c->env_h <= 0 || c->env_w <= 0 || c->rough_stride < 1) { set_error("radiance_loss: bad cfg"); return SVGIR_ERR_INVALID; }
    if (!in->means3D || !in->campos || !in->geo_normal || !in->incident_dirs || !in->incident_areas || !in->visibility ||
        !in->hit_index || !in->uv || !in->radiances || !in->normals || !in->albedo || !in->roughness || !in->env ||
        !in->env_act_scratch) { set_error("radiance_loss: missing input"); return SVGIR_ERR_INVALID; }
    if (((uintptr_t)in->normals & 15) || ((uintptr_t)in->albedo & 15)) { set_error("radiance_loss: normals / albedo must be 16-byte aligned"); return SVGIR_ERR_INVALID; }
    const int nenv = c->env_h * c->env_w * 3;
    if (!(c->flags & SVGIR_RADIANCE_ENV_READY)) launch_env_activate(nenv, in->env, in->env_act_scratch, c->env_mode, s);
    a.P = c->P; a.S = c->S; a.He = c->env_h; a.We = c->env_w; a.env_mode = c->env_mode; a.rough_stride = c->rough_stride;
    a.ref_grid = (c->flags & SVGIR_RADIANCE_BWD_REFERENCE_GRID) ? 1 : 0;
    a.nrm_vmajor = (c->flags & SVGIR_RADIANCE_NORMALS_VERTEX_MAJOR) ? 1 : 0;
    a.env_scale = c->env_mode == 0 ? 2.0f : 1.0f;
    a.means3D = in->means3D; a.campos = in->campos; a.geo_normal = in->geo_normal; a.dirs = in->incident_dirs;
    a.areas = in->incident_areas; a.vis = in->visibility; a.uv = in->uv; a.radiances = in->radiances;
    a.ratio = in->radiance_ratio; a.normals = in->normals; a.albedo = in->albedo; a.roughness = in->roughness;
    a.env_act = in->env_act_scratch; a.env_param = in->env; a.hit = in->hit_index; a.skip_flag = in->skip_flag; a.taps = in->env_taps;
    return SVGIR_OK;
}

static int rl_grid(int P) {
    const int need = (P + RL_WPC - 1) / RL_WPC;
    const int cap = 148 * 8;
    return need < cap ? (need > 0 ? need : 1) : cap;
}

}  // namespace svgir

using namespace svgir;

extern "C" {

int svgir_radiance_pack_surfels(int P, const float* means3D, const float* scales, int scale_stride, const float* rotations,
                                const float* normals, const float* opacity, const float* symm_inv, float* records,
                                void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return SVGIR_OK;
    if (!means3D || !scales || !rotations || !normals || !opacity || !symm_inv || !records) { set_error("radiance_pack_surfels: null pointer"); return SVGIR_ERR_INVALID; }
    if (scale_stride < 2) { set_error("radiance_pack_surfels: scale_stride < 2"); return SVGIR_ERR_INVALID; }
    if ((uintptr_t)records & 15) { set_error("radiance_pack_surfels: records must be 16-byte aligned"); return SVGIR_ERR_INVALID; }
    { TimedScope t_("radiance_pack", s);
      radiance_pack_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, scales, scale_stride, rotations, normals, opacity, symm_inv,
                                                           (float4*)records); }
    return check_launch("radiance_pack_surfels", false, s);
}

int svgir_radiance_cache_build(const svgir_bvh* b, int N, int S, int first_index, int self_mod, const float* origins,
                               const float* dirs, const float* records, const float* shs, float* radiance,
                               float* visibility, int32_t* hit_index, float* uv, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (N <= 0 || S <= 0) return SVGIR_OK;
    if (!b || b->P <= 0 || !b->packed) { set_error("radiance_cache_build: tree missing"); return SVGIR_ERR_INVALID; }
    if (!origins || !dirs || !records || !shs || !radiance || !visibility || !hit_index || !uv) { set_error("radiance_cache_build: null pointer"); return SVGIR_ERR_INVALID; }
    if (self_mod < 0 || first_index < 0) { set_error("radiance_cache_build: negative first_index / self_mod"); return SVGIR_ERR_INVALID; }
    const long long n_rays = (long long)N * S;
    const long long blocks = (n_rays + RC_THREADS - 1) / RC_THREADS;
    if (blocks > 0x7fffffffLL) { set_error("radiance_cache_build: too many rays for one launch"); return SVGIR_ERR_INVALID; }
    { TimedScope t_("radiance_cache", s);
      radiance_cache_kernel<<<(unsigned)blocks, RC_THREADS, 0, s>>>(n_rays, S, b->P, first_index, self_mod, (const float4*)b->packed,
                                                                    (const float4*)records, shs, origins, dirs, radiance,
                                                                    visibility, hit_index, uv); }
    return check_launch("radiance_cache_build", false, s);
}

int svgir_radiance_loss_forward(const svgir_radiance_loss_cfg* c, const svgir_radiance_loss_in* in, float* loss,
                                float* irradiance, int32_t* sample_index, float* saved, float* scratch, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    RLArgs a;
    int rc = rl_prepare(c, in, a, s);
    if (rc) return rc;
    if (!loss || !irradiance || !saved || !scratch) { set_error("radiance_loss_forward: missing output"); return SVGIR_ERR_INVALID; }
    if ((uintptr_t)saved & 15) { set_error("radiance_loss_forward: saved must be 16-byte aligned"); return SVGIR_ERR_INVALID; }
    if (c->P == 0) { cudaMemsetAsync(loss, 0, sizeof(float), s); return SVGIR_OK; }
    const int grid = rl_grid(c->P);
    static_assert(148 * 8 <= SVGIR_RADIANCE_SCRATCH_FLOATS, "scratch");
    { TimedScope t_("radiance_loss_fwd", s);
      radiance_select_kernel<<<(c->P + RL_WPC - 1) / RL_WPC, RL_THREADS, 0, s>>>(a, sample_index, (float4*)saved);
      radiance_loss_fwd_kernel<false><<<grid, RL_THREADS, 0, s>>>(a, (const float4*)saved, irradiance, scratch, nullptr, nullptr,
                                                                  nullptr, nullptr);
      radiance_loss_reduce_kernel<<<1, 256, 0, s>>>(grid, 1.f / (3.f * (float)c->P), scratch, loss); }
    return check_launch("radiance_loss_forward", false, s);
}

int svgir_radiance_loss_backward(const svgir_radiance_loss_cfg* c, const svgir_radiance_loss_in* in, const float* grad_loss,
                                 const float* irradiance, const float* saved, float* d_albedo, float* d_roughness,
                                 float* d_env, float* d_env_scratch, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    RLArgs a;
    int rc = rl_prepare(c, in, a, s);
    if (rc) return rc;
    if (!irradiance || !saved || ((uintptr_t)saved & 15)) { set_error("radiance_loss_backward: irradiance / saved of the forward are required"); return SVGIR_ERR_INVALID; }
    if (d_env && !d_env_scratch) { set_error("radiance_loss_backward: d_env needs d_env_scratch [SVGIR_RADIANCE_ENV_COPIES*env_h*env_w*4]"); return SVGIR_ERR_INVALID; }
    if (c->P == 0) return SVGIR_OK;
    const int ntex = a.He * a.We;
    if (d_env && cudaMemsetAsync(d_env_scratch, 0, (size_t)ENV_COPIES * ntex * 4 * sizeof(float), s) != cudaSuccess) { set_error("memset failed"); return SVGIR_ERR_CUDA; }
    { TimedScope t_("radiance_loss_bwd", s);
      radiance_loss_bwd_kernel<<<rl_grid(c->P), RL_THREADS, 0, s>>>(a, grad_loss, (const float4*)saved, irradiance, d_albedo, d_roughness,
                                                                   d_env ? d_env_scratch : nullptr); }
    if (d_env) launch_env_grad_finalize(ntex, ENV_COPIES, c->env_mode, d_env_scratch, in->env, d_env, s);
    return check_launch("radiance_loss_backward", false, s);
}

int svgir_radiance_loss_forward_backward(const svgir_radiance_loss_cfg* c, const svgir_radiance_loss_in* in,
                                         const float* grad_loss, float* loss, float* irradiance, int32_t* sample_index,
                                         float* saved, float* scratch, float* d_albedo, float* d_roughness, float* d_env,
                                         float* d_env_scratch, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    RLArgs a;
    int rc = rl_prepare(c, in, a, s);
    if (rc) return rc;
    if (!loss || !irradiance || !saved || !scratch) { set_error("radiance_loss_forward_backward: missing output"); return SVGIR_ERR_INVALID; }
    if ((uintptr_t)saved & 15) { set_error("radiance_loss_forward_backward: saved must be 16-byte aligned"); return SVGIR_ERR_INVALID; }
    if (d_env && !d_env_scratch) { set_error("radiance_loss_forward_backward: d_env needs d_env_scratch [SVGIR_RADIANCE_ENV_COPIES*env_h*env_w*4]"); return SVGIR_ERR_INVALID; }
    if (c->P == 0) { cudaMemsetAsync(loss, 0, sizeof(float), s); return SVGIR_OK; }
    const int ntex = a.He * a.We;
    if (d_env && cudaMemsetAsync(d_env_scratch, 0, (size_t)ENV_COPIES * ntex * 4 * sizeof(float), s) != cudaSuccess) { set_error("memset failed"); return SVGIR_ERR_CUDA; }
    const int grid = rl_grid(c->P);
    { TimedScope t_("radiance_loss_fused", s);
      radiance_select_kernel<<<(c->P + RL_WPC - 1) / RL_WPC, RL_THREADS, 0, s>>>(a, sample_index, (float4*)saved);
      radiance_loss_fwd_kernel<true><<<grid, RL_THREADS, 0, s>>>(a, (const float4*)saved, irradiance, scratch, grad_loss, d_albedo,
                                                                 d_roughness, d_env ? d_env_scratch : nullptr);
      radiance_loss_reduce_kernel<<<1, 256, 0, s>>>(grid, 1.f / (3.f * (float)c->P), scratch, loss); }
    if (d_env) launch_env_grad_finalize(ntex, ENV_COPIES, c->env_mode, d_env_scratch, in->env, d_env, s);
    return check_launch("radiance_loss_forward_backward", false, s);
}

}  // extern "C"
