// Fused tail of a stage-2 training iteration: G-buffer resolve + image loss, forward and backward.
//
// Replaces the ~120 elementwise / reduction torch kernels the reference launches between the
// rasteriser's forward and backward (gaussian_renderer/svgss.py:187-233: un-premultiply by the
// opacity, split, opacity filter, rgb_to_srgb (utils/graphics_utils.py:198-213); svgss.py:280-294 /
// utils/loss_utils.py:33-34: the L1 terms on "render" and "pbr"; the 0.02 * cos-style normal
// consistency term of svgss.py:313) by ONE kernel per direction:
//   loss = mean|C - gt| + lambda_pbr * mean|srgb(pbr*o + (1-o)*bg) - gt|
//          + lambda_normal * mean(1 - <n_shade, n_geo>),
//   pbr = VF[pbr_ch..+3] / max(o,1e-5),  n_shade = VF[normal_ch..+3] / max(o,1e-5).
// Each pixel is read once (64 B) forward; backward re-reads it and writes all six pixel-gradient
// images the rasteriser's backward consumes (no memsets, no autograd-saved intermediates).
// The loss sum is deterministic: per-block partials, last-arriving block adds them in a fixed order.
// HBM-bound: forward 64 B/pixel, backward 64 + 4*(8+S+NV) B/pixel.
#include "common.cuh"

namespace svgir {

#define LOSS_THREADS 256

struct PixelTerms {
    float o, inv, pbr[3], shn[3], x[3], s[3];
    bool pass[3];   // srgb clamp(0,1) passes the gradient
};

// utils/graphics_utils.py:198-213: where(x > 0.0031308, pow(clamp_min(x, 0.0031308), 1/2.4)*1.055 - 0.055, 12.92*x).clamp(0,1)
__device__ __forceinline__ float srgb_unclamped(float x) {
    return x > 0.0031308f ? powf(fmaxf(x, 0.0031308f), 1.0f / 2.4f) * 1.055f - 0.055f : 12.92f * x;
}
__device__ __forceinline__ float srgb_slope(float x) {
    return x > 0.0031308f ? (1.055f / 2.4f) * powf(x, 1.0f / 2.4f - 1.0f) : 12.92f;
}

__device__ __forceinline__ void pixel_terms(const svgir_train_loss_cfg& c, const svgir_train_loss_in& in, size_t HW,
                                            size_t p, const float* bg, PixelTerms& t) {
    t.o = in.opacity[p];
    t.inv = 1.0f / fmaxf(t.o, 1e-5f);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        t.pbr[k] = in.vfeature[(size_t)(c.pbr_ch + k) * HW + p] * t.inv;
        t.shn[k] = in.vfeature[(size_t)(c.normal_ch + k) * HW + p] * t.inv;
        t.x[k] = t.pbr[k] * t.o + (1.0f - t.o) * bg[k];
        const float u = srgb_unclamped(t.x[k]);
        t.pass[k] = u >= 0.0f && u <= 1.0f;
        t.s[k] = fminf(fmaxf(u, 0.0f), 1.0f);
    }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = 0.f;
    if (wid == 0) {
        r = lane < LOSS_THREADS / 32 ? red[lane] : 0.f;
        r = warp_sum(r);
    }
    return r;  // valid in warp 0
}

// ---- depth2normal (utils/image_utils.py:61-125) -------------------------------------------------------------------
// Camera-space position of pixel (x, y): ((x - cx) d / focal_x, (y - cy) d / focal_y, d). With replicate padding the
// neighbour of a border pixel is the pixel itself. p_c = P_c m_c; p_k = (P_k - p_c) m_k for k = up, left, bottom,
// right; n = p_u x p_l + p_r x p_u + p_b x p_r + p_l x p_b; d2n = n / max(|n|, 1e-12) * m_c.
struct D2N {
    int iu, il, ib, ir;            // pixel indices of the four (clamped) neighbours
    float mc, mu, ml, mb, mr;      // masks as 0 / 1
    float pu[3], pl[3], pb[3], pr[3];
    float n[3], len;               // un-normalised normal and max(|n|, eps)
    float out[3];                  // d2n
};

__device__ __forceinline__ void cam_pos(const svgir_train_loss_cfg& c, const svgir_train_loss_in& in, int x, int y, float P[3]) {
    const float d = in.depth[(size_t)y * c.W + x];
    P[0] = ((float)x - c.cx) * d * c.inv_focal_x;
    P[1] = ((float)y - c.cy) * d * c.inv_focal_y;
    P[2] = d;
}
__device__ __forceinline__ float mask_at(const svgir_train_loss_in& in, int idx) {
    return in.mask ? (in.mask[idx] != 0.f ? 1.f : 0.f) : 1.f;
}
__device__ __forceinline__ void cross3(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ __forceinline__ void d2n_eval(const svgir_train_loss_cfg& c, const svgir_train_loss_in& in, int x, int y, D2N& r) {
    const int W = c.W, H = c.H;
    const int yu = max(y - 1, 0), yb = min(y + 1, H - 1), xl = max(x - 1, 0), xr = min(x + 1, W - 1);
    r.iu = yu * W + x; r.il = y * W + xl; r.ib = yb * W + x; r.ir = y * W + xr;
    r.mc = mask_at(in, y * W + x); r.mu = mask_at(in, r.iu); r.ml = mask_at(in, r.il);
    r.mb = mask_at(in, r.ib); r.mr = mask_at(in, r.ir);
    float Pc[3], Pu[3], Pl[3], Pb[3], Pr[3];
    cam_pos(c, in, x, y, Pc); cam_pos(c, in, x, yu, Pu); cam_pos(c, in, xl, y, Pl);
    cam_pos(c, in, x, yb, Pb); cam_pos(c, in, xr, y, Pr);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float pc = Pc[k] * r.mc;
        r.pu[k] = (Pu[k] - pc) * r.mu; r.pl[k] = (Pl[k] - pc) * r.ml;
        r.pb[k] = (Pb[k] - pc) * r.mb; r.pr[k] = (Pr[k] - pc) * r.mr;
    }
    float a[3], b[3], d[3], e[3];
    cross3(r.pu, r.pl, a); cross3(r.pr, r.pu, b); cross3(r.pb, r.pr, d); cross3(r.pl, r.pb, e);
#pragma unroll
    for (int k = 0; k < 3; k++) r.n[k] = ((a[k] + b[k]) + d[k]) + e[k];
    r.len = fmaxf(sqrtf(r.n[0] * r.n[0] + r.n[1] * r.n[1] + r.n[2] * r.n[2]), 1e-12f);
#pragma unroll
    for (int k = 0; k < 3; k++) r.out[k] = r.n[k] / r.len * r.mc;
}

__global__ void __launch_bounds__(LOSS_THREADS) train_loss_fwd_kernel(const svgir_train_loss_cfg c, const svgir_train_loss_in in,
                                                                      float* __restrict__ loss, float* __restrict__ partials,
                                                                      unsigned int* __restrict__ counter) {
    __shared__ float red[LOSS_THREADS / 32];
    __shared__ bool last;
    const size_t HW = (size_t)c.W * c.H;
    const size_t p = (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
    const float bg[3] = {c.bg[0], c.bg[1], c.bg[2]};
    float l1 = 0.f, l1p = 0.f, nn = 0.f, cnt = 0.f;
    if (p < HW) {
        PixelTerms t;
        pixel_terms(c, in, HW, p, bg, t);
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float g = in.gt[k * HW + p];
            l1 += fabsf(in.color[k * HW + p] - g);
            l1p += fabsf(t.s[k] - g);
            dot += t.shn[k] * in.geo_normal[k * HW + p];
        }
        if (c.normal_mode == 0) {
            nn = 1.0f - dot;
            cnt = 1.f;
        } else {   // cos_loss(n_shade, depth2normal(depth)): only pixels with cos < cos(0) = 1 enter the mean
            D2N r;
            d2n_eval(c, in, (int)(p % c.W), (int)(p / c.W), r);
            const float cs = (t.shn[0] * r.out[0] + t.shn[1] * r.out[1]) + t.shn[2] * r.out[2];
            if (cs < 1.0f) { nn = 1.0f - cs; cnt = 1.f; }
        }
    }
    const float a = block_sum(l1, red), b = block_sum(l1p, red), d = block_sum(nn, red), e = block_sum(cnt, red);
    if (threadIdx.x == 0) {
        partials[4 * blockIdx.x + 0] = a;
        partials[4 * blockIdx.x + 1] = b;
        partials[4 * blockIdx.x + 2] = d;
        partials[4 * blockIdx.x + 3] = e;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // last block: fixed-order sum of the per-block partials (deterministic across runs)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += LOSS_THREADS) {
        s0 += __ldcg(partials + 4 * i);
        s1 += __ldcg(partials + 4 * i + 1);
        s2 += __ldcg(partials + 4 * i + 2);
        s3 += __ldcg(partials + 4 * i + 3);
    }
    s0 = block_sum(s0, red); s1 = block_sum(s1, red); s2 = block_sum(s2, red); s3 = block_sum(s3, red);
    if (threadIdx.x == 0) {
        const float n3 = 3.0f * (float)HW;
        // an empty selection gives torch's mean of an empty tensor: NaN (loss_utils.py:119)
        const float t0 = s0 / n3, t1 = s1 / n3, t2 = s2 / s3;
        loss[1] = t0; loss[2] = t1; loss[3] = t2; loss[4] = s3;
        loss[0] = t0 + c.lambda_pbr * t1 + c.lambda_normal * t2;
        *counter = 0;  // ready for the next launch (CUDA-graph replay)
    }
}

__global__ void __launch_bounds__(LOSS_THREADS) train_loss_bwd_kernel(const svgir_train_loss_cfg c, const svgir_train_loss_in in,
                                                                      const float* __restrict__ grad_loss,
                                                                      const float* __restrict__ loss,
                                                                      const svgir_train_loss_grads g) {
    const size_t HW = (size_t)c.W * c.H;
    const size_t p = (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
    if (p >= HW) return;
    const float up = grad_loss ? grad_loss[0] : 1.0f;
    const float bg[3] = {c.bg[0], c.bg[1], c.bg[2]};
    PixelTerms t;
    pixel_terms(c, in, HW, p, bg, t);
    const float k1 = up / (3.0f * (float)HW), kp = up * c.lambda_pbr / (3.0f * (float)HW);
    float kn = -up * c.lambda_normal / (float)HW;
    float go = 0.f, ginv = 0.f;
    float gvf_pbr[3], gvf_shn[3];
    // the other factor of the normal term: the rasteriser's geometric normal (mode 0) or depth2normal (mode 1)
    float nref[3] = {0.f, 0.f, 0.f};
    bool active = true;
    D2N r;
    if (c.normal_mode != 0) {
        d2n_eval(c, in, (int)(p % c.W), (int)(p / c.W), r);
        const float cs = (t.shn[0] * r.out[0] + t.shn[1] * r.out[1]) + t.shn[2] * r.out[2];
        active = cs < 1.0f;
        kn = active ? -up * c.lambda_normal / loss[4] : 0.f;
        nref[0] = r.out[0]; nref[1] = r.out[1]; nref[2] = r.out[2];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float gt = in.gt[k * HW + p];
        const float dc = in.color[k * HW + p] - gt;
        g.color[k * HW + p] = dc > 0.f ? k1 : (dc < 0.f ? -k1 : 0.f);
        const float ds = t.s[k] - gt;
        const float gs = ds > 0.f ? kp : (ds < 0.f ? -kp : 0.f);
        const float gx = t.pass[k] ? gs * srgb_slope(t.x[k]) : 0.f;
        const float gpbr = gx * t.o;
        go += gx * (t.pbr[k] - bg[k]);
        const float gn = c.normal_mode == 0 ? in.geo_normal[k * HW + p] : nref[k];
        const float gshn = kn * gn;
        g.geo_normal[k * HW + p] = c.normal_mode == 0 ? kn * t.shn[k] : 0.f;
        gvf_pbr[k] = gpbr * t.inv;
        gvf_shn[k] = gshn * t.inv;
        // d(1/max(o,eps)): VF_k = raw_k * inv
        ginv += gpbr * in.vfeature[(size_t)(c.pbr_ch + k) * HW + p] + gshn * in.vfeature[(size_t)(c.normal_ch + k) * HW + p];
    }
    if (t.o >= 1e-5f) go -= ginv * t.inv * t.inv;
    g.opacity[p] = go;
    if (c.normal_mode == 0) {
        if (g.depth) g.depth[p] = 0.f;
    } else if (active && r.mc != 0.f) {
        // d(1 - cos)/d(d2n) = -n_shade; through the normalisation, the four cross products and the back-projection
        // into the depth of this pixel and of its four neighbours (g.depth was zero-filled by the launcher)
        float dn[3];
        {
            const float dh[3] = {kn * t.shn[0] * r.mc, kn * t.shn[1] * r.mc, kn * t.shn[2] * r.mc};
            const float il = 1.0f / r.len;
            const float nh[3] = {r.n[0] * il, r.n[1] * il, r.n[2] * il};
            const float dd = nh[0] * dh[0] + nh[1] * dh[1] + nh[2] * dh[2];
            const bool clamped = r.len <= 1e-12f;   // n / eps: linear
#pragma unroll
            for (int k = 0; k < 3; k++) dn[k] = clamped ? dh[k] * il : (dh[k] - nh[k] * dd) * il;
        }
        float dpu[3], dpl[3], dpb[3], dpr[3], t0[3], t1[3];
        cross3(r.pl, dn, t0); cross3(dn, r.pr, t1);      // n = pu x pl + pr x pu + ...
#pragma unroll
        for (int k = 0; k < 3; k++) dpu[k] = t0[k] + t1[k];
        cross3(dn, r.pu, t0); cross3(r.pb, dn, t1);      // ... pu x pl + pl x pb
#pragma unroll
        for (int k = 0; k < 3; k++) dpl[k] = t0[k] + t1[k];
        cross3(r.pr, dn, t0); cross3(dn, r.pl, t1);      // pb x pr + pl x pb
#pragma unroll
        for (int k = 0; k < 3; k++) dpb[k] = t0[k] + t1[k];
        cross3(r.pu, dn, t0); cross3(dn, r.pb, t1);      // pr x pu + pb x pr
#pragma unroll
        for (int k = 0; k < 3; k++) dpr[k] = t0[k] + t1[k];
        float dPc[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dpu[k] *= r.mu; dpl[k] *= r.ml; dpb[k] *= r.mb; dpr[k] *= r.mr;
            dPc[k] = -(((dpu[k] + dpl[k]) + dpb[k]) + dpr[k]) * r.mc;
        }
        const int x = (int)(p % c.W), y = (int)(p / c.W);
        auto push = [&](int idx, const float dP[3]) {
            const int qx = idx % c.W, qy = idx / c.W;
            const float v = ((float)qx - c.cx) * c.inv_focal_x * dP[0] + ((float)qy - c.cy) * c.inv_focal_y * dP[1] + dP[2];
            if (v != 0.f) atomicAdd(g.depth + idx, v);
        };
        push(y * c.W + x, dPc); push(r.iu, dpu); push(r.il, dpl); push(r.ib, dpb); push(r.ir, dpr);
    }
    if (g.feature)
        for (int k = 0; k < c.S; k++) g.feature[(size_t)k * HW + p] = 0.f;
    for (int k = 0; k < c.NV; k++) {
        float v = 0.f;
        if (k >= c.pbr_ch && k < c.pbr_ch + 3) v = gvf_pbr[k - c.pbr_ch];
        else if (k >= c.normal_ch && k < c.normal_ch + 3) v = gvf_shn[k - c.normal_ch];
        g.vfeature[(size_t)k * HW + p] = v;
    }
}

static int check_cfg(const svgir_train_loss_cfg* c, const svgir_train_loss_in* in) {
    if (!c || !in) { set_error("train_loss: null argument"); return SVGIR_ERR_INVALID; }
    if (c->W <= 0 || c->H <= 0 || c->S < 0 || c->NV < 0) { set_error("train_loss: bad shape"); return SVGIR_ERR_INVALID; }
    if (c->pbr_ch < 0 || c->pbr_ch + 3 > c->NV || c->normal_ch < 0 || c->normal_ch + 3 > c->NV) {
        set_error("train_loss: pbr_ch=%d / normal_ch=%d do not fit NV=%d", c->pbr_ch, c->normal_ch, c->NV);
        return SVGIR_ERR_INVALID;
    }
    if (!c->bg || !in->color || !in->geo_normal || !in->opacity || !in->vfeature || !in->gt) {
        set_error("train_loss: null input pointer");
        return SVGIR_ERR_INVALID;
    }
    if (c->normal_mode != 0 && !in->depth) { set_error("train_loss: normal_mode 1 (depth2normal) needs the depth image"); return SVGIR_ERR_INVALID; }
    return SVGIR_OK;
}

}  // namespace svgir

extern "C" int svgir_train_loss_blocks(int W, int H) {
    return (int)(((long long)W * H + LOSS_THREADS - 1) / LOSS_THREADS);
}

extern "C" int svgir_train_loss_forward(const svgir_train_loss_cfg* cfg, const svgir_train_loss_in* in, float* loss,
                                        float* partials, unsigned int* counter, void* stream) {
    using namespace svgir;
    int rc = check_cfg(cfg, in);
    if (rc) return rc;
    if (!loss || !partials || !counter) { set_error("train_loss_forward: null output pointer"); return SVGIR_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = svgir_train_loss_blocks(cfg->W, cfg->H);
    { TimedScope ts_("train_loss_fwd", s); train_loss_fwd_kernel<<<nb, LOSS_THREADS, 0, s>>>(*cfg, *in, loss, partials, counter); }
    return check_launch("train_loss_fwd", false, s);
}

extern "C" int svgir_train_loss_backward(const svgir_train_loss_cfg* cfg, const svgir_train_loss_in* in,
                                         const float* grad_loss, const float* loss, const svgir_train_loss_grads* g,
                                         void* stream) {
    using namespace svgir;
    int rc = check_cfg(cfg, in);
    if (rc) return rc;
    if (!g || !g->color || !g->geo_normal || !g->opacity || !g->vfeature) {
        set_error("train_loss_backward: null gradient pointer");
        return SVGIR_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = svgir_train_loss_blocks(cfg->W, cfg->H);
    if (cfg->normal_mode != 0) {
        if (!loss || !g->depth) { set_error("train_loss_backward: normal_mode 1 needs the forward's loss buffer and a depth gradient image"); return SVGIR_ERR_INVALID; }
        if (cudaMemsetAsync(g->depth, 0, sizeof(float) * (size_t)cfg->W * cfg->H, s) != cudaSuccess) { set_error("memset failed"); return SVGIR_ERR_CUDA; }
    }
    { TimedScope ts_("train_loss_bwd", s); train_loss_bwd_kernel<<<nb, LOSS_THREADS, 0, s>>>(*cfg, *in, grad_loss, loss, *g); }
    return check_launch("train_loss_bwd", false, s);
}

// ---- evaluation / relighting frame resolve (svgss.py:187-262, eval branch) ---------------------------------
namespace svgir {

__device__ __forceinline__ float srgb_clamped(float x) { return fminf(fmaxf(srgb_unclamped(x), 0.f), 1.f); }

// HBM-bound: 24 floats in, up to 27 floats out per pixel, every image touched exactly once.
__global__ void __launch_bounds__(256) resolve_eval_kernel(int HWi, const float* __restrict__ bgp, const float* __restrict__ opacity,
                                                           const float* __restrict__ feature, const float* __restrict__ vfeature,
                                                           const svgir_resolve_eval_out o) {
    const size_t HW = (size_t)HWi;
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= HW) return;
    const float bg[3] = {bgp[0], bgp[1], bgp[2]};
    const float op = opacity[p];
    const float inv = 1.0f / fmaxf(op, 1e-5f);
    const float om = 1.0f - op;
    float f[7], v[16];
#pragma unroll
    for (int i = 0; i < 7; i++) f[i] = feature[i * HW + p] * inv;
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = vfeature[i * HW + p] * inv;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float fill = om * bg[k];
        if (o.pbr) o.pbr[k * HW + p] = srgb_clamped(v[k] * op + fill);
        if (o.normal) o.normal[k * HW + p] = v[6 + k];
        if (o.base_color) o.base_color[k * HW + p] = srgb_clamped(v[3 + k]) * op + fill;
        if (o.roughness) o.roughness[k * HW + p] = v[9] * op + fill;
        if (o.lights) o.lights[k * HW + p] = srgb_clamped(f[k]) * op + fill;
        if (o.local_lights) o.local_lights[k * HW + p] = srgb_clamped(f[3 + k]) * op + fill;
        if (o.visibility) o.visibility[k * HW + p] = f[6] * op + fill;
        if (o.direct) o.direct[k * HW + p] = srgb_clamped(v[10 + k]);
        if (o.indirect) o.indirect[k * HW + p] = srgb_clamped(v[13 + k]);
    }
}

}  // namespace svgir

extern "C" int svgir_resolve_eval(int W, int H, const float* bg, const float* opacity, const float* feature,
                                  const float* vfeature, const svgir_resolve_eval_out* out, void* stream) {
    using namespace svgir;
    if (W <= 0 || H <= 0 || !bg || !opacity || !feature || !vfeature || !out) {
        set_error("resolve_eval: bad size or null pointer");
        return SVGIR_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long HW = (long long)W * H;
    { TimedScope ts_("resolve_eval", s);
      resolve_eval_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, s>>>((int)HW, bg, opacity, feature, vfeature, *out); }
    return check_launch("resolve_eval", false, s);
}
