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

__global__ void __launch_bounds__(LOSS_THREADS) train_loss_fwd_kernel(const svgir_train_loss_cfg c, const svgir_train_loss_in in,
                                                                      float* __restrict__ loss, float* __restrict__ partials,
                                                                      unsigned int* __restrict__ counter) {
    __shared__ float red[LOSS_THREADS / 32];
    __shared__ bool last;
    const size_t HW = (size_t)c.W * c.H;
    const size_t p = (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
    const float bg[3] = {c.bg[0], c.bg[1], c.bg[2]};
    float l1 = 0.f, l1p = 0.f, nn = 0.f;
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
        nn = 1.0f - dot;
    }
    const float a = block_sum(l1, red), b = block_sum(l1p, red), d = block_sum(nn, red);
    if (threadIdx.x == 0) {
        partials[3 * blockIdx.x + 0] = a;
        partials[3 * blockIdx.x + 1] = b;
        partials[3 * blockIdx.x + 2] = d;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // last block: fixed-order sum of the per-block partials (deterministic across runs)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += LOSS_THREADS) {
        s0 += __ldcg(partials + 3 * i);
        s1 += __ldcg(partials + 3 * i + 1);
        s2 += __ldcg(partials + 3 * i + 2);
    }
    s0 = block_sum(s0, red); s1 = block_sum(s1, red); s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        const float n3 = 3.0f * (float)HW;
        const float t0 = s0 / n3, t1 = s1 / n3, t2 = s2 / (float)HW;
        loss[1] = t0; loss[2] = t1; loss[3] = t2;
        loss[0] = t0 + c.lambda_pbr * t1 + c.lambda_normal * t2;
        *counter = 0;  // ready for the next launch (CUDA-graph replay)
    }
}

__global__ void __launch_bounds__(LOSS_THREADS) train_loss_bwd_kernel(const svgir_train_loss_cfg c, const svgir_train_loss_in in,
                                                                      const float* __restrict__ grad_loss,
                                                                      const svgir_train_loss_grads g) {
    const size_t HW = (size_t)c.W * c.H;
    const size_t p = (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
    if (p >= HW) return;
    const float up = grad_loss ? grad_loss[0] : 1.0f;
    const float bg[3] = {c.bg[0], c.bg[1], c.bg[2]};
    PixelTerms t;
    pixel_terms(c, in, HW, p, bg, t);
    const float k1 = up / (3.0f * (float)HW), kp = up * c.lambda_pbr / (3.0f * (float)HW), kn = -up * c.lambda_normal / (float)HW;
    float go = 0.f, ginv = 0.f;
    float gvf_pbr[3], gvf_shn[3];
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
        const float gn = in.geo_normal[k * HW + p];
        const float gshn = kn * gn;
        g.geo_normal[k * HW + p] = kn * t.shn[k];
        gvf_pbr[k] = gpbr * t.inv;
        gvf_shn[k] = gshn * t.inv;
        // d(1/max(o,eps)): VF_k = raw_k * inv
        ginv += gpbr * in.vfeature[(size_t)(c.pbr_ch + k) * HW + p] + gshn * in.vfeature[(size_t)(c.normal_ch + k) * HW + p];
    }
    if (t.o >= 1e-5f) go -= ginv * t.inv * t.inv;
    g.opacity[p] = go;
    if (g.depth) g.depth[p] = 0.f;
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
                                         const float* grad_loss, const svgir_train_loss_grads* g, void* stream) {
    using namespace svgir;
    int rc = check_cfg(cfg, in);
    if (rc) return rc;
    if (!g || !g->color || !g->geo_normal || !g->opacity || !g->vfeature) {
        set_error("train_loss_backward: null gradient pointer");
        return SVGIR_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = svgir_train_loss_blocks(cfg->W, cfg->H);
    { TimedScope ts_("train_loss_bwd", s); train_loss_bwd_kernel<<<nb, LOSS_THREADS, 0, s>>>(*cfg, *in, grad_loss, *g); }
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
