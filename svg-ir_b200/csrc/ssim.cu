// SSIM of two images and its gradient with respect to the first, as two kernels.
//
// Replaces the reference's `ssim` (utils/loss_utils.py:21-62; called twice per stage-2 iteration on the splatted
// colour and on the PBR image, gaussian_renderer/svgss.py:282-293) and what torch autograd makes of it: five
// depth-wise 11x11 conv2d launches forward, ten backward, plus ~40 elementwise kernels over 3x800x800 images.
//   forward : one CTA per 16x16 tile and channel stages the 26x26 halo of both images in shared memory, applies the
//             11-tap Gaussian separably (rows, then columns) to x, y, x^2, y^2, xy, evaluates the SSIM map, and -- so
//             that the backward pass needs no second halo -- also stores the three per-pixel partials
//             g_mu1, g_e11, g_e12 of the map (the algebra is spelled out in the test-side restatement of the reference);
//   backward: d ssim / d x(p) = (1/N) [ (w*g_mu1)(p) + 2 x(p) (w*g_e11)(p) + y(p) (w*g_e12)(p) ], the same separable
//             window applied to the three stored maps.
// The mean is deterministic: per-CTA partial sums, combined in a fixed order by the last CTA to finish.
// STATUS: written and compiled at the end of round 1 after the GPU budget was spent -- NOT yet run on a GPU; the GPU
// test (tests/test_ssim_gpu.py) is skipped unless SVGIR_UNVERIFIED=1. Not used by bench.py or by any default path.
#include <cmath>
#include "common.cuh"

namespace svgir {

#define SSIM_R 5
#define SSIM_TAPS 11
#define SSIM_T 16
#define SSIM_IN (SSIM_T + 2 * SSIM_R)   // 26
#define SSIM_PITCH (SSIM_IN + 1)        // 27: odd pitch, conflict-free column walks

struct SsimWindow { float w[SSIM_TAPS]; };

// loss_utils.py:21-23: exp in double, stored as float32, normalised in float32
static SsimWindow make_window() {
    SsimWindow g;
    float sum = 0.f;
    for (int i = 0; i < SSIM_TAPS; i++) {
        g.w[i] = (float)std::exp(-(double)((i - SSIM_R) * (i - SSIM_R)) / (2.0 * 1.5 * 1.5));
        sum += g.w[i];
    }
    for (int i = 0; i < SSIM_TAPS; i++) g.w[i] /= sum;
    return g;
}

// stages the (zero-padded) 26x26 halo of one channel of `img` around tile (tx, ty)
__device__ __forceinline__ void load_halo(float (*dst)[SSIM_PITCH], const float* __restrict__ img, int H, int W, int tx, int ty) {
    for (int i = threadIdx.x; i < SSIM_IN * SSIM_IN; i += 256) {
        const int r = i / SSIM_IN, c = i - r * SSIM_IN;
        const int gy = ty * SSIM_T + r - SSIM_R, gx = tx * SSIM_T + c - SSIM_R;
        dst[r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(img + (size_t)gy * W + gx) : 0.f;
    }
}

__global__ void __launch_bounds__(256) ssim_fwd_kernel(int C, int H, int W, const float* __restrict__ img1,
                                                       const float* __restrict__ img2, const SsimWindow gw,
                                                       float* __restrict__ gmaps, float* __restrict__ partials,
                                                       unsigned int* __restrict__ counter, float* __restrict__ out) {
    __shared__ float X[SSIM_IN][SSIM_PITCH], Y[SSIM_IN][SSIM_PITCH];
    __shared__ float Hm[5][SSIM_IN][SSIM_T];
    __shared__ float red[8];
    __shared__ bool last;
    const int tx = blockIdx.x, ty = blockIdx.y, c = blockIdx.z;
    const size_t HW = (size_t)H * W;
    load_halo(X, img1 + c * HW, H, W, tx, ty);
    load_halo(Y, img2 + c * HW, H, W, tx, ty);
    __syncthreads();
    // rows: 26 x 16 positions, five moments each
    for (int i = threadIdx.x; i < SSIM_IN * SSIM_T; i += 256) {
        const int r = i / SSIM_T, col = i - r * SSIM_T;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int k = 0; k < SSIM_TAPS; k++) {
            const float w = gw.w[k], x = X[r][col + k], y = Y[r][col + k];
            a0 = fmaf(w, x, a0); a1 = fmaf(w, y, a1);
            a2 = fmaf(w, x * x, a2); a3 = fmaf(w, y * y, a3); a4 = fmaf(w, x * y, a4);
        }
        Hm[0][r][col] = a0; Hm[1][r][col] = a1; Hm[2][r][col] = a2; Hm[3][r][col] = a3; Hm[4][r][col] = a4;
    }
    __syncthreads();
    // columns: one output pixel per thread
    const int ly = threadIdx.x / SSIM_T, lx = threadIdx.x - ly * SSIM_T;
    const int gy = ty * SSIM_T + ly, gx = tx * SSIM_T + lx;
    float m = 0.f;
    if (gy < H && gx < W) {
        float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < SSIM_TAPS; k++) {
            const float w = gw.w[k];
#pragma unroll
            for (int j = 0; j < 5; j++) v[j] = fmaf(w, Hm[j][ly + k][lx], v[j]);
        }
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float mu1 = v[0], mu2 = v[1];
        const float s1 = v[2] - mu1 * mu1, s2 = v[3] - mu2 * mu2, s12 = v[4] - mu1 * mu2;
        const float A = 2.f * mu1 * mu2 + C1, B = 2.f * s12 + C2;
        const float Cc = mu1 * mu1 + mu2 * mu2 + C1, D = s1 + s2 + C2;
        const float inv = 1.0f / (Cc * D);
        m = A * B * inv;
        if (gmaps) {
            const float g_e11 = -m / D;
            const float g_e12 = 2.f * A * inv;
            const float g_mu1 = 2.f * mu2 * B * inv - 2.f * mu1 * m / Cc - 2.f * mu1 * g_e11 - mu2 * g_e12;
            const size_t p = (size_t)gy * W + gx;
            gmaps[((size_t)0 * C + c) * HW + p] = g_mu1;
            gmaps[((size_t)1 * C + c) * HW + p] = g_e11;
            gmaps[((size_t)2 * C + c) * HW + p] = g_e12;
        }
    }
    // deterministic mean: per-CTA partial, fixed-order sum by the last CTA
    m = warp_sum(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) s += red[i];
        partials[bid] = s;
        __threadfence();
        last = atomicAdd(counter, 1u) == nblocks - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    float s = 0.f;
    for (unsigned i = threadIdx.x; i < nblocks; i += 256) s += __ldcg(partials + i);
    s = warp_sum(s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) t += red[i];
        out[0] = t / ((float)C * (float)H * (float)W);
        *counter = 0;   // ready for the next launch (CUDA-graph replay)
    }
}

__global__ void __launch_bounds__(256) ssim_bwd_kernel(int C, int H, int W, const float* __restrict__ img1,
                                                       const float* __restrict__ img2, const float* __restrict__ gmaps,
                                                       const SsimWindow gw, const float* __restrict__ grad_out,
                                                       float* __restrict__ d_img1) {
    __shared__ float G[SSIM_IN][SSIM_PITCH];
    __shared__ float Hm[SSIM_IN][SSIM_T];
    const int tx = blockIdx.x, ty = blockIdx.y, c = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const int ly = threadIdx.x / SSIM_T, lx = threadIdx.x - ly * SSIM_T;
    const int gy = ty * SSIM_T + ly, gx = tx * SSIM_T + lx;
    float val[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; j++) {
        __syncthreads();   // the previous map's buffers are no longer read
        load_halo(G, gmaps + ((size_t)j * C + c) * HW, H, W, tx, ty);
        __syncthreads();
        for (int i = threadIdx.x; i < SSIM_IN * SSIM_T; i += 256) {
            const int r = i / SSIM_T, col = i - r * SSIM_T;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < SSIM_TAPS; k++) a = fmaf(gw.w[k], G[r][col + k], a);
            Hm[r][col] = a;
        }
        __syncthreads();
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < SSIM_TAPS; k++) v = fmaf(gw.w[k], Hm[ly + k][lx], v);
        val[j] = v;
    }
    if (gy < H && gx < W) {
        const size_t p = (size_t)c * HW + (size_t)gy * W + gx;
        const float up = (grad_out ? grad_out[0] : 1.0f) / ((float)C * (float)H * (float)W);
        d_img1[p] = up * (val[0] + 2.f * img1[p] * val[1] + img2[p] * val[2]);
    }
}

}  // namespace svgir

using namespace svgir;

extern "C" int svgir_ssim_blocks(int C, int H, int W) {
    return C * ((H + SSIM_T - 1) / SSIM_T) * ((W + SSIM_T - 1) / SSIM_T);
}

static int ssim_check(int C, int H, int W, const void* a, const void* b) {
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535 || !a || !b) { set_error("ssim: bad shape or null image"); return SVGIR_ERR_INVALID; }
    return SVGIR_OK;
}

extern "C" int svgir_ssim_forward(int C, int H, int W, const float* img1, const float* img2, float* ssim_out,
                                  float* gmaps, float* partials, unsigned int* counter, void* stream) {
    int rc = ssim_check(C, H, W, img1, img2);
    if (rc) return rc;
    if (!ssim_out || !partials || !counter) { set_error("ssim_forward: null output / scratch"); return SVGIR_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 grid((W + SSIM_T - 1) / SSIM_T, (H + SSIM_T - 1) / SSIM_T, C);
    { TimedScope ts_("ssim_fwd", s); ssim_fwd_kernel<<<grid, 256, 0, s>>>(C, H, W, img1, img2, make_window(), gmaps, partials, counter, ssim_out); }
    return check_launch("ssim_fwd", false, s);
}

extern "C" int svgir_ssim_backward(int C, int H, int W, const float* img1, const float* img2, const float* gmaps,
                                   const float* grad_out, float* d_img1, void* stream) {
    int rc = ssim_check(C, H, W, img1, img2);
    if (rc) return rc;
    if (!gmaps || !d_img1) { set_error("ssim_backward: null gradient maps / output"); return SVGIR_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 grid((W + SSIM_T - 1) / SSIM_T, (H + SSIM_T - 1) / SSIM_T, C);
    { TimedScope ts_("ssim_bwd", s); ssim_bwd_kernel<<<grid, 256, 0, s>>>(C, H, W, img1, img2, gmaps, make_window(), grad_out, d_img1); }
    return check_launch("ssim_bwd", false, s);
}
