// Optimiser step and densification bookkeeping for the surfel model.
//
// (1) Fused Adam over all parameter groups in ONE launch. The reference builds torch.optim.Adam(l, lr=..., eps=1e-15)
//     with one group per tensor (scene/gaussian_model.py:737-773) and first patches NaN gradients per tensor
//     (replace_nangrad_to_zero, :775-795); torch runs >= 4 elementwise kernels per group (13 groups in stage 2). Here the
//     gradients are read straight out of the flat (all-reduced) gradient bucket, NaNs are replaced on the fly, and every
//     element costs 16 B read + 12 B written: HBM-bound, 26 M parameters -> ~0.73 GB -> ~0.11 ms at 6.5 TB/s.
//     Arithmetic follows torch's single-tensor Adam: m <- m + (g - m)(1 - b1); v <- b2 v + (1 - b2) g^2;
//     p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).
// (2) add_densification_stats (gaussian_model.py:1270-1276) as one kernel.
// (3) densify_and_prune (:1136-1250) as device-side compaction: a decision kernel (clone / split / prune flags per
//     surfel, evaluated in the reference's order), a scan over the output counts (host side: torch.cumsum), and row
//     gather / split-transform kernels that build every per-surfel tensor of the new model.
#include "common.cuh"

namespace svgir {

struct AdamArgs {
    svgir_adam_group g[SVGIR_ADAM_MAX_GROUPS];
    int first_block[SVGIR_ADAM_MAX_GROUPS + 1];   // first CTA of each group (prefix sums of ceil(numel / 1024))
    int n;
    float beta1, beta2, omb1, omb2, eps, bc1, bc2_sqrt;   // omb = 1 - beta, rounded from the double difference as torch does
};

__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
    int gi = 0;
#pragma unroll 1
    while (gi + 1 < a.n && (int)blockIdx.x >= a.first_block[gi + 1]) gi++;
    const svgir_adam_group& G = a.g[gi];
    const long long base = ((long long)(blockIdx.x - a.first_block[gi]) * 256 + threadIdx.x) * 4;
    if (base >= G.numel) return;
    const float step_size = G.lr / a.bc1;
    const bool vec = base + 4 <= G.numel && (((uintptr_t)G.param | (uintptr_t)G.grad | (uintptr_t)G.exp_avg | (uintptr_t)G.exp_avg_sq) & 15) == 0;
    float p[4], g[4], m[4], v[4];
    const int n = (int)min(4LL, G.numel - base);
    if (vec) {
        *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(G.param + base);
        *reinterpret_cast<float4*>(g) = *reinterpret_cast<const float4*>(G.grad + base);
        *reinterpret_cast<float4*>(m) = *reinterpret_cast<const float4*>(G.exp_avg + base);
        *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(G.exp_avg_sq + base);
    } else {
        for (int i = 0; i < n; i++) { p[i] = G.param[base + i]; g[i] = G.grad[base + i]; m[i] = G.exp_avg[base + i]; v[i] = G.exp_avg_sq[base + i]; }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i >= n) break;
        float gr = g[i];
        if (G.nan_fix && gr != gr) gr = G.nan_value;              // replace_nangrad_to_zero
        m[i] = m[i] + (gr - m[i]) * a.omb1;               // lerp_
        v[i] = v[i] * a.beta2 + a.omb2 * gr * gr;         // mul_ / addcmul_
        const float denom = sqrtf(v[i]) / a.bc2_sqrt + a.eps;
        p[i] = p[i] - step_size * (m[i] / denom);                  // addcdiv_
    }
    if (vec) {
        *reinterpret_cast<float4*>(G.param + base) = *reinterpret_cast<const float4*>(p);
        *reinterpret_cast<float4*>(G.exp_avg + base) = *reinterpret_cast<const float4*>(m);
        *reinterpret_cast<float4*>(G.exp_avg_sq + base) = *reinterpret_cast<const float4*>(v);
    } else {
        for (int i = 0; i < n; i++) { G.param[base + i] = p[i]; G.exp_avg[base + i] = m[i]; G.exp_avg_sq[base + i] = v[i]; }
    }
}

// gaussian_model.py:1270-1276. update_filter = radii > 0 of the view just rendered.
__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const float* __restrict__ viewspace_grad,
                                                            const int32_t* __restrict__ radii, const float* __restrict__ weights,
                                                            float* __restrict__ weights_accum, float* __restrict__ xyz_grad_accum,
                                                            float* __restrict__ denom, float* __restrict__ max_radii2D) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    weights_accum[i] += weights[i];
    const int r = radii[i];
    if (r > 0) {
        const float gx = viewspace_grad[3 * i], gy = viewspace_grad[3 * i + 1];
        xyz_grad_accum[i] += sqrtf(gx * gx + gy * gy);
        denom[i] += 1.f;
        if (max_radii2D) max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);   // train.py: max_radii2D[vis] = max(.., radii[vis])
    }
}

// Flags per surfel, evaluated in the order of densify_and_prune (gaussian_model.py:1224-1250):
//   bit 0 clone   (|grad| >= thr or |grad_n| >= thr_n) and max(scaling) <= percent_dense * extent
//   bit 1 split   (grad  >= thr or grad_n  >= thr_n) and max(scaling) >  percent_dense * extent
//   bit 2 prune   opacity < min_opacity or weights_accum < weights_threshold or max(scaling) > 0.1 * extent
//                 (the screen-size test of :1237-1239 reads max_radii2D, which densification_postfix has just
//                  reset to zero, so it never fires after a densification; it is evaluated on `radii_ws` when given)
// counts[i] = rows this surfel contributes to the new model in section A (itself), B (clone), C (split: 2).
__global__ void __launch_bounds__(256) densify_decide_kernel(const svgir_densify_cfg c, const float* __restrict__ xyz_grad_accum,
                                                             const float* __restrict__ normal_grad_accum, const float* __restrict__ denom,
                                                             const float* __restrict__ scaling_raw, const float* __restrict__ opacity_raw,
                                                             const float* __restrict__ weights_accum, uint8_t* __restrict__ flags,
                                                             int32_t* __restrict__ keep, int32_t* __restrict__ nclone, int32_t* __restrict__ nsplit) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= c.P) return;
    const float d = denom[i];
    float g = xyz_grad_accum[i] / d, gn = normal_grad_accum ? normal_grad_accum[i] / d : 0.f;
    if (g != g) g = 0.f;          // grads[grads.isnan()] = 0 (0 / 0)
    if (gn != gn) gn = 0.f;
    const float s0 = expf(scaling_raw[3 * i]), s1 = expf(scaling_raw[3 * i + 1]), s2 = expf(scaling_raw[3 * i + 2]);
    const float smax = fmaxf(s0, fmaxf(s1, s2));
    const bool sel = fabsf(g) >= c.grad_threshold || fabsf(gn) >= c.grad_normal_threshold;
    const bool sel_split = g >= c.grad_threshold || gn >= c.grad_normal_threshold;
    const bool small = smax <= c.percent_dense * c.extent;
    const bool clone = sel && small, split = sel_split && !small;
    const float op = 1.f / (1.f + expf(-opacity_raw[i]));
    // the final prune mask is evaluated on the model AFTER clone / split; for an original surfel that survives the
    // split its own attributes decide; clones inherit the source's opacity / scaling and get weights_accum = 1
    const bool big_ws = c.use_screen_size && smax > 0.1f * c.extent;
    const bool prune_self = op < c.min_opacity || weights_accum[i] < c.weights_threshold || big_ws;
    const bool prune_clone = op < c.min_opacity || 1.f < c.weights_threshold || big_ws;
    // split children: scaling / (0.8 N), third axis exp(-1e10) = 0; same opacity
    const float cs = fmaxf(s0, s1) / (0.8f * 2.f);
    const bool prune_child = op < c.min_opacity || 1.f < c.weights_threshold || (c.use_screen_size && cs > 0.1f * c.extent);
    flags[i] = (uint8_t)((clone ? 1 : 0) | (split ? 2 : 0) | (prune_self ? 4 : 0));
    keep[i] = (!split && !prune_self) ? 1 : 0;
    nclone[i] = (clone && !prune_clone) ? 1 : 0;
    nsplit[i] = (split && !prune_child) ? 1 : 0;
}

// Source index of every row of the new model: section A = surviving originals (ascending), B = clones, C = first
// copies of the split surfels, D = second copies (repeat(N, 1): all selected once, then all again). src[j] = source
// surfel, kind[j] = 0 copy, 1 clone, 2 / 3 split child (first / second sample set).
__global__ void __launch_bounds__(256) densify_index_kernel(int P, const int32_t* __restrict__ keep, const int32_t* __restrict__ nclone,
                                                            const int32_t* __restrict__ nsplit, const int64_t* __restrict__ keep_scan,
                                                            const int64_t* __restrict__ clone_scan, const int64_t* __restrict__ split_scan,
                                                            long long nA, long long nB, long long nC, int32_t* __restrict__ src,
                                                            uint8_t* __restrict__ kind) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    if (keep[i]) { const long long j = keep_scan[i] - 1; src[j] = i; kind[j] = 0; }
    if (nclone[i]) { const long long j = nA + clone_scan[i] - 1; src[j] = i; kind[j] = 1; }
    if (nsplit[i]) {
        const long long j = nA + nB + split_scan[i] - 1;
        src[j] = i; kind[j] = 2;
        src[j + nC] = i; kind[j + nC] = 3;
    }
}

// dst[j][:] = src[index[j]][:] for K floats per row; rows of NEW surfels (kind != 0) are zero-filled when zero_new
// (Adam moments: cat_tensors_to_optimizer appends zeros, gaussian_model.py:1066-1069).
__global__ void __launch_bounds__(256) gather_rows_kernel(long long n_rows, int K, const float* __restrict__ src_rows,
                                                          const int32_t* __restrict__ index, const uint8_t* __restrict__ kind,
                                                          int zero_new, float new_value, float* __restrict__ dst) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= n_rows * K) return;
    const long long j = t / K;
    const int k = (int)(t - j * K);
    const bool is_new = kind && kind[j] != 0;
    dst[t] = (zero_new && is_new) ? new_value : src_rows[(long long)index[j] * K + k];
}

// densify_and_split (:1152-1160) for the rows of kind 2 / 3: new_xyz = R(q) (std * z) + xyz, new_scaling =
// log(exp(s) / (0.8 N)) with the third axis set to -1e10. `normal_samples` [2 nC, 3] ~ N(0,1) are drawn by the caller
// (torch.normal in the reference), row j - (nA + nB) of it belongs to new row j.
__global__ void __launch_bounds__(256) densify_split_kernel(long long n_new, long long first_split, const int32_t* __restrict__ src,
                                                            const uint8_t* __restrict__ kind, const float* __restrict__ xyz_old,
                                                            const float* __restrict__ scaling_old, const float* __restrict__ rotation_old,
                                                            const float* __restrict__ normal_samples, float* __restrict__ xyz_new,
                                                            float* __restrict__ scaling_new) {
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= n_new || kind[j] < 2) return;
    const int i = src[j];
    const float* z = normal_samples + 3 * (j - first_split);
    const float s0 = expf(scaling_old[3 * i]), s1 = expf(scaling_old[3 * i + 1]), s2 = expf(scaling_old[3 * i + 2]);
    const float sx = s0 * z[0], sy = s1 * z[1], sz = s2 * z[2];
    // build_rotation (utils/general_utils.py:117-149): normalised quaternion (r, x, y, z)
    float r = rotation_old[4 * i], x = rotation_old[4 * i + 1], y = rotation_old[4 * i + 2], w = rotation_old[4 * i + 3];
    const float nq = sqrtf(r * r + x * x + y * y + w * w);
    r /= nq; x /= nq; y /= nq; w /= nq;
    const float R00 = 1 - 2 * (y * y + w * w), R01 = 2 * (x * y - r * w), R02 = 2 * (x * w + r * y);
    const float R10 = 2 * (x * y + r * w), R11 = 1 - 2 * (x * x + w * w), R12 = 2 * (y * w - r * x);
    const float R20 = 2 * (x * w - r * y), R21 = 2 * (y * w + r * x), R22 = 1 - 2 * (x * x + y * y);
    xyz_new[3 * j] = R00 * sx + R01 * sy + R02 * sz + xyz_old[3 * i];
    xyz_new[3 * j + 1] = R10 * sx + R11 * sy + R12 * sz + xyz_old[3 * i + 1];
    xyz_new[3 * j + 2] = R20 * sx + R21 * sy + R22 * sz + xyz_old[3 * i + 2];
    scaling_new[3 * j] = logf(s0 / 1.6f);
    scaling_new[3 * j + 1] = logf(s1 / 1.6f);
    scaling_new[3 * j + 2] = -1e10f;
}

}  // namespace svgir

using namespace svgir;

extern "C" int svgir_adam_step(const svgir_adam_group* groups, int n_groups, double beta1, double beta2, double eps, int step,
                               void* stream) {
    if (!groups || n_groups <= 0 || n_groups > SVGIR_ADAM_MAX_GROUPS || step < 1) { set_error("adam_step: 1..%d groups, step >= 1", SVGIR_ADAM_MAX_GROUPS); return SVGIR_ERR_INVALID; }
    AdamArgs a;
    a.n = n_groups;
    int blocks = 0;
    for (int i = 0; i < n_groups; i++) {
        const svgir_adam_group& g = groups[i];
        if (g.numel < 0 || (g.numel > 0 && (!g.param || !g.grad || !g.exp_avg || !g.exp_avg_sq))) { set_error("adam_step: group %d has null pointers", i); return SVGIR_ERR_INVALID; }
        a.g[i] = g;
        a.first_block[i] = blocks;
        blocks += (int)((g.numel + 1023) / 1024);
    }
    a.first_block[n_groups] = blocks;
    a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps;
    a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    a.bc1 = (float)(1.0 - pow(beta1, (double)step));
    a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    if (blocks == 0) return SVGIR_OK;
    cudaStream_t s = (cudaStream_t)stream;
    { TimedScope ts_("adam", s); adam_kernel<<<blocks, 256, 0, s>>>(a); }
    return check_launch("adam", false, s);
}

extern "C" int svgir_densify_stats(int P, const float* viewspace_grad, const int32_t* radii, const float* weights,
                                   float* weights_accum, float* xyz_grad_accum, float* denom, float* max_radii2D, void* stream) {
    if (P < 0 || (P > 0 && (!viewspace_grad || !radii || !weights || !weights_accum || !xyz_grad_accum || !denom))) { set_error("densify_stats: null pointer"); return SVGIR_ERR_INVALID; }
    if (P == 0) return SVGIR_OK;
    cudaStream_t s = (cudaStream_t)stream;
    { TimedScope ts_("densify_stats", s); densify_stats_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, viewspace_grad, radii, weights, weights_accum, xyz_grad_accum, denom, max_radii2D); }
    return check_launch("densify_stats", false, s);
}

extern "C" int svgir_densify_decide(const svgir_densify_cfg* c, const float* xyz_grad_accum, const float* normal_grad_accum,
                                    const float* denom, const float* scaling_raw, const float* opacity_raw, const float* weights_accum,
                                    uint8_t* flags, int32_t* keep, int32_t* nclone, int32_t* nsplit, void* stream) {
    if (!c || c->P < 0 || (c->P > 0 && (!xyz_grad_accum || !denom || !scaling_raw || !opacity_raw || !weights_accum || !flags || !keep || !nclone || !nsplit))) { set_error("densify_decide: null pointer"); return SVGIR_ERR_INVALID; }
    if (c->P == 0) return SVGIR_OK;
    cudaStream_t s = (cudaStream_t)stream;
    { TimedScope ts_("densify_decide", s); densify_decide_kernel<<<(c->P + 255) / 256, 256, 0, s>>>(*c, xyz_grad_accum, normal_grad_accum, denom, scaling_raw, opacity_raw, weights_accum, flags, keep, nclone, nsplit); }
    return check_launch("densify_decide", false, s);
}

extern "C" int svgir_densify_index(int P, const int32_t* keep, const int32_t* nclone, const int32_t* nsplit, const int64_t* keep_scan,
                                   const int64_t* clone_scan, const int64_t* split_scan, long long nA, long long nB, long long nC,
                                   int32_t* src, uint8_t* kind, void* stream) {
    if (P <= 0) return SVGIR_OK;
    if (!keep || !nclone || !nsplit || !keep_scan || !clone_scan || !split_scan || !src || !kind) { set_error("densify_index: null pointer"); return SVGIR_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    { TimedScope ts_("densify_index", s); densify_index_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, keep, nclone, nsplit, keep_scan, clone_scan, split_scan, nA, nB, nC, src, kind); }
    return check_launch("densify_index", false, s);
}

extern "C" int svgir_gather_rows(long long n_rows, int K, const float* src_rows, const int32_t* index, const uint8_t* kind,
                                 int zero_new, float new_value, float* dst, void* stream) {
    if (n_rows <= 0 || K <= 0) return SVGIR_OK;
    if (!src_rows || !index || !dst) { set_error("gather_rows: null pointer"); return SVGIR_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    const long long n = n_rows * K;
    { TimedScope ts_("gather_rows", s); gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n_rows, K, src_rows, index, kind, zero_new, new_value, dst); }
    return check_launch("gather_rows", false, s);
}

extern "C" int svgir_densify_split(long long n_new, long long first_split, const int32_t* src, const uint8_t* kind, const float* xyz_old,
                                   const float* scaling_old, const float* rotation_old, const float* normal_samples, float* xyz_new,
                                   float* scaling_new, void* stream) {
    if (n_new <= 0) return SVGIR_OK;
    if (!src || !kind || !xyz_old || !scaling_old || !rotation_old || !normal_samples || !xyz_new || !scaling_new) { set_error("densify_split: null pointer"); return SVGIR_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    { TimedScope ts_("densify_split", s); densify_split_kernel<<<(unsigned)((n_new + 255) / 256), 256, 0, s>>>(n_new, first_split, src, kind, xyz_old, scaling_old, rotation_old, normal_samples, xyz_new, scaling_new); }
    return check_launch("densify_split", false, s);
}
