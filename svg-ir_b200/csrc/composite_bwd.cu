// Backward compositing: per-pixel back-to-front re-traversal producing per-surfel gradients.
//
// Replaces renderCUDA<3> of backward.cu (svgss_rasterization/cuda_rasterizer/backward.cu:530-934;
// stage 1: rgss-rasterization/cuda_rasterizer/backward.cu:432-757), which issues 13+S+VS global
// fp32 atomics per (pixel, surfel) hit (69 at the stage-2 training shape) and keeps 1.3 KB of
// per-thread recurrences in local memory.  Here:
//   * the per-channel recurrences  acc_c <- a_prev*v_prev,c + (1-a_prev)*acc_c  are collapsed into
//     ONE scalar recurrence on A = sum_c acc_c*g_c (they are linear with channel-independent
//     coefficients), so a pixel carries 4 floats of state instead of 2*(7+S+VS/4);
//   * every gradient component is sum_hits s_sel(hit) * G[pixel(hit)][channel], with s a handful of
//     per-hit scalars and G the tile's pixel-gradient matrix kept in shared memory.  Hit lanes
//     publish their scalars to a per-warp scratch; then each LANE OWNS A COMPONENT and loops over
//     the warp's hits -- no shuffle tree, no atomics;
//   * per-warp partial sums go to warp-private shared accumulators and are combined into one
//     global atomic per (tile, instance, component);
//   * the traversal starts at the tile's max n_contrib instead of the end of the tile's range.
// Gradient semantics (x10 on the normal gradient, un-weighted depth-differencing term on mean2D,
// no gradient through the bilinear weights, ...) follow the reference; see Appendix C of SURVEY.md.
#include "common.cuh"

namespace svgir {

#define BWD_BATCH 16
#define HIT_STRIDE 13   // 12 words per hit, odd stride -> conflict-free publication
#define NWARP 8

template <int S_T, int NV_T, bool RGSS>
__global__ void __launch_bounds__(TILE_PIX) composite_bwd_kernel(
    const svgir_raster_cfg c, const float* __restrict__ features, const float* __restrict__ vfeatures,
    const float4* __restrict__ rec, const uint2* __restrict__ ranges,
    const uint32_t* __restrict__ point_list, const float* __restrict__ final_T,
    const float* __restrict__ final_D, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ gpix_color, const float* __restrict__ gpix_normal,
    const float* __restrict__ gpix_depth, const float* __restrict__ gpix_opac,
    const float* __restrict__ gpix_feature, const float* __restrict__ gpix_vfeature,
    float* __restrict__ geo_grad, float* __restrict__ dL_dfeatures, float* __restrict__ dL_dvfeatures,
    const int32_t* __restrict__ num_rendered, const uint32_t* __restrict__ tile_order) {
    constexpr bool GENERIC = S_T < 0;
    const int S = GENERIC ? c.S : S_T;
    const int NV = GENERIC ? c.VS / 4 : NV_T;
    const int SP = (S + 3) & ~3;
    const int STRIDE = SVGIR_REC_FLOATS + SP + 4 * NV;
    const int CH = STRIDE / 4;
    const int NCOMP = SVGIR_GEO_GRAD_FLOATS + SP + 4 * NV;  // accumulator row
    const int NG = 8 + S + NV;                              // pixel-gradient row: 1,gC3,gN3,gD',gF,gVF
    const int GS = NG | 1;                                  // odd stride

    extern __shared__ __align__(16) float smem[];
    float* stage = smem;                                   // [BWD_BATCH][STRIDE]
    float* G = stage + BWD_BATCH * STRIDE;                 // [256][GS]
    float* hits = G + TILE_PIX * GS;                       // [NWARP][32][HIT_STRIDE]
    float* acc = hits + NWARP * 32 * HIT_STRIDE;           // [NWARP][BWD_BATCH][NCOMP]
    int* ids = reinterpret_cast<int*>(acc + NWARP * BWD_BATCH * NCOMP);  // [BWD_BATCH]
    __shared__ int tile_max_s;

    const int W = c.W, H = c.H;
    const int gx = (W + TILE - 1) / TILE;
    const int tile = (int)tile_order[blockIdx.x];   // heaviest tiles first (binning.cu: tile_scan_kernel)
    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);
    if (total == 0 || num_rendered[1]) return;  // empty tile, or the forward's bins overflowed (nothing valid)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int px = (tile % gx) * TILE + (tid & 15), py = (tile / gx) * TILE + (tid >> 4);
    const bool inside = px < W && py < H;
    const size_t HW = (size_t)H * W;
    const size_t pix_id = (size_t)W * py + px;
    const float pxf = (float)px, pyf = (float)py;

    bool surface = true, ppd = true, normalize_depth = true;
    if (!RGSS) {
        surface = c.n_config > 0 && c.config[0] > 0;
        normalize_depth = c.n_config > 1 && c.config[1] > 0;
        ppd = c.n_config > 2 && c.config[2] > 0;
    }
    const bool sv = surface && ppd;
    const bool feat_to_alpha = !RGSS || c.backward_geometry != 0;

    const float T_final = inside ? final_T[pix_id] : 0.f;
    const float D_final = (inside && normalize_depth) ? final_D[pix_id] : 0.f;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

    // ---- pixel-gradient row -----------------------------------------------------------------
    float* Grow = G + tid * GS;
    float gD = 0.f, gDn = 0.f, Kpix = 0.f;
    {
        float gC[3] = {0, 0, 0}, gN[3] = {0, 0, 0}, gO = 0.f;
        if (inside) {
#pragma unroll
            for (int i = 0; i < 3; i++) gC[i] = gpix_color[i * HW + pix_id];
#pragma unroll
            for (int i = 0; i < 3; i++) gN[i] = gpix_normal[i * HW + pix_id];
            gD = gpix_depth[pix_id];
            gO = gpix_opac[pix_id];
        }
        const float omT = 1.f - T_final;
        gDn = normalize_depth ? gD / omT : gD;
        Grow[0] = 1.0f;
        Grow[1] = gC[0]; Grow[2] = gC[1]; Grow[3] = gC[2];
        Grow[4] = surface ? gN[0] : 0.f;   // the x10 of backward.cu:804 is applied at the flush
        Grow[5] = surface ? gN[1] : 0.f;
        Grow[6] = surface ? gN[2] : 0.f;
        Grow[7] = gDn;
        for (int i = 0; i < S; i++) Grow[8 + i] = inside ? gpix_feature[i * HW + pix_id] : 0.f;
        for (int i = 0; i < NV; i++) Grow[8 + S + i] = inside ? gpix_vfeature[i * HW + pix_id] : 0.f;
        // pixel-constant part of dL/dalpha that is divided by (1-alpha):
        //   opacity output (+gO*T_f), background (-T_f*bg.gC), depth normalisation or the +10T term
        const float* bg = c.bg;
        const float bgdot = bg[0] * gC[0] + bg[1] * gC[1] + bg[2] * gC[2];
        Kpix = gO * T_final - T_final * bgdot;
        if (normalize_depth) Kpix += gD * D_final / omT / omT * -T_final;
        else Kpix += -T_final * (10.f * gD);
    }

    // ---- component ownership for the reduction ------------------------------------------------
    // component k of the accumulator row: geo[16] | F[SP] | VF[4*NV]
    //   geo: 0,1 mean2D.xy  2,3,4 conic.x,y,w  5 opacity  6..8 colour  9..11 normal  12 depth
    // value(k) = sum_hits hit[sel(k)] * G[pix][gch(k)]
    //   hit words: 0 w | 1..4 w*w0..w3 | 5,6 dmean | 7,8,9 dconic | 10 dopacity | 11 pixel
    constexpr int MAXPASS = GENERIC ? (SVGIR_GEO_GRAD_FLOATS + SVGIR_MAX_S + 4 * SVGIR_MAX_NV + 31) / 32
                                    : (SVGIR_GEO_GRAD_FLOATS + ((S_T + 3) & ~3) + 4 * (NV_T > 0 ? NV_T : 0) + 31) / 32;
    int csel[MAXPASS], cgch[MAXPASS];
#pragma unroll
    for (int p = 0; p < MAXPASS; p++) {
        const int k = lane + 32 * p;
        int sel = -1, gch = 0;
        if (k < 6) { sel = 5 + k; gch = 0; }
        else if (k < 9) { sel = 0; gch = 1 + (k - 6); }
        else if (k < 12) { sel = 0; gch = 4 + (k - 9); }
        else if (k == 12) { sel = 0; gch = 7; }
        else if (k >= 16 && k < 16 + S) { sel = 0; gch = 8 + (k - 16); }
        else if (k >= 16 + SP && k < NCOMP) { const int v = k - 16 - SP; sel = 1 + (v & 3); gch = 8 + S + (v >> 2); }
        csel[p] = sel; cgch[p] = gch;
    }

    // ---- tile-wide traversal start --------------------------------------------------------------
    if (tid == 0) tile_max_s = 0;
    __syncthreads();
    {
        int m = last_contributor;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) atomicMax(&tile_max_s, m);
    }
    __syncthreads();
    const int tile_max = min(tile_max_s, total);
    if (tile_max == 0) return;

    float* my_hits = hits + wid * 32 * HIT_STRIDE;
    float* my_acc = acc + wid * BWD_BATCH * NCOMP;

    float T = T_final, A = 0.f, last_alpha = 0.f, V_last = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    for (int top = tile_max; top > 0; top -= BWD_BATCH) {
        const int nb = min(BWD_BATCH, top);
        __syncthreads();  // previous batch fully flushed
        // stage instances top-1, top-2, ... (back to front) and clear the accumulators
        for (int q = tid; q < nb * CH; q += TILE_PIX) {
            const int i = q / CH, ch = q - i * CH;
            const int id = (int)point_list[range.x + top - 1 - i];
            float4 v;
            if (ch < REC_F4) {
                v = __ldg(rec + (size_t)id * REC_F4 + ch);
                if (ch == 0) ids[i] = id;
            } else if (ch < REC_F4 + SP / 4) {
                const int f0 = (ch - REC_F4) * 4;
                const float* src = features + (size_t)id * S + f0;
                if ((S & 3) == 0) v = __ldg(reinterpret_cast<const float4*>(src));
                else {
                    v.x = f0 + 0 < S ? __ldg(src + 0) : 0.f;
                    v.y = f0 + 1 < S ? __ldg(src + 1) : 0.f;
                    v.z = f0 + 2 < S ? __ldg(src + 2) : 0.f;
                    v.w = f0 + 3 < S ? __ldg(src + 3) : 0.f;
                }
            } else {
                v = __ldg(reinterpret_cast<const float4*>(vfeatures + (size_t)id * (4 * NV)) + (ch - REC_F4 - SP / 4));
            }
            reinterpret_cast<float4*>(stage)[q] = v;
        }
        for (int q = tid; q < NWARP * BWD_BATCH * NCOMP; q += TILE_PIX) acc[q] = 0.f;
        __syncthreads();

        for (int j = 0; j < nb; j++) {
            const int pos = top - 1 - j;  // 0-based position in the tile's sorted list
            const float4* r = reinterpret_cast<const float4*>(stage + j * STRIDE);
            const float4 q0 = r[0];
            const float4 q1 = r[1];
            PairEval e;
            bool hit = false;
            if (pos < last_contributor) hit = eval_alpha<RGSS>(pxf, pyf, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, e);
            const unsigned ballot = __ballot_sync(0xffffffffu, hit);
            if (ballot == 0) continue;
            if (hit) {
                const float inv1ma = __frcp_rn(1.f - e.alpha);
                T = T * inv1ma;
                const float w = e.alpha * T;
                float depth_k = q1.z;
                float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
                float gzx = 0.f, gzy = 0.f;
                if (sv) {
                    const float4 q2 = r[2];
                    const float4 q3 = r[3];
                    const float u0 = fma_(e.dx, q2.x, mul_(e.dy, q2.y));
                    const float u1 = fma_(e.dx, q2.z, mul_(e.dy, q2.w));
                    depth_k = sub_(q1.z, fma_(q3.x, u0, mul_(q3.y, u1)));
                    gzx = q3.x * q2.x + q3.y * q2.z;   // J6*J0 + J9*J2 (backward.cu:915)
                    gzy = q3.x * q2.y + q3.y * q2.w;   // J6*J1 + J9*J3
                    if (!RGSS) {
                        float u = fmaf(u0, q1.w, 0.5f), v = fmaf(u1, q3.z, 0.5f);
                        u = fminf(0.999f, fmaxf(0.001f, u));
                        v = fminf(0.999f, fmaxf(0.001f, v));
                        w0 = (1.0f - u) * (1.0f - v);
                        w1 = u * (1.0f - v);
                        w2 = (1.0f - u) * v;
                        w3 = u * v;
                    }
                }
                // V = sum_c value_c * g_c over all blended channels (normal uses the plain gN)
                const float4 q4 = r[4];
                float V = q4.x * Grow[1] + q4.y * Grow[2] + q4.z * Grow[3];
                if (surface) {
                    const float4 q5 = r[5];
                    V += q4.w * Grow[4] + q5.x * Grow[5] + q5.y * Grow[6];
                }
                V = fmaf(depth_k, gDn, V);
                const float* f = stage + j * STRIDE + SVGIR_REC_FLOATS;
                if (feat_to_alpha)
                    for (int ch = 0; ch < S; ch++) V = fmaf(f[ch], Grow[8 + ch], V);
                const float4* vf = reinterpret_cast<const float4*>(f + SP);
                for (int cidx = 0; cidx < NV; cidx++) {
                    const float4 t = vf[cidx];
                    const float s4 = ((t.x * w0 + t.y * w1) + t.z * w2) + t.w * w3;
                    V = fmaf(s4, Grow[8 + S + cidx], V);
                }
                A = last_alpha * V_last + (1.f - last_alpha) * A;
                V_last = V;
                last_alpha = e.alpha;
                const float dL_dalpha = (V - A) * T + Kpix * inv1ma;
                const float dL_ddist = dL_dalpha * q1.y * -0.5f * e.G;
                float dmx = dL_ddist * 2.f * (q0.z * e.dx + q0.w * e.dy) * ddelx_dx;
                float dmy = dL_ddist * 2.f * (q1.x * e.dy + q0.w * e.dx) * ddely_dy;
                if (sv) { dmx -= gD * gzx; dmy -= gD * gzy; }
                const int rank = __popc(ballot & ((1u << lane) - 1u));
                float* h = my_hits + rank * HIT_STRIDE;
                h[0] = w; h[1] = w * w0; h[2] = w * w1; h[3] = w * w2; h[4] = w * w3;
                h[5] = dmx; h[6] = dmy;
                h[7] = dL_ddist * (e.dx * e.dx);
                h[8] = dL_ddist * (e.dx * e.dy);
                h[9] = dL_ddist * (e.dy * e.dy);
                h[10] = e.G * dL_dalpha;
                h[11] = __int_as_float(tid);
            }
            __syncwarp();
            const int nh = __popc(ballot);
            float* arow = my_acc + j * NCOMP;
#pragma unroll
            for (int p = 0; p < MAXPASS; p++) {
                const int sel = csel[p];
                if (sel < 0) continue;
                const int gch = cgch[p];
                float sum = 0.f;
                for (int hh = 0; hh < nh; hh++) {
                    const float* h = my_hits + hh * HIT_STRIDE;
                    const int pix = __float_as_int(h[11]);
                    sum = fmaf(h[sel], G[pix * GS + gch], sum);
                }
                arow[lane + 32 * p] += sum;
            }
            __syncwarp();
        }
        __syncthreads();
        // combine the warp-private partial sums; one 128-bit vector reduction (REDG.E.ADD.F32x4) per
        // (tile, instance, 4 components) -- 18 per instance at the training shape instead of 69 scalar atomics
        for (int q4 = tid; q4 < nb * (NCOMP / 4); q4 += TILE_PIX) {
            const int q = q4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int wv = 0; wv < NWARP; wv++) {
                const float4 t = *reinterpret_cast<const float4*>(acc + wv * BWD_BATCH * NCOMP + q);
                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
            if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
            const int j = q / NCOMP, k = q - j * NCOMP;
            const int id = ids[j];
            if (k < SVGIR_GEO_GRAD_FLOATS) {
                if (k == 8) { v.y *= 10.f; v.z *= 10.f; v.w *= 10.f; }  // dL_dnormal x10 (backward.cu:804): comps 9..11
                red_add_f32x4(geo_grad + (size_t)id * SVGIR_GEO_GRAD_FLOATS + k, v);
            } else if (k < SVGIR_GEO_GRAD_FLOATS + SP) {
                const int f0 = k - SVGIR_GEO_GRAD_FLOATS;
                float* dst = dL_dfeatures + (size_t)id * S + f0;
                if ((S & 3) == 0) red_add_f32x4(dst, v);
                else {
                    if (f0 + 0 < S && v.x != 0.f) atomicAdd(dst + 0, v.x);
                    if (f0 + 1 < S && v.y != 0.f) atomicAdd(dst + 1, v.y);
                    if (f0 + 2 < S && v.z != 0.f) atomicAdd(dst + 2, v.z);
                    if (f0 + 3 < S && v.w != 0.f) atomicAdd(dst + 3, v.w);
                }
            } else {
                red_add_f32x4(dL_dvfeatures + (size_t)id * (4 * NV) + (k - SVGIR_GEO_GRAD_FLOATS - SP), v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Lane-mode backward compositor, used when 8+S+NV+5 <= 32 (stage-2 training S=4/VS=52, stage 1 S=5).
// Same per-hit math as the kernel above; the reduction is organised so that LANE L OWNS PIXEL-GRADIENT
// CHANNEL L: for every hit of its warp it reads G[pixel][L] once and accumulates
//     a0      += h[sel0(L)] * G     (colour / normal / depth / feature gradient, or a geo scalar with G = 1)
//     a1..a4  += (w*w0..w3) * G     (the four vertex gradients of SV channel L)
// i.e. 5 FMAs for 3 shared-memory loads.  A warp's result for an instance is complete in registers,
// so it goes straight to L2 as one scalar and one 128-bit fire-and-forget reduction instruction per
// (warp, instance) -- no warp-private accumulators, no flush pass, one barrier pair per 32 instances, and
// 46 KB instead of 81 KB of shared memory per CTA (4 resident CTAs per SM instead of 2).
#ifndef LBATCH
#define LBATCH 32
#endif
#ifndef LMINB
#define LMINB 3
#endif
#ifndef LSTAGES
#define LSTAGES 3   // staging ring depth of the lane kernel
#endif
#define HIT_STRIDE_L 12   // 3 float4 per hit: 128-bit stores/loads are conflict-free at this stride

template <int S_T, int NV_T>
struct LaneBwdLayout {
    static constexpr int S = S_T, NV = NV_T;
    static constexpr int SP = (S + 3) & ~3;
    static constexpr int NVP = (NV + 3) & ~3;                 // transposed vfeature rows: [4 vertices][NVP channels]
    static constexpr int STRIDE = SVGIR_REC_FLOATS + SP + 4 * NVP;
    static constexpr int LCH = REC_F4 + SP / 4 + NV;          // float4 chunks loaded per instance
    static constexpr int NG = 8 + S + NV;                     // pixel-gradient row: 1,gC3,gN3,gD',gF,gVF
    static constexpr int GS = NG | 1;                         // odd stride
    static constexpr int G_OFF = LSTAGES * LBATCH * STRIDE;      // stage is a ring of LSTAGES buffers
    static constexpr int HITS_OFF = (G_OFF + TILE_PIX * GS + 3) & ~3;   // 16-B aligned
    static constexpr int IDS_OFF = HITS_OFF + NWARP * 32 * HIT_STRIDE_L;
    static constexpr int SMEM_FLOATS = IDS_OFF + 8 * LBATCH;      // 8-slot ring of staged surfel ids
};

template <int S_T, int NV_T, bool RGSS>
__global__ void __launch_bounds__(TILE_PIX, LMINB) composite_bwd_lane_kernel(
    const svgir_raster_cfg c, const float* __restrict__ features, const float* __restrict__ vfeatures,
    const float4* __restrict__ rec, const uint2* __restrict__ ranges,
    const uint32_t* __restrict__ point_list, const float* __restrict__ final_T,
    const float* __restrict__ final_D, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ gpix_color, const float* __restrict__ gpix_normal,
    const float* __restrict__ gpix_depth, const float* __restrict__ gpix_opac,
    const float* __restrict__ gpix_feature, const float* __restrict__ gpix_vfeature,
    float* __restrict__ geo_grad, float* __restrict__ dL_dfeatures, float* __restrict__ dL_dvfeatures,
    const int32_t* __restrict__ num_rendered, const uint32_t* __restrict__ tile_order) {
    using LY = LaneBwdLayout<S_T, NV_T>;
    constexpr int S = S_T, NV = NV_T;
    constexpr int SP = LY::SP, NVP = LY::NVP, STRIDE = LY::STRIDE, LCH = LY::LCH, NG = LY::NG, GS = LY::GS;
    static_assert(NG + 5 <= 32, "lane mode needs one lane per gradient channel plus 5 geo lanes");

    extern __shared__ __align__(16) float smem[];
    float* stage = smem;                                   // [LSTAGES][LBATCH][STRIDE]
    float* G = smem + LY::G_OFF;                           // [256][GS]
    float* hits = smem + LY::HITS_OFF;                     // [NWARP][32][HIT_STRIDE_L]
    int* ids = reinterpret_cast<int*>(smem + LY::IDS_OFF); // [8][LBATCH]
    __shared__ int tile_max_s;
    __shared__ __align__(8) uint64_t full_bar[LSTAGES];    // batch staged   (32 LDGSTS-completion arrivals: the issuing warp)
    __shared__ __align__(8) uint64_t empty_bar[LSTAGES];   // batch consumed (one arrival per warp)
    __shared__ int ticket[8];                              // per-batch election of the issuing warp

    const int W = c.W, H = c.H;
    const int gx = (W + TILE - 1) / TILE;
    const int tile = (int)tile_order[blockIdx.x];   // heaviest tiles first (binning.cu: tile_scan_kernel)
    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);
    if (total == 0 || num_rendered[1]) return;  // empty tile, or the forward's bins overflowed (nothing valid)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wx0 = (tile % gx) * TILE + (wid & 1) * WARP_PX_W, wy0 = (tile / gx) * TILE + (wid >> 1) * WARP_PX_H;
    const int px = wx0 + (lane & (WARP_PX_W - 1)), py = wy0 + (lane / WARP_PX_W);
    const bool inside = px < W && py < H;
    const size_t HW = (size_t)H * W;
    const size_t pix_id = (size_t)W * py + px;
    const float pxf = (float)px, pyf = (float)py;
    const float wx0f = (float)wx0, wy0f = (float)wy0;

    bool surface = true, ppd = true, normalize_depth = true;
    if (!RGSS) {
        surface = c.n_config > 0 && c.config[0] > 0;
        normalize_depth = c.n_config > 1 && c.config[1] > 0;
        ppd = c.n_config > 2 && c.config[2] > 0;
    }
    const bool sv = surface && ppd;
    const bool feat_to_alpha = !RGSS || c.backward_geometry != 0;

    const float T_final = inside ? final_T[pix_id] : 0.f;
    const float D_final = (inside && normalize_depth) ? final_D[pix_id] : 0.f;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

    // ---- pixel-gradient row (shared: read by the channel lanes; registers: this pixel's own copy) ----
    float* Grow = G + tid * GS;
    float gD = 0.f, gDn = 0.f, Kpix = 0.f;
    float gC[3] = {0, 0, 0}, gN[3] = {0, 0, 0}, gF[S > 0 ? S : 1];
    unsigned long long gV2[NVP > 0 ? NVP / 2 : 1];   // SV-channel pixel gradients as FFMA2 operand pairs
    {
        float gO = 0.f;
        if (inside) {
#pragma unroll
            for (int i = 0; i < 3; i++) gC[i] = gpix_color[i * HW + pix_id];
#pragma unroll
            for (int i = 0; i < 3; i++) gN[i] = surface ? gpix_normal[i * HW + pix_id] : 0.f;
            gD = gpix_depth[pix_id];
            gO = gpix_opac[pix_id];
        }
        const float omT = 1.f - T_final;
        gDn = normalize_depth ? gD / omT : gD;
        Grow[0] = 1.0f;
        Grow[1] = gC[0]; Grow[2] = gC[1]; Grow[3] = gC[2];
        Grow[4] = gN[0]; Grow[5] = gN[1]; Grow[6] = gN[2];   // the x10 of backward.cu:804 is applied at the reduction
        Grow[7] = gDn;
#pragma unroll
        for (int i = 0; i < S; i++) { gF[i] = inside ? gpix_feature[i * HW + pix_id] : 0.f; Grow[8 + i] = gF[i]; }
        float gV[NVP > 0 ? NVP : 1];
#pragma unroll
        for (int i = 0; i < NVP; i++) {
            gV[i] = (i < NV && inside) ? gpix_vfeature[i * HW + pix_id] : 0.f;
            if (i < NV) Grow[8 + S + i] = gV[i];
        }
#pragma unroll
        for (int i = 0; i < NVP / 2; i++) gV2[i] = pack2(gV[2 * i], gV[2 * i + 1]);
        const float* bg = c.bg;
        const float bgdot = bg[0] * gC[0] + bg[1] * gC[1] + bg[2] * gC[2];
        Kpix = gO * T_final - T_final * bgdot;
        if (normalize_depth) Kpix += gD * D_final / omT / omT * -T_final;
        else Kpix += -T_final * (10.f * gD);
    }

    // ---- lane roles ---------------------------------------------------------------------------------
    //   hit record: [0..3] w*w0..w3 | [4] w | [5,6] dmean | [7,8,9] dconic | [10] dopacity | [11] pixel's G-row offset
    //   scalar target of a0: geo_grad component (>=0), feature index (encoded as 16+i), or none (-1)
    int l_gch = 0, l_sel0 = 4, l_k0 = -1;
    bool l_vf = false;
    float l_scale = 1.f;
    if (lane == 0) { l_sel0 = 5; l_k0 = 0; }                                    // dmean.x (G = 1)
    else if (lane < 4) { l_gch = lane; l_k0 = 6 + (lane - 1); }                // colour
    else if (lane < 7) { l_gch = lane; l_k0 = 9 + (lane - 4); l_scale = 10.f; }  // normal, x10 (backward.cu:804)
    else if (lane == 7) { l_gch = 7; l_k0 = 12; }                              // depth
    else if (lane < 8 + S) { l_gch = lane; l_k0 = 16 + (lane - 8); }           // flat features
    else if (lane < NG) { l_gch = lane; l_vf = true; }                         // SV channel lane-8-S
    else if (lane < NG + 5) { l_sel0 = 6 + (lane - NG); l_k0 = 1 + (lane - NG); }  // dmean.y, dconic xyw, dopacity
    const float* Glane = G + l_gch;

    // ---- tile-wide traversal start --------------------------------------------------------------
    if (tid == 0) {
        tile_max_s = 0;
#pragma unroll
        for (int i = 0; i < LSTAGES; i++) { mbar_init(&full_bar[i], 32); mbar_init(&empty_bar[i], TILE_PIX / 32); }
        mbar_fence_init();
    }
    if (tid < 8) ticket[tid] = 0;
    if constexpr (NVP != NV) {  // zero the padding channels of the transposed rows once
        constexpr int PADC = NVP - NV;
        for (int q = tid; q < LSTAGES * LBATCH * 4 * PADC; q += TILE_PIX) {
            const int i = q / (4 * PADC), r = q - i * 4 * PADC;
            stage[i * STRIDE + SVGIR_REC_FLOATS + SP + (r / PADC) * NVP + NV + r % PADC] = 0.f;
        }
    }
    __syncthreads();
    int warp_max;
    {
        int m = last_contributor;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        warp_max = m;
        if (lane == 0) atomicMax(&tile_max_s, m);
    }
    __syncthreads();
    const int tile_max = min(tile_max_s, total);
    if (tile_max == 0) return;

    float* my_hits = hits + wid * 32 * HIT_STRIDE_L;
    float T = T_final, A = 0.f, last_alpha = 0.f, V_last = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    // ---- asynchronous staging ring ------------------------------------------------------------------------
    // batch b = sorted positions top_b-1 ... top_b-nb (back to front), top_b = tile_max - b*LBATCH, lives in buffer
    // b % LSTAGES. There is no CTA-wide barrier in the loop:
    //   * the FIRST warp to finish round k (elected with a ticket) requests batch k+2 and the surfel ids of batch
    //     k+4 with LDGSTS (cp.async: no registers, no scoreboard wait; vfeature rows are transposed on the way by
    //     4-byte copies) -- the fastest warp has the slack, and nobody waits for the slowest warp to issue;
    //   * full_bar[s] completes when that warp's copies have landed (cp.async.mbarrier.arrive.noinc), and is what a
    //     warp waits on before it composites a batch;
    //   * empty_bar[s] counts the 8 warps that finished reading buffer s, and is what the issuing warp waits on
    //     before it overwrites it.
    // Warps drift up to LSTAGES-1 batches apart instead of idling at a __syncthreads (the barrier stall was 2.9 of
    // 9 stall cycles per issue with the CTA-synchronous double buffer).
    auto issue_ids = [&](int b, int t) {
        const int top = tile_max - b * LBATCH;
        const int nb = min(LBATCH, top);
        if (t < nb) cp_async4(ids + (b & 7) * LBATCH + t, point_list + range.x + top - 1 - t);
    };
    auto issue_data = [&](int b, int buf) {   // by one warp
        const int nb = min(LBATCH, tile_max - b * LBATCH);
        float* sb = stage + buf * LBATCH * STRIDE;
        const int* idl = ids + (b & 7) * LBATCH;
        for (int q = lane; q < nb * LCH; q += 32) {
            const int i = q / LCH, ch = q - i * LCH;
            const int id = idl[i];
            float* dst = sb + i * STRIDE;
            if (ch < REC_F4) {
                cp_async16(dst + 4 * ch, rec + (size_t)id * REC_F4 + ch);
            } else if (ch < REC_F4 + SP / 4) {
                const int f0 = (ch - REC_F4) * 4;
                const float* src = features + (size_t)id * S + f0;
                if ((S & 3) == 0) cp_async16(dst + 4 * ch, src);
                else {  // the padding floats [S, SP) of the row are never read
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (f0 + e < S) cp_async4(dst + 4 * ch + e, src + e);
                }
            } else {
                const int cidx = ch - REC_F4 - SP / 4;
                const float* src = vfeatures + (size_t)id * (4 * NV) + 4 * cidx;
                float* t = dst + SVGIR_REC_FLOATS + SP + cidx;   // transpose: vertex-major rows of NVP channels
                cp_async4(t, src); cp_async4(t + NVP, src + 1);
                cp_async4(t + 2 * NVP, src + 2); cp_async4(t + 3 * NVP, src + 3);
            }
        }
    };
    const int nrounds = (tile_max + LBATCH - 1) / LBATCH;
#pragma unroll
    for (int b = 0; b < 4; b++)
        if (b < nrounds) issue_ids(b, tid);
    cp_async_wait_all();
    __syncthreads();   // ids of batches 0..3, the barriers and the tickets are visible
    if (wid < 2 && wid < nrounds) { issue_data(wid, wid); cp_async_mbar_arrive_noinc(&full_bar[wid]); }

    int st = 0;          // kb % LSTAGES
    unsigned ph = 0;     // (kb / LSTAGES) & 1
    for (int kb = 0; kb < nrounds; kb++) {
        const int top = tile_max - kb * LBATCH;
        const int nb = min(LBATCH, top);
        mbar_wait(&full_bar[st], ph);
        const float* sb = stage + st * LBATCH * STRIDE;
        const int* idb = ids + (kb & 7) * LBATCH;

        // per-warp cull: instances behind every pixel's last contributor, or whose alpha >= 1/255 ellipse cannot
        // reach this warp's 8x4 pixels, are never evaluated
        unsigned m;
        {
            bool keep = false;
            if (lane < nb && top - 1 - lane < warp_max) {
                const float4* r = reinterpret_cast<const float4*>(sb + lane * STRIDE);
                const float4 q0 = r[0];
                const float2 q1 = *reinterpret_cast<const float2*>(r + 1);
                keep = footprint_overlaps(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, wx0f, wy0f, WARP_PX_W - 1.f, WARP_PX_H - 1.f);
            }
            m = __ballot_sync(0xffffffffu, keep);
        }

        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const int pos = top - 1 - j;  // 0-based position in the tile's sorted list
            const float4* r = reinterpret_cast<const float4*>(sb + j * STRIDE);
            const float4 q0 = r[0];
            const float4 q1 = r[1];
            PairEval e;
            bool hit = false;
            if (pos < last_contributor) hit = eval_alpha<RGSS>(pxf, pyf, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, e);
            const unsigned ballot = __ballot_sync(0xffffffffu, hit);
            if (ballot == 0) continue;
            if (hit) {
                const float inv1ma = __frcp_rn(1.f - e.alpha);
                T = T * inv1ma;
                const float w = e.alpha * T;
                float depth_k = q1.z;
                float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
                float gzx = 0.f, gzy = 0.f;
                if (sv) {
                    const float4 q2 = r[2];
                    const float4 q3 = r[3];
                    const float u0 = fma_(e.dx, q2.x, mul_(e.dy, q2.y));
                    const float u1 = fma_(e.dx, q2.z, mul_(e.dy, q2.w));
                    depth_k = sub_(q1.z, fma_(q3.x, u0, mul_(q3.y, u1)));
                    gzx = q3.x * q2.x + q3.y * q2.z;   // J6*J0 + J9*J2 (backward.cu:915)
                    gzy = q3.x * q2.y + q3.y * q2.w;   // J6*J1 + J9*J3
                    if (!RGSS) {
                        float u = fmaf(u0, q1.w, 0.5f), v = fmaf(u1, q3.z, 0.5f);
                        u = fminf(0.999f, fmaxf(0.001f, u));
                        v = fminf(0.999f, fmaxf(0.001f, v));
                        w0 = (1.0f - u) * (1.0f - v);
                        w1 = u * (1.0f - v);
                        w2 = (1.0f - u) * v;
                        w3 = u * v;
                    }
                }
                // V = sum_c value_c * g_c over all blended channels (normal uses the plain gN)
                const float4 q4 = r[4];
                float V = q4.x * gC[0] + q4.y * gC[1] + q4.z * gC[2];
                if (surface) {
                    const float4 q5 = r[5];
                    V += q4.w * gN[0] + q5.x * gN[1] + q5.y * gN[2];
                }
                V = fmaf(depth_k, gDn, V);
                const float* f = sb + j * STRIDE + SVGIR_REC_FLOATS;
                if (feat_to_alpha) {
#pragma unroll
                    for (int ch = 0; ch < S; ch++) V = fmaf(f[ch], gF[ch], V);
                }
                if constexpr (NV > 0) {
                    // sum_c gV_c * sum_k w_k vf[4c+k] = sum_k w_k * <vfT[k][:], gV>: channel pairs per FFMA2
                    const float wk[4] = {w0, w1, w2, w3};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const ulonglong2* row = reinterpret_cast<const ulonglong2*>(f + SP + k * NVP);
                        unsigned long long acc = 0ull;
#pragma unroll
                        for (int q = 0; q < NVP / 4; q++) {
                            const ulonglong2 t = row[q];
                            acc = ffma2(t.x, gV2[2 * q], acc);
                            acc = ffma2(t.y, gV2[2 * q + 1], acc);
                        }
                        const float2 d = unpack2(acc);
                        V = fmaf(wk[k], d.x + d.y, V);
                    }
                }
                A = last_alpha * V_last + (1.f - last_alpha) * A;
                V_last = V;
                last_alpha = e.alpha;
                const float dL_dalpha = (V - A) * T + Kpix * inv1ma;
                const float dL_ddist = dL_dalpha * q1.y * -0.5f * e.G;
                float dmx = dL_ddist * 2.f * (q0.z * e.dx + q0.w * e.dy) * ddelx_dx;
                float dmy = dL_ddist * 2.f * (q1.x * e.dy + q0.w * e.dx) * ddely_dy;
                if (sv) { dmx -= gD * gzx; dmy -= gD * gzy; }
                const int rank = __popc(ballot & ((1u << lane) - 1u));
                float4* h4 = reinterpret_cast<float4*>(my_hits + rank * HIT_STRIDE_L);
                h4[0] = make_float4(w * w0, w * w1, w * w2, w * w3);
                h4[1] = make_float4(w, dmx, dmy, dL_ddist * (e.dx * e.dx));
                h4[2] = make_float4(dL_ddist * (e.dx * e.dy), dL_ddist * (e.dy * e.dy), e.G * dL_dalpha, __int_as_float(tid * GS));
            }
            __syncwarp();
            const int nh = __popc(ballot);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll 4
            for (int hh = 0; hh < nh; hh++) {
                const float* h = my_hits + hh * HIT_STRIDE_L;
                const float Gv = Glane[__float_as_int(h[11])];
                a0 = fmaf(h[l_sel0], Gv, a0);
                if (NV > 0) {
                    const float4 hw = *reinterpret_cast<const float4*>(h);
                    a1 = fmaf(hw.x, Gv, a1); a2 = fmaf(hw.y, Gv, a2);
                    a3 = fmaf(hw.z, Gv, a3); a4 = fmaf(hw.w, Gv, a4);
                }
            }
            __syncwarp();
            // this warp's complete contribution to instance j: straight to L2
            const int id = idb[j];
            if (l_k0 >= 16) {
                if (a0 != 0.f) atomicAdd(dL_dfeatures + (size_t)id * S + (l_k0 - 16), a0);
            } else if (l_k0 >= 0) {
                if (a0 != 0.f) atomicAdd(geo_grad + (size_t)id * SVGIR_GEO_GRAD_FLOATS + l_k0, a0 * l_scale);
            } else if (l_vf) {
                red_add_f32x4(dL_dvfeatures + (size_t)id * (4 * NV) + 4 * (lane - 8 - S), make_float4(a1, a2, a3, a4));
            }
        }
        // round epilogue: release buffer st, then request batch kb+2 into the buffer batch kb+2-LSTAGES occupied
        __syncwarp();
        int tk = 1;
        if (lane == 0) {
            mbar_arrive(&empty_bar[st]);
            if (kb + 2 < nrounds) tk = atomicAdd(&ticket[(kb + 2) & 7], 1);   // 8 tickets per batch: first = multiple of 8
        }
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if ((tk & 7) == 0) {   // this warp finished round kb first: it stages batch kb+2
            const int nst = st + 2 >= LSTAGES ? st + 2 - LSTAGES : st + 2;
            const unsigned nph = st + 2 >= LSTAGES ? ph ^ 1u : ph;
            if (kb + 2 >= LSTAGES) mbar_wait(&empty_bar[nst], nph ^ 1u);   // batch kb+2-LSTAGES consumed by all warps
            issue_data(kb + 2, nst);
            if (kb + 4 < nrounds) issue_ids(kb + 4, lane);
            cp_async_mbar_arrive_noinc(&full_bar[nst]);
        }
        if (++st == LSTAGES) { st = 0; ph ^= 1u; }
    }
    cp_async_wait_all();
}

template <int S_T, int NV_T, bool RGSS>
static int launch_lane_bwd(const svgir_raster_cfg& c, const svgir_raster_in& in,
                           const svgir_raster_state& st, svgir_raster_grads& g, cudaStream_t s) {
    const int gx = (c.W + TILE - 1) / TILE, gy = (c.H + TILE - 1) / TILE;
    const size_t smem = sizeof(float) * (size_t)LaneBwdLayout<S_T, NV_T>::SMEM_FLOATS;
    auto k = composite_bwd_lane_kernel<S_T, NV_T, RGSS>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        set_error("composite_bwd: cannot reserve %zu B of shared memory", smem);
        return SVGIR_ERR_CUDA;
    }
    { TimedScope ts_("composite_bwd", s); k<<<gx * gy, TILE_PIX, smem, s>>>(c, in.features, in.vfeatures, (const float4*)st.rec,
                                      (const uint2*)st.ranges, st.point_list, st.final_T, st.final_D,
                                      st.n_contrib, g.dL_dcolor, g.dL_dnormal, g.dL_ddepth,
                                      g.dL_dopacity, g.dL_dfeature, g.dL_dvfeature, g.geo_grad,
                                      g.dL_dfeatures, g.dL_dvfeatures, st.num_rendered, st.big_tiles + 2 + 2 * gx * gy); }
    return check_launch("composite_bwd", c.debug, s);
}

template <int S_T, int NV_T, bool RGSS>
static int launch_one_bwd(const svgir_raster_cfg& c, const svgir_raster_in& in,
                          const svgir_raster_state& st, svgir_raster_grads& g, cudaStream_t s) {
    const int gx = (c.W + TILE - 1) / TILE, gy = (c.H + TILE - 1) / TILE;
    const int SP = (c.S + 3) & ~3;
    const int stride = SVGIR_REC_FLOATS + SP + c.VS;
    const int ncomp = SVGIR_GEO_GRAD_FLOATS + SP + c.VS;
    const int gs = (8 + c.S + c.VS / 4) | 1;
    const size_t smem = sizeof(float) * ((size_t)BWD_BATCH * stride + (size_t)TILE_PIX * gs +
                                         (size_t)NWARP * 32 * HIT_STRIDE +
                                         (size_t)NWARP * BWD_BATCH * ncomp + BWD_BATCH);
    auto k = composite_bwd_kernel<S_T, NV_T, RGSS>;
    if (smem > 48 * 1024) {
        if (smem > 227 * 1024) {
            set_error("composite_bwd: S=%d VS=%d needs %zu B of shared memory (>227 KB)", c.S, c.VS, smem);
            return SVGIR_ERR_INVALID;
        }
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("composite_bwd: cannot reserve %zu B of shared memory", smem);
            return SVGIR_ERR_CUDA;
        }
    }
    { TimedScope ts_("composite_bwd", s); k<<<gx * gy, TILE_PIX, smem, s>>>(c, in.features, in.vfeatures, (const float4*)st.rec,
                                      (const uint2*)st.ranges, st.point_list, st.final_T, st.final_D,
                                      st.n_contrib, g.dL_dcolor, g.dL_dnormal, g.dL_ddepth,
                                      g.dL_dopacity, g.dL_dfeature, g.dL_dvfeature, g.geo_grad,
                                      g.dL_dfeatures, g.dL_dvfeatures, st.num_rendered, st.big_tiles + 2 + 2 * gx * gy); }
    return check_launch("composite_bwd", c.debug, s);
}

int launch_composite_bwd(const svgir_raster_cfg& c, const svgir_raster_in& in,
                         const svgir_raster_state& st, svgir_raster_grads& g, cudaStream_t s) {
    const int NV = c.VS / 4;
    if (c.variant == SVGIR_VARIANT_RGSS) {
        if (c.S == 5) return launch_lane_bwd<5, 0, true>(c, in, st, g, s);
        return launch_one_bwd<-1, -1, true>(c, in, st, g, s);
    }
    // lane mode needs 8+S+NV+5 <= 32 lanes; the eval shape (7,16) and generic shapes use the component-owner kernel
    if (c.S == 4 && NV == 13) return launch_lane_bwd<4, 13, false>(c, in, st, g, s);
    if (c.S == 7 && NV == 16) return launch_one_bwd<7, 16, false>(c, in, st, g, s);
    if (c.S == 0 && NV == 0) return launch_lane_bwd<0, 0, false>(c, in, st, g, s);
    return launch_one_bwd<-1, -1, false>(c, in, st, g, s);
}

}  // namespace svgir
