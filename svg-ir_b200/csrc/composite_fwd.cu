// Forward compositing: one CTA per 16x16 tile, one thread per pixel, front-to-back alpha blending
// of colour, geometric normal, depth, S flat features and VS/4 bilinearly interpolated
// spatially-varying features.
//
// Replaces renderCUDA<3> (svgss_rasterization/cuda_rasterizer/forward.cu:402-750; stage 1:
// rgss-rasterization/cuda_rasterizer/forward.cu:324-535).  Differences in *how*, not *what*:
//   * the whole per-instance payload (24-float packed record + the surfel's feature and vfeature
//     rows) is staged into shared memory with 128-bit loads, so the per-hit global reads of
//     features/vfeatures (forward.cu:635-646) disappear;
//   * (S, VS/4) are template parameters for the shapes the application uses, so accumulators live
//     in registers instead of the reference's 304 B local-memory stack;
//   * warps skip instances none of their 32 pixels hit;
//   * out_weights is reduced per warp and per tile before one global atomic per (tile, instance)
//     instead of one per (pixel, hit) (forward.cu:653).
// The alpha chain (power, exp, alpha, T test) uses the reference build's exact roundings.
//
// Roofline: algorithmic bytes = R*(104+4S+4VS) + HW*(4*(8+S+VS/4)+12); the kernel is
// issue-bound, not HBM-bound, whenever tiles terminate early.
#include "common.cuh"

namespace svgir {

#define FWD_BATCH 32   // 32 vs 64 instances per batch: composite_fwd<4,13> 0.351 -> 0.344 ms, <7,16> 0.432 -> 0.409 ms

template <int S_T, int NV_T, bool RGSS>
__global__ void __launch_bounds__(TILE_PIX, S_T >= 0 && S_T <= 8 ? 4 : 1) composite_fwd_kernel(
    const svgir_raster_cfg c, const float* __restrict__ features, const float* __restrict__ vfeatures,
    const float4* __restrict__ rec, const uint2* __restrict__ ranges,
    const uint32_t* __restrict__ point_list, const int32_t* __restrict__ num_rendered,
    float* __restrict__ final_T, float* __restrict__ final_D, uint32_t* __restrict__ n_contrib,
    float* __restrict__ out_color, float* __restrict__ out_normal, float* __restrict__ out_depth,
    float* __restrict__ out_opac, float* __restrict__ out_feature, float* __restrict__ out_vfeature,
    float* __restrict__ out_weights, const uint32_t* __restrict__ tile_order) {
    constexpr bool GENERIC = S_T < 0;
    // PACKED: the staged vfeature row is transposed to [4 vertices][NVP channels] so that channel PAIRS are
    // accumulated with one FFMA2 each (2 fused multiply-adds per issue slot)
    constexpr bool PACKED = !GENERIC && NV_T > 0;
    constexpr int MAXS = GENERIC ? SVGIR_MAX_S : (S_T > 0 ? S_T : 1);
    constexpr int MAXNV = GENERIC ? SVGIR_MAX_NV : (NV_T > 0 ? NV_T : 1);
    constexpr int NVP_T = PACKED ? ((NV_T + 3) & ~3) : 0;   // channels padded to a float4
    const int S = GENERIC ? c.S : S_T;
    const int NV = GENERIC ? c.VS / 4 : NV_T;
    const int SP = (S + 3) & ~3;          // feature row padded to float4
    const int STRIDE = SVGIR_REC_FLOATS + SP + (PACKED ? 4 * NVP_T : 4 * NV);  // floats per staged instance
    const int LCH = REC_F4 + SP / 4 + NV;  // float4 chunks loaded per instance

    extern __shared__ __align__(16) float smem[];
    float* stage = smem;                                  // [2][FWD_BATCH][STRIDE]  double-buffered (LDGSTS pipeline)
    float* wsum = smem + 2 * FWD_BATCH * STRIDE;          // [2][8 warps][FWD_BATCH] warp-private blend-weight sums
    int* ids = reinterpret_cast<int*>(wsum + 2 * 8 * FWD_BATCH);  // [4][FWD_BATCH] ring of staged surfel ids

    if (num_rendered[1]) return;  // binning overflowed: nothing valid to render
    const int W = c.W, H = c.H;
    const int gx = (W + TILE - 1) / TILE;
    const int tile = (int)tile_order[blockIdx.x];   // heaviest tiles first (binning.cu: tile_scan_kernel)
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

    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);

    float T = 1.0f, D = 0.f;
    float C[3] = {0, 0, 0}, N[3] = {0, 0, 0};
    float F[MAXS], VF[PACKED ? 1 : MAXNV];
    unsigned long long VF2[PACKED ? NVP_T / 2 : 1];
#pragma unroll
    for (int i = 0; i < MAXS; i++) F[i] = 0.f;
    if (PACKED) {
#pragma unroll
        for (int i = 0; i < NVP_T / 2; i++) VF2[i] = 0ull;
    } else {
#pragma unroll
        for (int i = 0; i < MAXNV; i++) VF[i] = 0.f;
    }
    uint32_t last_contributor = 0;
    bool done = !inside;

    for (int i = tid; i < 2 * 8 * FWD_BATCH; i += TILE_PIX) wsum[i] = 0.f;
    if (PACKED && NVP_T != NV_T) {  // zero the padding channels of the transposed rows once
        constexpr int PADC = NVP_T > NV_T ? NVP_T - NV_T : 1;
        for (int q = tid; q < 2 * FWD_BATCH * 4 * PADC; q += TILE_PIX) {
            const int i = q / (4 * PADC), r = q - i * 4 * PADC;
            stage[i * STRIDE + SVGIR_REC_FLOATS + SP + (r / PADC) * NVP_T + NV_T + r % PADC] = 0.f;
        }
    }

    // ---- asynchronous staging pipeline -----------------------------------------------------------
    // While batch k is composited, the records + feature rows of batch k+1 stream into the other stage
    // buffer and the surfel ids of batch k+2 into the id ring, all with LDGSTS (cp.async: no registers,
    // no scoreboard wait); the vfeature rows are transposed on the way by 4-byte copies. One barrier per
    // batch: it publishes batch k and retires every warp's reads of batch k-1.
    auto issue_ids = [&](int base, int slot) {
        const int nb = min(FWD_BATCH, total - base);
        if (tid < nb) cp_async4(ids + slot * FWD_BATCH + tid, point_list + range.x + base + tid);
    };
    auto issue_data = [&](int base, int slot, int buf) {
        const int nb = min(FWD_BATCH, total - base);
        float* sb = stage + buf * FWD_BATCH * STRIDE;
        const int* idl = ids + slot * FWD_BATCH;
        for (int q = tid; q < nb * LCH; q += TILE_PIX) {
            const int i = q / LCH, ch = q - i * LCH;
            const int id = idl[i];
            float* dst = sb + i * STRIDE;
            if (ch < REC_F4) {
                cp_async16(dst + 4 * ch, rec + (size_t)id * REC_F4 + ch);
            } else if (ch < REC_F4 + SP / 4) {
                const int f0 = (ch - REC_F4) * 4;
                const float* src = features + (size_t)id * S + f0;
                if ((S & 3) == 0) cp_async16(dst + 4 * ch, src);
                else {
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        if (f0 + e < S) cp_async4(dst + 4 * ch + e, src + e);
                        else dst[4 * ch + e] = 0.f;
                    }
                }
            } else {
                const int cidx = ch - REC_F4 - SP / 4;
                const float* src = vfeatures + (size_t)id * (4 * NV) + 4 * cidx;
                if (PACKED) {  // transpose: vertex-major rows of NVP channels
                    float* t = dst + SVGIR_REC_FLOATS + SP + cidx;
                    cp_async4(t, src); cp_async4(t + NVP_T, src + 1);
                    cp_async4(t + 2 * NVP_T, src + 2); cp_async4(t + 3 * NVP_T, src + 3);
                } else {
                    cp_async16(dst + 4 * ch, src);
                }
            }
        }
    };
    if (total > 0) {
        issue_ids(0, 0);
        if (FWD_BATCH < total) issue_ids(FWD_BATCH, 1);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        issue_data(0, 0, 0);
        cp_async_commit();
    }

    int nbatch = 0;
    for (int base = 0; base < total; base += FWD_BATCH, nbatch++) {
        const int par = nbatch & 1;
        cp_async_wait<0>();
        const int ndone = __syncthreads_count(done);
        // flush the previous batch's per-instance weight sums (one atomic per tile x instance)
        if (nbatch > 0 && tid < FWD_BATCH) {
            float* ws = wsum + (par ^ 1) * 8 * FWD_BATCH + tid;
            float wv = 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) { wv += ws[k * FWD_BATCH]; ws[k * FWD_BATCH] = 0.f; }
            if (wv != 0.f) atomicAdd(&out_weights[ids[((nbatch - 1) & 3) * FWD_BATCH + tid]], wv);
        }
        if (ndone == TILE_PIX) break;
        const int nb = min(FWD_BATCH, total - base);
        if (base + FWD_BATCH < total) issue_data(base + FWD_BATCH, (nbatch + 1) & 3, par ^ 1);
        if (base + 2 * FWD_BATCH < total) issue_ids(base + 2 * FWD_BATCH, (nbatch + 2) & 3);
        cp_async_commit();
        const float* sb = stage + par * FWD_BATCH * STRIDE;

        // per-warp footprint cull: one bit per staged instance whose alpha >= 1/255 ellipse can reach this
        // warp's 8x4 pixels; the others are never evaluated
        unsigned masks[FWD_BATCH / 32];
        const bool warp_live = !__all_sync(0xffffffffu, done);
#pragma unroll
        for (int h = 0; h < FWD_BATCH / 32; h++) {
            const int i = h * 32 + lane;
            bool keep = false;
            if (warp_live && i < nb) {
                const float4* r = reinterpret_cast<const float4*>(sb + i * STRIDE);
                const float4 q0 = r[0];
                const float2 q1 = *reinterpret_cast<const float2*>(r + 1);
                keep = footprint_overlaps(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, wx0f, wy0f, WARP_PX_W - 1.f, WARP_PX_H - 1.f);
            }
            masks[h] = __ballot_sync(0xffffffffu, keep);
        }

#pragma unroll
        for (int h = 0; h < FWD_BATCH / 32; h++) {
            unsigned m = masks[h];
            while (m) {
                const int j = h * 32 + __ffs(m) - 1;
                m &= m - 1;
                const float4* r = reinterpret_cast<const float4*>(sb + j * STRIDE);
                const float4 q0 = r[0];
                const float4 q1 = r[1];
                PairEval e;
                bool hit = false;
                float test_T = 0.f;
                if (!done) {
                    hit = eval_alpha<RGSS>(pxf, pyf, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, e);
                    if (hit) {
                        test_T = mul_(T, sub_(1.f, e.alpha));
                        if (test_T < 0.0001f) { done = true; hit = false; }
                    }
                }
                if (!__any_sync(0xffffffffu, hit)) continue;
                float w = 0.f;
                if (hit) {
                    w = mul_(e.alpha, T);
                    float depth_k = q1.z;
                    float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
                    if (sv) {
                        const float4 q2 = r[2];
                        const float4 q3 = r[3];
                        const float u0 = fma_(e.dx, q2.x, mul_(e.dy, q2.y));
                        const float u1 = fma_(e.dx, q2.z, mul_(e.dy, q2.w));
                        depth_k = sub_(q1.z, fma_(q3.x, u0, mul_(q3.y, u1)));
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
                    D = fmaf(depth_k, w, D);
                    const float4 q4 = r[4];
                    C[0] = fmaf(q4.x, w, C[0]);
                    C[1] = fmaf(q4.y, w, C[1]);
                    C[2] = fmaf(q4.z, w, C[2]);
                    if (surface) {
                        const float4 q5 = r[5];
                        N[0] = fmaf(q4.w, w, N[0]);
                        N[1] = fmaf(q5.x, w, N[1]);
                        N[2] = fmaf(q5.y, w, N[2]);
                    }
                    const float* f = sb + j * STRIDE + SVGIR_REC_FLOATS;
#pragma unroll
                    for (int ch = 0; ch < MAXS; ch++)
                        if (ch < S) F[ch] = fmaf(f[ch], w, F[ch]);
                    if (PACKED) {
                        // VF_c += (w*w_k) * vf[4c+k], two channels per FFMA2 (reference: w * sum_k w_k vf[4c+k])
                        const float wk[4] = {w * w0, w * w1, w * w2, w * w3};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const unsigned long long wk2 = pack2(wk[k], wk[k]);
                            const ulonglong2* row = reinterpret_cast<const ulonglong2*>(f + SP + k * NVP_T);
#pragma unroll
                            for (int q = 0; q < NVP_T / 4; q++) {
                                const ulonglong2 t = row[q];
                                VF2[2 * q] = ffma2(t.x, wk2, VF2[2 * q]);
                                VF2[2 * q + 1] = ffma2(t.y, wk2, VF2[2 * q + 1]);
                            }
                        }
                    } else {
                        const float4* vf = reinterpret_cast<const float4*>(f + SP);
#pragma unroll
                        for (int cidx = 0; cidx < MAXNV; cidx++)
                            if (cidx < NV) {
                                const float4 t = vf[cidx];
                                const float s4 = ((t.x * w0 + t.y * w1) + t.z * w2) + t.w * w3;
                                VF[cidx] = fmaf(w, s4, VF[cidx]);
                            }
                    }
                    T = test_T;
                    last_contributor = (uint32_t)(base + j + 1);
                }
                const float wt = warp_sum(w);
                if (lane == 0) wsum[(par * 8 + wid) * FWD_BATCH + j] = wt;
            }
        }
    }
    // flush the last computed batch's weight sums (already flushed and zeroed if the loop broke out early)
    cp_async_wait<0>();
    __syncthreads();
    if (nbatch > 0 && tid < FWD_BATCH) {
        const int par = (nbatch - 1) & 1;
        float wv = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) wv += wsum[(par * 8 + k) * FWD_BATCH + tid];
        if (wv != 0.f) atomicAdd(&out_weights[ids[((nbatch - 1) & 3) * FWD_BATCH + tid]], wv);
    }

    if (inside) {
        T = fminf(1.f - 0.000001f, T);  // forward.cu:671 (double literal, rounds to 0.999999f)
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        const float* bg = c.bg;
        out_color[0 * HW + pix_id] = fmaf(T, bg[0], C[0]);
        out_color[1 * HW + pix_id] = fmaf(T, bg[1], C[1]);
        out_color[2 * HW + pix_id] = fmaf(T, bg[2], C[2]);
#pragma unroll
        for (int ch = 0; ch < MAXS; ch++)
            if (ch < S) out_feature[ch * HW + pix_id] = F[ch];
        if (PACKED) {
#pragma unroll
            for (int i = 0; i < NVP_T / 2; i++) {
                const float2 v = unpack2(VF2[i]);
                if (2 * i < NV_T) out_vfeature[(size_t)(2 * i) * HW + pix_id] = v.x;
                if (2 * i + 1 < NV_T) out_vfeature[(size_t)(2 * i + 1) * HW + pix_id] = v.y;
            }
        } else {
#pragma unroll
            for (int cidx = 0; cidx < MAXNV; cidx++)
                if (cidx < NV) out_vfeature[cidx * HW + pix_id] = VF[cidx];
        }
        out_normal[0 * HW + pix_id] = surface ? N[0] : 0.f;
        out_normal[1 * HW + pix_id] = surface ? N[1] : 0.f;
        out_normal[2 * HW + pix_id] = surface ? N[2] : 0.f;
        out_depth[pix_id] = normalize_depth ? __fdiv_rn(D, 1.f - T) : fmaf(T, 10.f, D);
        out_opac[pix_id] = 1.f - T;
        if (normalize_depth) final_D[pix_id] = D;
    }
}

template <int S_T, int NV_T, bool RGSS>
static int launch_one(const svgir_raster_cfg& c, const svgir_raster_in& in, svgir_raster_state& st,
                      svgir_raster_out& out, cudaStream_t s) {
    const int gx = (c.W + TILE - 1) / TILE, gy = (c.H + TILE - 1) / TILE;
    const int SP = (c.S + 3) & ~3;
    const int nvf = (S_T >= 0 && NV_T > 0) ? 4 * ((NV_T + 3) & ~3) : c.VS;  // packed kernels pad the transposed rows
    const int stride = SVGIR_REC_FLOATS + SP + nvf;
    const size_t smem = sizeof(float) * ((size_t)2 * FWD_BATCH * stride + 2 * 8 * FWD_BATCH + 4 * FWD_BATCH);
    auto k = composite_fwd_kernel<S_T, NV_T, RGSS>;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("composite_fwd: cannot reserve %zu B of shared memory", smem);
            return SVGIR_ERR_CUDA;
        }
    }
    { TimedScope ts_("composite_fwd", s); k<<<gx * gy, TILE_PIX, smem, s>>>(c, in.features, in.vfeatures, (const float4*)st.rec,
                                      (const uint2*)st.ranges, st.point_list, st.num_rendered,
                                      st.final_T, st.final_D, st.n_contrib, out.color, out.normal,
                                      out.depth, out.opacity, out.feature, out.vfeature, out.weights,
                                      st.big_tiles + 2 + 2 * gx * gy); }
    return check_launch("composite_fwd", c.debug, s);
}

int launch_composite_fwd(const svgir_raster_cfg& c, const svgir_raster_in& in, svgir_raster_state& st,
                         svgir_raster_out& out, cudaStream_t s) {
    const int NV = c.VS / 4;
    if (c.variant == SVGIR_VARIANT_RGSS) {
        if (c.S == 5) return launch_one<5, 0, true>(c, in, st, out, s);
        return launch_one<-1, -1, true>(c, in, st, out, s);
    }
    if (c.S == 4 && NV == 13) return launch_one<4, 13, false>(c, in, st, out, s);   // stage-2 training
    if (c.S == 7 && NV == 16) return launch_one<7, 16, false>(c, in, st, out, s);   // relight eval
    if (c.S == 0 && NV == 0) return launch_one<0, 0, false>(c, in, st, out, s);
    return launch_one<-1, -1, false>(c, in, st, out, s);
}

// ---- stage-1 screen-space helpers (rgss-rasterization/cuda_rasterizer/forward.cu:538-631) ------
// surface_xyz = back-projected (depth/opacity); pseudo normal = -normalize(cross(Sobel_x, Sobel_y))
// rotated to world space. Only run when computer_pseudo_normal is set.
__global__ void __launch_bounds__(256) surface_xyz_kernel(int W, int H, float fx, float fy, float cx, float cy,
                                                          const float* __restrict__ opac,
                                                          const float* __restrict__ depth,
                                                          float* __restrict__ xyz) {
    const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (px >= W || py >= H) return;
    const size_t HW = (size_t)H * W, id = (size_t)W * py + px;
    const float d = depth[id] / fmaxf(opac[id], 0.0000001f);
    xyz[id] = (px - cx) / fx * d;
    xyz[HW + id] = (py - cy) / fy * d;
    xyz[2 * HW + id] = d;
}

__global__ void __launch_bounds__(256) pseudo_normal_kernel(int W, int H, const float* __restrict__ V,
                                                            const float* __restrict__ xyz,
                                                            float* __restrict__ normals) {
    const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (px >= W || py >= H) return;
    const size_t HW = (size_t)H * W;
    const int xm = px == 0 ? 0 : px - 1, xp = px == W - 1 ? W - 1 : px + 1;
    const int ym = py == 0 ? 0 : py - 1, yp = py == H - 1 ? H - 1 : py + 1;
    float ga[3], gb[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float* p = xyz + i * HW;
        const float v00 = p[(size_t)W * ym + xm], v01 = p[(size_t)W * ym + px], v02 = p[(size_t)W * ym + xp];
        const float v10 = p[(size_t)W * py + xm], v12 = p[(size_t)W * py + xp];
        const float v20 = p[(size_t)W * yp + xm], v21 = p[(size_t)W * yp + px], v22 = p[(size_t)W * yp + xp];
        ga[i] = -0.125f * v00 + 0.125f * v02 - 0.25f * v10 + 0.25f * v12 - 0.125f * v20 + 0.125f * v22;
        gb[i] = -0.125f * v00 - 0.25f * v01 - 0.125f * v02 + 0.125f * v20 + 0.25f * v21 + 0.125f * v22;
    }
    float n[3] = {ga[1] * gb[2] - ga[2] * gb[1], -ga[0] * gb[2] + ga[2] * gb[0], ga[0] * gb[1] - ga[1] * gb[0]};
    const float norm = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (norm <= 0.00000f) return;
    n[0] = -n[0] / norm; n[1] = -n[1] / norm; n[2] = -n[2] / norm;
    const size_t id = (size_t)W * py + px;
    normals[id] = V[0] * n[0] + V[1] * n[1] + V[2] * n[2];
    normals[HW + id] = V[4] * n[0] + V[5] * n[1] + V[6] * n[2];
    normals[2 * HW + id] = V[8] * n[0] + V[9] * n[1] + V[10] * n[2];
}

int launch_pseudo_normal(const svgir_raster_cfg& c, svgir_raster_out& out, cudaStream_t s) {
    const dim3 grid((c.W + TILE - 1) / TILE, (c.H + TILE - 1) / TILE);
    const float fy = c.H / (2.0f * c.tan_fovy), fx = c.W / (2.0f * c.tan_fovx);
    { TimedScope ts_("surface_xyz", s);
      surface_xyz_kernel<<<grid, 256, 0, s>>>(c.W, c.H, fx, fy, c.cx, c.cy, out.opacity, out.depth, out.surface_xyz); }
    { TimedScope ts_("pseudo_normal", s);
      pseudo_normal_kernel<<<grid, 256, 0, s>>>(c.W, c.H, c.viewmatrix, out.surface_xyz, out.pseudo_normal); }
    return check_launch("pseudo_normal", c.debug, s);
}

}  // namespace svgir
