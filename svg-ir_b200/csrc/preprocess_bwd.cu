// Backward preprocess: conic -> cov2D -> cov3D/mean, mean2D/depth -> mean3D, SH colour -> SH
// coefficients (+ view direction -> mean), cov3D -> scale / quaternion (+ normal -> R[:,2]).
// One thread per surfel; one kernel instead of the reference's computeCov2DCUDA +
// preprocessCUDA pair (svgss_rasterization/cuda_rasterizer/backward.cu:163-322, 438-526, with
// helpers :20-158 and :326-432), and it also materialises the API-level gradient tensors
// (dL_dmeans2D, dL_dcolors, dL_dopacities) from the packed accumulator row, writing zeros for
// culled surfels so the host allocates with empty() instead of 17 zero-filled tensors
// (rasterize_points.cu:195-211).
//
// Camera-pose gradients (`lrn_cam`, config[3] > 0) are not produced: the reference model passes
// a 3-entry config so that flag is an out-of-bounds read; the returned
// dL_dviewmat / dL_dprojmat / dL_dcampos are zeros.
//
// Roofline: HBM-bound, ~ (64 + 12*M + 70) B read + (100 + 12*M) B written per surfel.
#include "common.cuh"

namespace svgir {

__device__ __constant__ float bSHC0 = 0.28209479177387814f;
__device__ __constant__ float bSHC1 = 0.4886025119029199f;
__device__ __constant__ float bSHC2[5] = {1.0925484305920792f, -1.0925484305920792f,
                                          0.31539156525252005f, -1.0925484305920792f,
                                          0.5462742152960396f};
__device__ __constant__ float bSHC3[7] = {-0.5900435899266435f, 2.890611442640554f,
                                          -0.4570457994644658f, 0.3731763325901154f,
                                          -0.4570457994644658f, 1.445305721320277f,
                                          -0.5900435899266435f};

// ACC = false: the reference-shaped call (svgir_raster_backward): one thread per surfel of [0,P), every output row is
//   written (zeros for culled surfels).
// ACC = true (svgir_raster_backward_params): one thread per entry of the visible-surfel list; parameter gradients are
//   ADDED into caller-zeroed buffers (`pg`), the API-level intermediates (dL_dcolors, dL_dcov3D, ...) are not
//   materialised, and an overflowed forward contributes nothing.
#define PB_STORE(ptr, val) do { if (ACC) *(ptr) += (val); else *(ptr) = (val); } while (0)
template <bool RGSS, bool ACC>
__global__ void __launch_bounds__(256) preprocess_bwd_kernel(
    const svgir_raster_cfg c, const svgir_raster_in in, const float* __restrict__ cov3Ds,
    const uint8_t* __restrict__ clamped, const int32_t* __restrict__ radii,
    const float* __restrict__ geo_grad, svgir_raster_grads g, const svgir_param_grads pg,
    const int32_t* __restrict__ vis_list, const int32_t* __restrict__ vis_count,
    const int32_t* __restrict__ num_rendered) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (ACC) {
        if (num_rendered[1] != 0 || idx >= min(__ldg(vis_count), c.P)) return;
        idx = __ldg(vis_list + idx);
        g.dL_dsh = pg.d_sh; g.dL_dscales = pg.d_scales; g.dL_drotations = pg.d_rotations;
    }
    if (idx >= c.P) return;
    const int M = c.M;
    const bool visible = ACC || radii[idx] > 0;
    if (!visible) {
        // reference leaves the torch::zeros initialisation untouched for culled surfels
#pragma unroll
        for (int i = 0; i < 3; i++) {
            g.dL_dmeans2D[3 * idx + i] = 0.f; g.dL_dcolors[3 * idx + i] = 0.f;
            g.dL_dmeans3D[3 * idx + i] = 0.f; g.dL_dscales[3 * idx + i] = 0.f;
        }
        g.dL_dopacities[idx] = 0.f;
#pragma unroll
        for (int i = 0; i < 6; i++) g.dL_dcov3D[6 * idx + i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) g.dL_drotations[4 * idx + i] = 0.f;
        for (int i = 0; i < 3 * M; i++) g.dL_dsh[(size_t)idx * 3 * M + i] = 0.f;
        if (g.dL_dconic) for (int i = 0; i < 4; i++) g.dL_dconic[4 * idx + i] = 0.f;
        if (g.dL_dnormal3) for (int i = 0; i < 3; i++) g.dL_dnormal3[3 * idx + i] = 0.f;
        if (g.dL_ddepths) g.dL_ddepths[idx] = 0.f;
        return;
    }
    const float4* gg = reinterpret_cast<const float4*>(geo_grad + (size_t)idx * SVGIR_GEO_GRAD_FLOATS);
    const float4 g0 = gg[0], g1 = gg[1], g2 = gg[2], g3 = gg[3];
    const float dm2x = g0.x, dm2y = g0.y;
    const float dcon[3] = {g0.z, g0.w, g1.x};
    const float dopac = g1.y;
    const float dcol[3] = {g1.z, g1.w, g2.x};
    const float dnrm[3] = {g2.y, g2.z, g2.w};
    const float ddep = g3.x;
    if (ACC) {
        if (pg.d_means2D) { pg.d_means2D[3 * idx] += dm2x; pg.d_means2D[3 * idx + 1] += dm2y; }
        if (pg.d_opacities) pg.d_opacities[idx] += dopac;
    } else {
        g.dL_dmeans2D[3 * idx] = dm2x; g.dL_dmeans2D[3 * idx + 1] = dm2y; g.dL_dmeans2D[3 * idx + 2] = 0.f;
        g.dL_dopacities[idx] = dopac;
#pragma unroll
        for (int i = 0; i < 3; i++) g.dL_dcolors[3 * idx + i] = dcol[i];
    }
    if (!ACC && g.dL_dconic) {
        g.dL_dconic[4 * idx] = dcon[0]; g.dL_dconic[4 * idx + 1] = dcon[1];
        g.dL_dconic[4 * idx + 2] = 0.f; g.dL_dconic[4 * idx + 3] = dcon[2];
    }
    if (!ACC && g.dL_dnormal3) for (int i = 0; i < 3; i++) g.dL_dnormal3[3 * idx + i] = dnrm[i];
    if (!ACC && g.dL_ddepths) g.dL_ddepths[idx] = ddep;

    const float* __restrict__ V = c.viewmatrix;
    const float* __restrict__ PV = c.projmatrix;
    const float fy = c.H / (2.0f * c.tan_fovy), fx = c.W / (2.0f * c.tan_fovx);
    const float mx = in.means3D[3 * idx], my = in.means3D[3 * idx + 1], mz = in.means3D[3 * idx + 2];
    const float* cov3D = (in.cov3D_precomp ? in.cov3D_precomp : cov3Ds) + 6 * (size_t)idx;

    // ---- computeCov2DCUDA (backward.cu:163-322) ----
    float t[3];
#pragma unroll
    for (int i = 0; i < 3; i++) t[i] = V[i] * mx + V[4 + i] * my + V[8 + i] * mz + V[12 + i];
    const float limx = 1.3f * c.tan_fovx, limy = 1.3f * c.tan_fovy;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float J0 = fx / t[2], J1 = -(fx * t[0]) / (t[2] * t[2]);
    const float J2 = fy / t[2], J3 = -(fy * t[1]) / (t[2] * t[2]);
    // T = W*J with W columns (V0,V4,V8),(V1,V5,V9),(V2,V6,V10); T[c][r]
    float T0[3], T1[3];
    T0[0] = V[0] * J0 + V[2] * J1; T0[1] = V[4] * J0 + V[6] * J1; T0[2] = V[8] * J0 + V[10] * J1;
    T1[0] = V[1] * J2 + V[2] * J3; T1[1] = V[5] * J2 + V[6] * J3; T1[2] = V[9] * J2 + V[10] * J3;
    const float Vrk[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]},
                             {cov3D[2], cov3D[4], cov3D[5]}};
    // TV0[k] = sum_m T0[m]*Vrk[k][m] ; TV1 likewise
    float TV0[3], TV1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        TV0[k] = T0[0] * Vrk[k][0] + T0[1] * Vrk[k][1] + T0[2] * Vrk[k][2];
        TV1[k] = T1[0] * Vrk[k][0] + T1[1] * Vrk[k][1] + T1[2] * Vrk[k][2];
    }
    const float a = (TV0[0] * T0[0] + TV0[1] * T0[1] + TV0[2] * T0[2]) + 0.3f;
    const float b = TV1[0] * T0[0] + TV1[1] * T0[1] + TV1[2] * T0[2];
    const float cc = (TV1[0] * T1[0] + TV1[1] * T1[1] + TV1[2] * T1[2]) + 0.3f;
    const float denom = a * cc - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dcv[6];
    if (denom2inv != 0) {
        dL_da = denom2inv * (-cc * cc * dcon[0] + 2 * b * cc * dcon[1] + (denom - a * cc) * dcon[2]);
        dL_dc = denom2inv * (-a * a * dcon[2] + 2 * a * b * dcon[1] + (denom - a * cc) * dcon[0]);
        dL_db = denom2inv * 2 * (b * cc * dcon[0] - (denom + 2 * b * b) * dcon[1] + a * b * dcon[2]);
        dcv[0] = (T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc);
        dcv[3] = (T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc);
        dcv[5] = (T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc);
        dcv[1] = 2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2 * T1[0] * T1[1] * dL_dc;
        dcv[2] = 2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2 * T1[0] * T1[2] * dL_dc;
        dcv[4] = 2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2 * T1[1] * T1[2] * dL_dc;
    } else {
#pragma unroll
        for (int i = 0; i < 6; i++) dcv[i] = 0.f;
    }
    if (!ACC) {
#pragma unroll
        for (int i = 0; i < 6; i++) g.dL_dcov3D[6 * idx + i] = dcv[i];
    }
    const float dL_dT00 = 2 * TV0[0] * dL_da + TV1[0] * dL_db;
    const float dL_dT01 = 2 * TV0[1] * dL_da + TV1[1] * dL_db;
    const float dL_dT02 = 2 * TV0[2] * dL_da + TV1[2] * dL_db;
    const float dL_dT10 = 2 * TV1[0] * dL_dc + TV0[0] * dL_db;
    const float dL_dT11 = 2 * TV1[1] * dL_dc + TV0[1] * dL_db;
    const float dL_dT12 = 2 * TV1[2] * dL_dc + TV0[2] * dL_db;
    const float dL_dJ00 = V[0] * dL_dT00 + V[4] * dL_dT01 + V[8] * dL_dT02;
    const float dL_dJ02 = V[2] * dL_dT00 + V[6] * dL_dT01 + V[10] * dL_dT02;
    const float dL_dJ11 = V[1] * dL_dT10 + V[5] * dL_dT11 + V[9] * dL_dT12;
    const float dL_dJ12 = V[2] * dL_dT10 + V[6] * dL_dT11 + V[10] * dL_dT12;
    const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const float dL_dtx = x_grad_mul * -fx * tz2 * dL_dJ02;
    const float dL_dty = y_grad_mul * -fy * tz2 * dL_dJ12;
    const float dL_dtz = -fx * tz2 * dL_dJ00 - fy * tz2 * dL_dJ11 + (2 * fx * t[0]) * tz3 * dL_dJ02 +
                         (2 * fy * t[1]) * tz3 * dL_dJ12;
    float dm[3] = {V[0] * dL_dtx + V[1] * dL_dty + V[2] * dL_dtz,
                   V[4] * dL_dtx + V[5] * dL_dty + V[6] * dL_dtz,
                   V[8] * dL_dtx + V[9] * dL_dty + V[10] * dL_dtz};

    // ---- preprocessCUDA backward (backward.cu:438-526) ----
    const float m_w = 1.0f / ((PV[3] * mx + PV[7] * my + PV[11] * mz + PV[15]) + 0.0000001f);
    const float mul1 = (PV[0] * mx + PV[4] * my + PV[8] * mz + PV[12]) * m_w * m_w;
    const float mul2 = (PV[1] * mx + PV[5] * my + PV[9] * mz + PV[13]) * m_w * m_w;
    dm[0] += (PV[0] * m_w - PV[3] * mul1) * dm2x + (PV[1] * m_w - PV[3] * mul2) * dm2y + ddep * V[2];
    dm[1] += (PV[4] * m_w - PV[7] * mul1) * dm2x + (PV[5] * m_w - PV[7] * mul2) * dm2y + ddep * V[6];
    dm[2] += (PV[8] * m_w - PV[11] * mul1) * dm2x + (PV[9] * m_w - PV[11] * mul2) * dm2y + ddep * V[10];

    if (in.shs) {
        // computeColorFromSH backward (backward.cu:20-158)
        const float ox = mx - c.campos[0], oy = my - c.campos[1], oz = mz - c.campos[2];
        const float len = sqrtf(ox * ox + oy * oy + oz * oz);
        const float x = ox / len, y = oy / len, z = oz / len;
        const float* sh = in.shs + (size_t)idx * M * 3;
        float* dsh = g.dL_dsh + (size_t)idx * M * 3;
        const unsigned cb = clamped[idx];
        const float dRGB[3] = {(cb & 1u) ? 0.f : dcol[0], (cb & 2u) ? 0.f : dcol[1], (cb & 4u) ? 0.f : dcol[2]};
        float dRx[3] = {0, 0, 0}, dRy[3] = {0, 0, 0}, dRz[3] = {0, 0, 0};
        const int D = c.sh_degree;
        int nco = 1;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) PB_STORE(dsh + ch, bSHC0 * dRGB[ch]);
        if (D > 0) {
            nco = 4;
            const float w1 = -bSHC1 * y, w2 = bSHC1 * z, w3 = -bSHC1 * x;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                PB_STORE(dsh + 3 + ch, w1 * dRGB[ch]); PB_STORE(dsh + 6 + ch, w2 * dRGB[ch]); PB_STORE(dsh + 9 + ch, w3 * dRGB[ch]);
                dRx[ch] = -bSHC1 * sh[9 + ch];
                dRy[ch] = -bSHC1 * sh[3 + ch];
                dRz[ch] = bSHC1 * sh[6 + ch];
            }
            if (D > 1) {
                nco = 9;
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                const float w4 = bSHC2[0] * xy, w5 = bSHC2[1] * yz, w6 = bSHC2[2] * (2.f * zz - xx - yy);
                const float w7 = bSHC2[3] * xz, w8 = bSHC2[4] * (xx - yy);
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    PB_STORE(dsh + 12 + ch, w4 * dRGB[ch]); PB_STORE(dsh + 15 + ch, w5 * dRGB[ch]); PB_STORE(dsh + 18 + ch, w6 * dRGB[ch]);
                    PB_STORE(dsh + 21 + ch, w7 * dRGB[ch]); PB_STORE(dsh + 24 + ch, w8 * dRGB[ch]);
                    const float* s = sh + ch;
                    dRx[ch] += bSHC2[0] * y * s[12] + bSHC2[2] * 2.f * -x * s[18] + bSHC2[3] * z * s[21] + bSHC2[4] * 2.f * x * s[24];
                    dRy[ch] += bSHC2[0] * x * s[12] + bSHC2[1] * z * s[15] + bSHC2[2] * 2.f * -y * s[18] + bSHC2[4] * 2.f * -y * s[24];
                    dRz[ch] += bSHC2[1] * y * s[15] + bSHC2[2] * 2.f * 2.f * z * s[18] + bSHC2[3] * x * s[21];
                }
                if (D > 2) {
                    nco = 16;
                    const float w9 = bSHC3[0] * y * (3.f * xx - yy), w10 = bSHC3[1] * xy * z;
                    const float w11 = bSHC3[2] * y * (4.f * zz - xx - yy);
                    const float w12 = bSHC3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                    const float w13 = bSHC3[4] * x * (4.f * zz - xx - yy);
                    const float w14 = bSHC3[5] * z * (xx - yy), w15 = bSHC3[6] * x * (xx - 3.f * yy);
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        PB_STORE(dsh + 27 + ch, w9 * dRGB[ch]); PB_STORE(dsh + 30 + ch, w10 * dRGB[ch]); PB_STORE(dsh + 33 + ch, w11 * dRGB[ch]);
                        PB_STORE(dsh + 36 + ch, w12 * dRGB[ch]); PB_STORE(dsh + 39 + ch, w13 * dRGB[ch]); PB_STORE(dsh + 42 + ch, w14 * dRGB[ch]);
                        PB_STORE(dsh + 45 + ch, w15 * dRGB[ch]);
                        const float* s = sh + ch;
                        dRx[ch] += (bSHC3[0] * s[27] * 3.f * 2.f * xy + bSHC3[1] * s[30] * yz +
                                    bSHC3[2] * s[33] * -2.f * xy + bSHC3[3] * s[36] * -3.f * 2.f * xz +
                                    bSHC3[4] * s[39] * (-3.f * xx + 4.f * zz - yy) +
                                    bSHC3[5] * s[42] * 2.f * xz + bSHC3[6] * s[45] * 3.f * (xx - yy));
                        dRy[ch] += (bSHC3[0] * s[27] * 3.f * (xx - yy) + bSHC3[1] * s[30] * xz +
                                    bSHC3[2] * s[33] * (-3.f * yy + 4.f * zz - xx) +
                                    bSHC3[3] * s[36] * -3.f * 2.f * yz + bSHC3[4] * s[39] * -2.f * xy +
                                    bSHC3[5] * s[42] * -2.f * yz + bSHC3[6] * s[45] * -3.f * 2.f * xy);
                        dRz[ch] += (bSHC3[1] * s[30] * xy + bSHC3[2] * s[33] * 4.f * 2.f * yz +
                                    bSHC3[3] * s[36] * 3.f * (2.f * zz - xx - yy) +
                                    bSHC3[4] * s[39] * 4.f * 2.f * xz + bSHC3[5] * s[42] * (xx - yy));
                    }
                }
            }
        }
        if (!ACC) for (int i = 3 * nco; i < 3 * M; i++) dsh[i] = 0.f;
        const float ddx = dRx[0] * dRGB[0] + dRx[1] * dRGB[1] + dRx[2] * dRGB[2];
        const float ddy = dRy[0] * dRGB[0] + dRy[1] * dRGB[1] + dRy[2] * dRGB[2];
        const float ddz = dRz[0] * dRGB[0] + dRz[1] * dRGB[1] + dRz[2] * dRGB[2];
        const float sum2 = ox * ox + oy * oy + oz * oz;
        const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
        dm[0] += ((+sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * inv32;
        dm[1] += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * inv32;
        dm[2] += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * inv32;
    } else if (!ACC) {
        for (int i = 0; i < 3 * M; i++) g.dL_dsh[(size_t)idx * 3 * M + i] = 0.f;
    }
    if (ACC) {
        if (pg.d_means3D) {
#pragma unroll
            for (int i = 0; i < 3; i++) atomicAdd(pg.d_means3D + 3 * idx + i, dm[i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) g.dL_dmeans3D[3 * idx + i] = dm[i];
    }

    if (in.scales) {
        // computeCov3D backward (backward.cu:326-432)
        bool surface = true;
        if (!RGSS) surface = c.n_config > 0 && c.config[0] > 0;
        const float4 q = reinterpret_cast<const float4*>(in.rotations)[idx];
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        const float Rm[3][3] = {
            {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
            {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
            {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        const float s[3] = {c.scale_modifier * in.scales[3 * idx], c.scale_modifier * in.scales[3 * idx + 1],
                            c.scale_modifier * in.scales[3 * idx + 2]};
        const float dS[3][3] = {{dcv[0], 0.5f * dcv[1], 0.5f * dcv[2]},
                                {0.5f * dcv[1], dcv[3], 0.5f * dcv[4]},
                                {0.5f * dcv[2], 0.5f * dcv[4], dcv[5]}};
        float dM[3][3];  // dL_dM[c][r] = 2 * sum_k M[k][r] * dS[c][k], M[k][r] = s_r R[k][r]
#pragma unroll
        for (int cc2 = 0; cc2 < 3; cc2++)
#pragma unroll
            for (int rr = 0; rr < 3; rr++) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 3; k++) acc += (2.0f * (s[rr] * Rm[k][rr])) * dS[cc2][k];
                dM[cc2][rr] = acc;
            }
        float ds[3];
#pragma unroll
        for (int cc2 = 0; cc2 < 3; cc2++) ds[cc2] = Rm[0][cc2] * dM[0][cc2] + Rm[1][cc2] * dM[1][cc2] + Rm[2][cc2] * dM[2][cc2];
        PB_STORE(g.dL_dscales + 3 * idx, ds[0]);
        PB_STORE(g.dL_dscales + 3 * idx + 1, ds[1]);
        PB_STORE(g.dL_dscales + 3 * idx + 2, surface ? 0.f : ds[2]);
        float dRt[3][3];
#pragma unroll
        for (int cc2 = 0; cc2 < 3; cc2++)
#pragma unroll
            for (int rr = 0; rr < 3; rr++) dRt[cc2][rr] = dM[rr][cc2] * s[cc2];
        dRt[2][0] += dnrm[0] * V[0] + dnrm[1] * V[1] + dnrm[2] * V[2];
        dRt[2][1] += dnrm[0] * V[4] + dnrm[1] * V[5] + dnrm[2] * V[6];
        dRt[2][2] += dnrm[0] * V[8] + dnrm[1] * V[9] + dnrm[2] * V[10];
        float4 dq;
        dq.x = 2 * z * (dRt[0][1] - dRt[1][0]) + 2 * y * (dRt[2][0] - dRt[0][2]) + 2 * x * (dRt[1][2] - dRt[2][1]);
        dq.y = 2 * y * (dRt[1][0] + dRt[0][1]) + 2 * z * (dRt[2][0] + dRt[0][2]) + 2 * r * (dRt[1][2] - dRt[2][1]) - 4 * x * (dRt[2][2] + dRt[1][1]);
        dq.z = 2 * x * (dRt[1][0] + dRt[0][1]) + 2 * r * (dRt[2][0] - dRt[0][2]) + 2 * z * (dRt[1][2] + dRt[2][1]) - 4 * y * (dRt[2][2] + dRt[0][0]);
        dq.w = 2 * r * (dRt[0][1] - dRt[1][0]) + 2 * x * (dRt[2][0] + dRt[0][2]) + 2 * y * (dRt[1][2] + dRt[2][1]) - 4 * z * (dRt[1][1] + dRt[0][0]);
        if (ACC) {
            float4* dst = reinterpret_cast<float4*>(g.dL_drotations) + idx;
            float4 o = *dst;
            o.x += dq.x; o.y += dq.y; o.z += dq.z; o.w += dq.w;
            *dst = o;
        } else {
            reinterpret_cast<float4*>(g.dL_drotations)[idx] = dq;
        }
    } else if (!ACC) {
#pragma unroll
        for (int i = 0; i < 3; i++) g.dL_dscales[3 * idx + i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) g.dL_drotations[4 * idx + i] = 0.f;
    }
}

int launch_preprocess_bwd(const svgir_raster_cfg& c, const svgir_raster_in& in,
                          const svgir_raster_state& st, const int32_t* radii, svgir_raster_grads& g,
                          cudaStream_t s) {
    const int grid = (c.P + 255) / 256;
    const svgir_param_grads none = {};
    if (c.variant == SVGIR_VARIANT_RGSS)
        { TimedScope ts_("preprocess_bwd", s); preprocess_bwd_kernel<true, false><<<grid, 256, 0, s>>>(c, in, st.cov3D, st.clamped, radii, g.geo_grad, g, none, nullptr, nullptr, nullptr); }
    else
        { TimedScope ts_("preprocess_bwd", s); preprocess_bwd_kernel<false, false><<<grid, 256, 0, s>>>(c, in, st.cov3D, st.clamped, radii, g.geo_grad, g, none, nullptr, nullptr, nullptr); }
    return check_launch("preprocess_bwd", c.debug, s);
}

// svgir_raster_backward_params: visible surfels only, gradients added into the caller's buffers. The grid covers P
// (the list length is only known on the device); threads beyond the list return at once.
int launch_preprocess_bwd_params(const svgir_raster_cfg& c, const svgir_raster_in& in, const svgir_raster_state& st,
                                 const float* geo_grad, const svgir_param_grads& pg, cudaStream_t s) {
    const int grid = (c.P + 255) / 256;
    svgir_raster_grads g = {};
    if (c.variant == SVGIR_VARIANT_RGSS)
        { TimedScope ts_("preprocess_bwd", s); preprocess_bwd_kernel<true, true><<<grid, 256, 0, s>>>(c, in, st.cov3D, st.clamped, nullptr, geo_grad, g, pg, st.vis_list, st.vis_count, st.num_rendered); }
    else
        { TimedScope ts_("preprocess_bwd", s); preprocess_bwd_kernel<false, true><<<grid, 256, 0, s>>>(c, in, st.cov3D, st.clamped, nullptr, geo_grad, g, pg, st.vis_list, st.vis_count, st.num_rendered); }
    return check_launch("preprocess_bwd", c.debug, s);
}

}  // namespace svgir
