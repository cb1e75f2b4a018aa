// Forward preprocess: project every surfel, cull, build the packed render record and count the
// tiles it touches.  One thread per surfel, grid = ceil(P/256).
//
// Replaces preprocessCUDA<3> of the reference (svgss_rasterization/cuda_rasterizer/forward.cu:230-396,
// rgss-rasterization/cuda_rasterizer/forward.cu:177-318) and the tiles_touched half of the binning
// (rasterizer_impl.cu:307).  The floating-point chain that decides culling, radius, tile rect and
// the depth key is written with explicit single-rounding intrinsics in the contraction pattern of
// the reference build, because sort keys and tile ranges must be bit-exact.
//
// Roofline: HBM-bound, algorithmic bytes = P*(44 + 12*M) read + P_vis*(96+24+8+...) written.
#include "common.cuh"

namespace svgir {

__constant__ float kSHC0 = 0.28209479177387814f;
__constant__ float kSHC1 = 0.4886025119029199f;
__constant__ float kSHC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
__constant__ float kSHC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

__device__ __forceinline__ void quat_to_R(const float4 q, float R[3][3]) {
    // forward.cu:165-180; R[c][r] column-major as GLM. Roundings as in the reference SASS.
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const float yy = mul_(y, y), zz = mul_(z, z);
    const float rz = mul_(r, z), xz = mul_(x, z), rx = mul_(r, x);
    float t;
    t = add_(yy, zz);      R[0][0] = sub_(1.f, add_(t, t));
    t = fma_(x, y, -rz);   R[0][1] = add_(t, t);
    t = fma_(r, y, xz);    R[0][2] = add_(t, t);
    t = fma_(x, y, rz);    R[1][0] = add_(t, t);
    t = fma_(x, x, zz);    R[1][1] = sub_(1.f, add_(t, t));
    t = fma_(y, z, -rx);   R[1][2] = add_(t, t);
    t = fma_(-r, y, xz);   R[2][0] = add_(t, t);
    t = fma_(y, z, rx);    R[2][1] = add_(t, t);
    t = fma_(x, x, yy);    R[2][2] = sub_(1.f, add_(t, t));
}

__device__ __forceinline__ void sh_to_rgb(int deg, const float* __restrict__ sh, float3 pos, float3 cam,
                                          float rgb[3], unsigned& clamp_bits) {
    // forward.cu:20-71
    float dx = pos.x - cam.x, dy = pos.y - cam.y, dz = pos.z - cam.z;
    float len = sqrtf(dx * dx + dy * dy + dz * dz);
    float x = dx / len, y = dy / len, z = dz / len;
    clamp_bits = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float res = kSHC0 * sh[c];
        if (deg > 0) {
            res = res - kSHC1 * y * sh[3 + c] + kSHC1 * z * sh[6 + c] - kSHC1 * x * sh[9 + c];
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                res = res + kSHC2[0] * xy * sh[12 + c] + kSHC2[1] * yz * sh[15 + c] +
                      kSHC2[2] * (2.0f * zz - xx - yy) * sh[18 + c] + kSHC2[3] * xz * sh[21 + c] +
                      kSHC2[4] * (xx - yy) * sh[24 + c];
                if (deg > 2) {
                    res = res + kSHC3[0] * y * (3.0f * xx - yy) * sh[27 + c] +
                          kSHC3[1] * xy * z * sh[30 + c] +
                          kSHC3[2] * y * (4.0f * zz - xx - yy) * sh[33 + c] +
                          kSHC3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
                          kSHC3[4] * x * (4.0f * zz - xx - yy) * sh[39 + c] +
                          kSHC3[5] * z * (xx - yy) * sh[42 + c] +
                          kSHC3[6] * x * (xx - 3.0f * yy) * sh[45 + c];
                }
            }
        }
        res += 0.5f;
        if (res < 0) clamp_bits |= 1u << c;
        rgb[c] = fmaxf(res, 0.0f);
    }
}

template <bool RGSS>
__global__ void __launch_bounds__(256) preprocess_kernel(
    const svgir_raster_cfg c, const svgir_raster_in in, float4* __restrict__ rec,
    float* __restrict__ cov3D_out, uint8_t* __restrict__ clamped, ushort4* __restrict__ rect_out,
    uint32_t* __restrict__ tiles_touched, uint32_t* __restrict__ tile_count,
    int32_t* __restrict__ radii) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= c.P) return;
    radii[idx] = 0;
    tiles_touched[idx] = 0;

    const float* __restrict__ V = c.viewmatrix;
    const float* __restrict__ PV = c.projmatrix;
    const int W = c.W, H = c.H;
    const float fy = H / (2.0f * c.tan_fovy), fx = W / (2.0f * c.tan_fovx);
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;

    const float x = in.means3D[3 * idx], y = in.means3D[3 * idx + 1], z = in.means3D[3 * idx + 2];
    const float hx = add_(dot3_col(PV, 0, x, y, z), PV[12]);
    const float hy = add_(dot3_col(PV, 1, x, y, z), PV[13]);
    const float hw = add_(dot3_col(PV, 3, x, y, z), PV[15]);
    const float p_w = __frcp_rn(add_(hw, 0.0000001f));
    const float projx = mul_(hx, p_w), projy = mul_(hy, p_w);
    const float pvx = add_(dot3_col(V, 0, x, y, z), V[12]);
    const float pvy = add_(dot3_col(V, 1, x, y, z), V[13]);
    const float pvz = add_(dot3_col(V, 2, x, y, z), V[14]);
    // ndc2Pix is evaluated in fp64 in the reference (auxiliary.h:42-46)
    const float pix_x = __double2float_rn(__dmul_rn(__fma_rn(__dadd_rn((double)projx, 1.0), (double)W, -1.0), 0.5));
    const float pix_y = __double2float_rn(__dmul_rn(__fma_rn(__dadd_rn((double)projy, 1.0), (double)H, -1.0), 0.5));

    bool surface = true, pix_depth = true;
    if (!RGSS) {
        surface = c.n_config > 0 && c.config[0] > 0;
        pix_depth = c.n_config > 2 && c.config[2] > 0;
        const float x0 = c.patch_bbox[1], y0 = c.patch_bbox[0], x1 = c.patch_bbox[3], y1 = c.patch_bbox[2];
        const float w = sub_(x1, x0), h = sub_(y1, y0);
        if (pvz < 0 || pix_x < fma_(w, -0.2f, x0) || pix_x >= fma_(w, 0.2f, x1) ||
            pix_y < fma_(h, -0.2f, y0) || pix_y >= fma_(h, 0.2f, y1))
            return;
    } else {
        if (pvz <= 0.2f) return;
    }

    float R[3][3];
    float4 q = make_float4(1, 0, 0, 0);
    if (in.rotations) q = reinterpret_cast<const float4*>(in.rotations)[idx];
    quat_to_R(q, R);
    float nv[3] = {0, 0, 0};
    float J[4] = {0, 0, 0, 0}, J6 = 0, J9 = 0;
    if (surface) {
        float ax0[3], ax1[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            nv[i] = dot3_col(V, i, R[0][2], R[1][2], R[2][2]);
            ax0[i] = dot3_col(V, i, R[0][0], R[1][0], R[2][0]);
            ax1[i] = dot3_col(V, i, R[0][1], R[1][1], R[2][1]);
        }
        const float dotpn = fma_(pvz, nv[2], fma_(pvx, nv[0], mul_(pvy, nv[1])));
        if ((double)dotpn > -0.01) return;  // back-facing (auxiliary.h:173-208)
        if (pix_depth) {
            // local_homo (auxiliary.h:291-388)
            const float prjx = div_(pvx, pvz), prjy = div_(pvy, pvz);
            const float a0 = add_(prjx, 0.001f);
            const float mod0 = fmaxf(sqrt_(add_(fma_(prjy, prjy, mul_(a0, a0)), 1.0f)), 0.00000001f);
            const float d00 = div_(a0, mod0), d01 = div_(prjy, mod0), d02 = __frcp_rn(mod0);
            const float b1 = add_(prjy, 0.001f);
            const float mod1 = fmaxf(sqrt_(add_(fma_(prjx, prjx, mul_(b1, b1)), 1.0f)), 0.00000001f);
            const float d10 = div_(prjx, mod1), d11 = div_(b1, mod1), d12 = __frcp_rn(mod1);
            const float prj_x0 = fma_(nv[2], d02, fma_(nv[0], d00, mul_(nv[1], d01)));
            const float prj_x1 = fma_(nv[2], d12, fma_(nv[0], d10, mul_(nv[1], d11)));
            if (fabsf(div_(prj_x0, mod0)) < 0.01f) return;  // grazing
            if (fabsf(div_(prj_x1, mod1)) < 0.01f) return;
            const float t_x0 = div_(dotpn, prj_x0), t_x1 = div_(dotpn, prj_x1);
            const float xu0[3] = {fma_(t_x0, d00, -pvx), fma_(t_x0, d01, -pvy), fma_(t_x0, d02, -pvz)};
            const float xu1[3] = {fma_(t_x1, d10, -pvx), fma_(t_x1, d11, -pvy), fma_(t_x1, d12, -pvz)};
            const float sc = div_(mul_(add_(fx, fy), 0.5f), 1000.0f);
            J[0] = div_(fma_(ax0[2], xu0[2], fma_(ax0[0], xu0[0], mul_(ax0[1], xu0[1]))), sc);
            J[1] = div_(fma_(ax0[2], xu1[2], fma_(ax0[0], xu1[0], mul_(ax0[1], xu1[1]))), sc);
            J[2] = div_(fma_(ax1[2], xu0[2], fma_(ax1[0], xu0[0], mul_(ax1[1], xu0[1]))), sc);
            J[3] = div_(fma_(ax1[2], xu1[2], fma_(ax1[0], xu1[0], mul_(ax1[1], xu1[1]))), sc);
            J6 = ax0[2];
            J9 = ax1[2];
        }
    }

    // 3D covariance (forward.cu:186-226), including the `mod * surface ? 0 : scale.z` precedence.
    float cv[6];
    float3 sc3 = make_float3(0, 0, 0);
    if (in.scales) sc3 = make_float3(in.scales[3 * idx], in.scales[3 * idx + 1], in.scales[3 * idx + 2]);
    if (in.cov3D_precomp) {
#pragma unroll
        for (int i = 0; i < 6; i++) cv[i] = in.cov3D_precomp[6 * idx + i];
    } else {
        float s[3];
        s[0] = mul_(c.scale_modifier, sc3.x);
        s[1] = mul_(c.scale_modifier, sc3.y);
        s[2] = (mul_(c.scale_modifier, surface ? 1.0f : 0.0f) != 0.0f) ? 0.0f : sc3.z;
        float Mx[3][3];
#pragma unroll
        for (int cc = 0; cc < 3; cc++)
#pragma unroll
            for (int r = 0; r < 3; r++) Mx[cc][r] = mul_(s[r], R[cc][r]);
        cv[0] = dot3_glm(Mx[0][0], Mx[0][0], Mx[0][1], Mx[0][1], Mx[0][2], Mx[0][2]);
        cv[1] = dot3_glm(Mx[0][0], Mx[1][0], Mx[0][1], Mx[1][1], Mx[0][2], Mx[1][2]);
        cv[2] = dot3_glm(Mx[0][0], Mx[2][0], Mx[0][1], Mx[2][1], Mx[0][2], Mx[2][2]);
        cv[3] = dot3_glm(Mx[1][0], Mx[1][0], Mx[1][1], Mx[1][1], Mx[1][2], Mx[1][2]);
        cv[4] = dot3_glm(Mx[1][0], Mx[2][0], Mx[1][1], Mx[2][1], Mx[1][2], Mx[2][2]);
        cv[5] = dot3_glm(Mx[2][0], Mx[2][0], Mx[2][1], Mx[2][1], Mx[2][2], Mx[2][2]);
#pragma unroll
        for (int i = 0; i < 6; i++) cov3D_out[6 * idx + i] = cv[i];
    }

    // EWA 2D covariance (forward.cu:74-139) on the view-space mean
    float a, b, cc2;
    {
        const float limx = mul_(1.3f, c.tan_fovx), limy = mul_(1.3f, c.tan_fovy);
        const float txtz = div_(pvx, pvz), tytz = div_(pvy, pvz);
        const float cxx = fminf(limx, fmaxf(-limx, txtz));
        const float cyy = fminf(limy, fmaxf(-limy, tytz));
        const float tz2 = mul_(pvz, pvz);
        const float j00 = div_(fx, pvz);
        const float j02 = div_(mul_(mul_(pvz, -cxx), fx), tz2);
        const float j11 = div_(fy, pvz);
        const float j12 = div_(mul_(mul_(pvz, -cyy), fy), tz2);
        float T0[3], T1[3];
        T0[0] = fma_(V[2], j02, mul_(V[0], j00));
        T0[1] = fma_(V[6], j02, mul_(V[4], j00));
        T0[2] = fma_(V[10], j02, mul_(V[8], j00));
        T1[0] = fma_(V[2], j12, mul_(V[1], j11));
        T1[1] = fma_(V[6], j12, mul_(V[5], j11));
        T1[2] = fma_(V[10], j12, mul_(V[9], j11));
        const float A00 = dot3_glm(T0[0], cv[0], T0[1], cv[1], T0[2], cv[2]);
        const float A01 = dot3_glm(T1[0], cv[0], T1[1], cv[1], T1[2], cv[2]);
        const float A10 = dot3_glm(T0[0], cv[1], T0[1], cv[3], T0[2], cv[4]);
        const float A11 = dot3_glm(T1[0], cv[1], T1[1], cv[3], T1[2], cv[4]);
        const float A20 = dot3_glm(T0[0], cv[2], T0[1], cv[4], T0[2], cv[5]);
        const float A21 = dot3_glm(T1[0], cv[2], T1[1], cv[4], T1[2], cv[5]);
        a = add_(dot3_glm(T0[0], A00, T0[1], A10, T0[2], A20), 0.3f);
        b = dot3_glm(T0[0], A01, T0[1], A11, T0[2], A21);
        cc2 = add_(dot3_glm(T1[0], A01, T1[1], A11, T1[2], A21), 0.3f);
    }
    const float det = fma_(a, cc2, -mul_(b, b));
    if (det == 0.0f) return;
    const float det_inv = __frcp_rn(det);
    const float conx = mul_(cc2, det_inv), cony = mul_(b, -det_inv), conz = mul_(a, det_inv);
    const float mid = mul_(add_(a, cc2), 0.5f);
    const float sq = sqrt_(fmaxf(0.1f, fma_(mid, mid, -det)));
    const float l1 = add_(mid, sq), l2 = sub_(mid, sq);
    const int my_radius = __float2int_ru(mul_(sqrt_(fmaxf(l1, l2)), 3.f));

    // getRect (auxiliary.h:53-63)
    const float fr = (float)my_radius;
    int rx0 = __float2int_rz(mul_(sub_(pix_x, fr), 0.0625f));
    int ry0 = __float2int_rz(mul_(sub_(pix_y, fr), 0.0625f));
    int rx1 = __float2int_rz(mul_(sub_(add_(add_(pix_x, fr), 16.0f), 1.0f), 0.0625f));
    int ry1 = __float2int_rz(mul_(sub_(add_(add_(pix_y, fr), 16.0f), 1.0f), 0.0625f));
    rx0 = min(gx, max(0, rx0)); ry0 = min(gy, max(0, ry0));
    rx1 = min(gx, max(0, rx1)); ry1 = min(gy, max(0, ry1));
    const int ntiles = (rx1 - rx0) * (ry1 - ry0);
    if (ntiles == 0) return;

    float rgb[3];
    if (in.colors_precomp) {
        rgb[0] = in.colors_precomp[3 * idx]; rgb[1] = in.colors_precomp[3 * idx + 1];
        rgb[2] = in.colors_precomp[3 * idx + 2];
    } else {
        unsigned bits;
        sh_to_rgb(c.sh_degree, in.shs + (size_t)idx * c.M * 3, make_float3(x, y, z),
                  make_float3(c.campos[0], c.campos[1], c.campos[2]), rgb, bits);
        clamped[idx] = (uint8_t)bits;
    }

    radii[idx] = my_radius;
    tiles_touched[idx] = (uint32_t)ntiles;
    rect_out[idx] = make_ushort4((unsigned short)rx0, (unsigned short)ry0, (unsigned short)rx1,
                                 (unsigned short)ry1);
    // uv_max = 0.5*lambda + 0.1 evaluated in double then rounded to float (forward.cu:608);
    // the renderer multiplies by su = 0.5/uv_max instead of dividing.
    const float lx = sc3.x, ly = sc3.y;  // lambda = raw scales.xy (forward.cu:394)
    const float umx = __double2float_rn(__fma_rn(0.5, (double)lx, 0.1));
    const float umy = __double2float_rn(__fma_rn(0.5, (double)ly, 0.1));
    float4* r4 = rec + (size_t)idx * REC_F4;
    r4[0] = make_float4(pix_x, pix_y, conx, cony);
    r4[1] = make_float4(conz, in.opacities[idx], pvz, 0.5f / umx);
    r4[2] = make_float4(J[0], J[1], J[2], J[3]);
    r4[3] = make_float4(J6, J9, 0.5f / umy, fr);
    r4[4] = make_float4(rgb[0], rgb[1], rgb[2], nv[0]);
    r4[5] = make_float4(nv[1], nv[2], lx, ly);

    // per-tile instance counts (the counting half of the tile-major radix pass)
    for (int ty = ry0; ty < ry1; ty++)
        for (int tx = rx0; tx < rx1; tx++) atomicAdd(&tile_count[ty * gx + tx], 1u);
}

int launch_preprocess(const svgir_raster_cfg& c, const svgir_raster_in& in, svgir_raster_state& st,
                      svgir_raster_out& out, cudaStream_t s) {
    const int T = ((c.W + TILE - 1) / TILE) * ((c.H + TILE - 1) / TILE);
    if (cudaMemsetAsync(st.tile_count, 0, sizeof(uint32_t) * T, s) != cudaSuccess) {
        set_error("memset tile_count failed");
        return SVGIR_ERR_CUDA;
    }
    const int grid = (c.P + 255) / 256;
    if (c.variant == SVGIR_VARIANT_RGSS)
        { TimedScope ts_("preprocess", s); preprocess_kernel<true><<<grid, 256, 0, s>>>(c, in, (float4*)st.rec, st.cov3D, st.clamped,
                                                      (ushort4*)st.rect, st.tiles_touched,
                                                      st.tile_count, out.radii); }
    else
        { TimedScope ts_("preprocess", s); preprocess_kernel<false><<<grid, 256, 0, s>>>(c, in, (float4*)st.rec, st.cov3D, st.clamped,
                                                       (ushort4*)st.rect, st.tiles_touched,
                                                       st.tile_count, out.radii); }
    return check_launch("preprocess", c.debug, s);
}

}  // namespace svgir
