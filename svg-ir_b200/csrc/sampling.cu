// Incident-ray sampling: Fibonacci hemisphere directions rotated from +z onto each surfel normal.
// Replaces the torch code of fibonacci_sphere_sampling (utils/graphics_utils.py:9-37) +
// rotation_between_z (utils/sh_utils.py:36-68), called through sample_incident_rays
// (scene/gaussian_model.py:23-31). One thread per (surfel, sample); consecutive threads write
// consecutive 12-byte directions, so the [N,Ns,3] store is fully coalesced and nothing but the
// normals ([N,3]) is read: 12 B in, 16 B out per ray instead of the reference's ~10 elementwise
// passes plus a batched 3x3 matmul over [N,3,Ns].
#include "common.cuh"

namespace svgir {

// Every elementwise step uses the _rn intrinsics so the roundings are those of the reference's
// separate torch kernels (no cross-op FMA contraction); only the 3x3 product is an fma chain.
__global__ void __launch_bounds__(256) sample_incident_rays_kernel(long long total, int Ns, const float* __restrict__ normals,
                                                                   const float* __restrict__ rand_u, float* __restrict__ dirs,
                                                                   float* __restrict__ areas) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int n = (int)(i / Ns), s = (int)(i - (long long)n * Ns);
    const float delta = 2.3999631f;        // float(pi * (3 - sqrt(5)))
    const float zmin = 0.17364818f;        // float(sin(10 deg))
    const float pi_f = 3.14159274f;
    float idx = (float)s;
    float z = fmaxf(sub_(1.0f, div_(mul_(2.0f, idx), (float)(2 * Ns - 1))), zmin);
    float rad = sqrt_(sub_(1.0f, mul_(z, z)));
    float theta = mul_(delta, idx);
    if (rand_u) theta = add_(mul_(mul_(rand_u[n], 2.0f), pi_f), theta);
    float y = mul_(cosf(theta), rad);
    float x = mul_(sinf(theta), rad);
    float nx = normals[3 * n], ny = normals[3 * n + 1], nz = normals[3 * n + 2];
    float v1 = -ny, v2 = nx;
    float c = fmaxf(add_(nz, 1.0f), 1e-7f);
    float r00, r01, r02, r10, r11, r12, r20, r21, r22;
    if (add_(nz, 1.0f) > 0.0f) {
        float v12c = div_(mul_(v1, v2), c);
        r00 = add_(1.0f, div_(-mul_(v2, v2), c)); r01 = v12c; r02 = v2;
        r10 = v12c; r11 = add_(1.0f, div_(-mul_(v1, v1), c)); r12 = -v1;
        r20 = -v2; r21 = v1; r22 = add_(1.0f, div_(sub_(-mul_(v2, v2), mul_(v1, v1)), c));
    } else {  // normal == -z: the reference substitutes -I
        r00 = r11 = r22 = -1.0f;
        r01 = r02 = r10 = r12 = r20 = r21 = 0.0f;
    }
    float dx = fmaf(r02, z, fmaf(r01, y, mul_(r00, x)));
    float dy = fmaf(r12, z, fmaf(r11, y, mul_(r10, x)));
    float dz = fmaf(r22, z, fmaf(r21, y, mul_(r20, x)));
    float len = fmaxf(sqrt_(fmaf(dz, dz, fmaf(dy, dy, mul_(dx, dx)))), 1e-12f);  // F.normalize eps
    dirs[3 * i] = div_(dx, len);
    dirs[3 * i + 1] = div_(dy, len);
    dirs[3 * i + 2] = div_(dz, len);
    if (areas) areas[i] = 6.28318548f;  // ones * 2 * pi in fp32
}

}  // namespace svgir

extern "C" int svgir_sample_incident_rays(int N, int Ns, const float* normals, const float* rand_u, float* incident_dirs,
                                          float* incident_areas, void* stream) {
    using namespace svgir;
    if (N < 0 || Ns <= 0) { set_error("sample_incident_rays: bad N/Ns"); return SVGIR_ERR_INVALID; }
    if (N == 0) return SVGIR_OK;
    if (!normals || !incident_dirs) { set_error("sample_incident_rays: null pointer"); return SVGIR_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    long long total = (long long)N * Ns;
    long long blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) { set_error("sample_incident_rays: too many rays"); return SVGIR_ERR_INVALID; }
    {
        TimedScope ts_("sample_incident_rays", s);
        sample_incident_rays_kernel<<<(unsigned)blocks, 256, 0, s>>>(total, Ns, normals, rand_u, incident_dirs, incident_areas);
    }
    return check_launch("sample_incident_rays", false, s);
}
