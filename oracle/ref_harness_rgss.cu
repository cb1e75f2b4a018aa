// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI harness around the *unmodified* reference stage-1 rasteriser
// (CudaRasterizer::Rasterizer::{forward,backward},
//  /root/reference/rgss-rasterization/cuda_rasterizer/rasterizer.h:24-103).
// The reference sources are compiled where they lie (see oracle/Makefile); this file only
// supplies what the reference's torch glue (rasterize_points.cu:27-145,147-265) supplies:
// growable state buffers and zero-initialised outputs -- here with cudaMalloc instead of
// torch tensors, so the resulting oracle/_ref/librgss_ref.so has no torch dependency and
// can be driven from ctypes with raw device pointers.
//
// It additionally exposes the reference's internal Geometry/Binning/Image state arrays
// (rasterizer_impl.h:32-75) so tests can compare sort keys, ranges, radii, n_contrib
// bit-for-bit.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <cuda_runtime.h>
#include "rasterizer_impl.h"   // reference header (found through -I, not copied)

namespace {
struct Buf {
    char* p = nullptr;
    size_t cap = 0;
    char* get(size_t n) {
        if (n > cap) {
            if (p) cudaFree(p);
            cudaMalloc(&p, n ? n : 1);
            cap = n;
        }
        return p;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct RefState {
    Buf geom, binning, img;
    int P = 0, W = 0, H = 0, R = 0;
};
}  // namespace

extern "C" {

void* ref_rgss_create() { return new RefState(); }

void ref_rgss_destroy(void* h) {
    RefState* s = (RefState*)h;
    s->geom.release(); s->binning.release(); s->img.release();
    delete s;
}

// Returns num_rendered (>=0) or -1. Outputs zero-filled by the caller as
// rgss-rasterization/rasterize_points.cu:77-92 does.
int ref_rgss_forward(void* h,
    int P, int S, int D, int M,
    const float* background, int W, int H,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* features, const float* opacities,
    const float* scales, float scale_modifier, const float* rotations,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, float cx, float cy, int prefiltered, int computer_pseudo_normal,
    float* out_color, float* out_normal, float* out_opacity, float* out_depth, float* out_feature,
    float* out_pseudo_normal, float* out_surface_xyz, float* out_weights, int* radii, int debug)
{
    RefState* s = (RefState*)h;
    s->P = P; s->W = W; s->H = H;
    try {
        std::function<char*(size_t)> g = [s](size_t n) { return s->geom.get(n); };
        std::function<char*(size_t)> b = [s](size_t n) { return s->binning.get(n); };
        std::function<char*(size_t)> i = [s](size_t n) { return s->img.get(n); };
        s->R = CudaRasterizer::Rasterizer::forward(g, b, i, P, S, D, M, background, W, H,
            means3D, shs, colors_precomp, features, opacities, scales, scale_modifier,
            rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, cx, cy,
            prefiltered != 0, computer_pseudo_normal != 0, out_color, out_normal, out_opacity, out_depth,
            out_feature, out_pseudo_normal, out_surface_xyz, out_weights, radii, debug != 0);
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_rgss_forward] %s\n", e.what());
        return -1;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    return s->R;
}

int ref_rgss_backward(void* h,
    int P, int S, int D, int M,
    const float* background, int W, int H,
    const float* means3D, const float* shs, const float* features,
    const float* colors_precomp, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
    const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy, const int* radii,
    const float* dL_dpix, const float* dL_dpix_n, const float* dL_dpix_o, const float* dL_dpix_d,
    const float* dL_dpix_f,
    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dnormal,
    float* dL_ddepth, float* dL_dfeature, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
    float* dL_dscale, float* dL_drot, int backward_geometry, int debug)
{
    RefState* s = (RefState*)h;
    try {
        CudaRasterizer::Rasterizer::backward(P, S, D, M, s->R, background, W, H, means3D,
            shs, features, colors_precomp, scales, scale_modifier, rotations,
            cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii,
            s->geom.p, s->binning.p, s->img.p, dL_dpix, dL_dpix_n, dL_dpix_o, dL_dpix_d, dL_dpix_f,
            dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dnormal, dL_ddepth, dL_dfeature,
            dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, backward_geometry != 0, debug != 0);
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_rgss_backward] %s\n", e.what());
        return -1;
    }
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

// Device pointer to one of the reference's internal state arrays, or nullptr.
// geometry: depths means2D conic_opacity rgb normal Jinv lambda tiles_touched point_offsets
//           cov3D clamped
// binning : keys_unsorted keys point_list_unsorted point_list
// image   : ranges n_contrib final_T final_D
void* ref_rgss_state(void* h, const char* name) {
    RefState* s = (RefState*)h;
    using namespace CudaRasterizer;
    char* gp = s->geom.p; char* bp = s->binning.p; char* ip = s->img.p;
    if (!gp || !ip) return nullptr;
    GeometryState g = GeometryState::fromChunk(gp, s->P);
    ImageState im = ImageState::fromChunk(ip, (size_t)s->W * s->H);
    if (!strcmp(name, "depths")) return g.depths;
    if (!strcmp(name, "means2D")) return g.means2D;
    if (!strcmp(name, "conic_opacity")) return g.conic_opacity;
    if (!strcmp(name, "rgb")) return g.rgb;
    if (!strcmp(name, "normal")) return g.normal;
    if (!strcmp(name, "Jinv")) return g.Jinv;
    if (!strcmp(name, "tiles_touched")) return g.tiles_touched;
    if (!strcmp(name, "point_offsets")) return g.point_offsets;
    if (!strcmp(name, "cov3D")) return g.cov3D;
    if (!strcmp(name, "clamped")) return g.clamped;
    if (!strcmp(name, "ranges")) return im.ranges;
    if (!strcmp(name, "n_contrib")) return im.n_contrib;
    if (!strcmp(name, "final_T")) return im.accum_alpha;
    if (!strcmp(name, "final_D")) return im.accum_depth;
    if (!bp) return nullptr;
    BinningState b = BinningState::fromChunk(bp, s->R);
    if (!strcmp(name, "keys_unsorted")) return b.point_list_keys_unsorted;
    if (!strcmp(name, "keys")) return b.point_list_keys;
    if (!strcmp(name, "point_list_unsorted")) return b.point_list_unsorted;
    if (!strcmp(name, "point_list")) return b.point_list;
    return nullptr;
}

int ref_rgss_num_rendered(void* h) { return ((RefState*)h)->R; }

}  // extern "C"
