// extern "C" entry points of libsvgir_b200.so (declared in include/svgir_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "common.cuh"

namespace svgir {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Launch check. With debug set it also synchronises the stream, like the reference's CHECK_CUDA
// (cuda_rasterizer/auxiliary.h:425-432).
int check_launch(const char* what, bool debug, cudaStream_t stream) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && debug) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        set_error("[svgir_b200] %s: %s", what, cudaGetErrorString(e));
        return SVGIR_ERR_CUDA;
    }
    return SVGIR_OK;
}

// ---- per-kernel timing registry -------------------------------------------------------------
struct TimingRec { const char* name; cudaEvent_t e0, e1; };
static bool g_timing = false;
static TimingRec g_recs[4096];
static int g_nrec = 0;
static int g_open = -1;
static long long g_launches = 0;

void timing_begin(const char* name, cudaStream_t s) {
    g_launches++;
    if (!g_timing || g_nrec >= 4096) { g_open = -1; return; }
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { g_open = -1; return; }  // no timing events inside a graph capture
    TimingRec& r = g_recs[g_nrec];
    r.name = name;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) { g_open = -1; return; }
    cudaEventRecord(r.e0, s);
    g_open = g_nrec++;
}
void timing_end(cudaStream_t s) {
    if (g_open >= 0) cudaEventRecord(g_recs[g_open].e1, s);
    g_open = -1;
}

static int validate(const svgir_raster_cfg* c, const svgir_raster_in* in, bool need_features = true) {
    if (!c || !in) { set_error("null cfg/in"); return SVGIR_ERR_INVALID; }
    if (c->P < 0 || c->W <= 0 || c->H <= 0) { set_error("bad sizes P=%d W=%d H=%d", c->P, c->W, c->H); return SVGIR_ERR_INVALID; }
    if (c->S < 0 || c->S > SVGIR_MAX_S) { set_error("S=%d outside [0,%d]", c->S, SVGIR_MAX_S); return SVGIR_ERR_INVALID; }
    if (c->VS < 0 || (c->VS & 3) || c->VS / 4 > SVGIR_MAX_NV) { set_error("VS=%d must be a multiple of 4 and <= %d", c->VS, 4 * SVGIR_MAX_NV); return SVGIR_ERR_INVALID; }
    if (c->variant == SVGIR_VARIANT_RGSS && c->VS != 0) { set_error("rgss has no vfeatures"); return SVGIR_ERR_INVALID; }
    if (c->sh_degree > 3) { set_error("SH degree %d > 3", c->sh_degree); return SVGIR_ERR_INVALID; }
    if (!in->means3D || !in->opacities) { set_error("means3D/opacities missing"); return SVGIR_ERR_INVALID; }
    if ((in->shs != nullptr) == (in->colors_precomp != nullptr)) { set_error("provide exactly one of shs / colors_precomp"); return SVGIR_ERR_INVALID; }
    if (in->shs && c->M < (c->sh_degree + 1) * (c->sh_degree + 1)) { set_error("M=%d too small for SH degree %d", c->M, c->sh_degree); return SVGIR_ERR_INVALID; }
    if (!in->cov3D_precomp && (!in->scales || !in->rotations)) { set_error("provide scales+rotations or cov3D_precomp"); return SVGIR_ERR_INVALID; }
    if (!in->rotations) { set_error("rotations are required (surfel normals)"); return SVGIR_ERR_INVALID; }
    if (need_features && c->S > 0 && !in->features) { set_error("features missing"); return SVGIR_ERR_INVALID; }
    if (need_features && c->VS > 0 && !in->vfeatures) { set_error("vfeatures missing"); return SVGIR_ERR_INVALID; }
    if (((uintptr_t)in->rotations & 15) || (c->VS > 0 && ((uintptr_t)in->vfeatures & 15)) ||
        (c->S > 0 && (c->S & 3) == 0 && ((uintptr_t)in->features & 15))) {
        set_error("rotations/features/vfeatures must be 16-byte aligned");
        return SVGIR_ERR_INVALID;
    }
    if (!c->bg || !c->viewmatrix || !c->projmatrix || !c->campos) { set_error("camera constants missing"); return SVGIR_ERR_INVALID; }
    if (c->variant == SVGIR_VARIANT_SVGSS && (!c->patch_bbox || (c->n_config > 0 && !c->config))) { set_error("patch_bbox/config missing"); return SVGIR_ERR_INVALID; }
    return SVGIR_OK;
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D,
                                    const float* __restrict__ V, uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float x = means3D[3 * idx], y = means3D[3 * idx + 1], z = means3D[3 * idx + 2];
    const float pz = V[2] * x + V[6] * y + V[10] * z + V[14];
    present[idx] = pz > 0.2f;
}

// Work list of the surfels that survived culling (radii > 0): block-local ballot scan, ONE global atomic
// per block. Order across blocks is arbitrary -- consumers treat it as a set.
__global__ void __launch_bounds__(256) compact_visible_kernel(int P, const int32_t* __restrict__ radii,
                                                              int32_t* __restrict__ list, int32_t* __restrict__ count) {
    __shared__ int wsum[8];
    __shared__ int base_s;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool v = idx < P && radii[idx] > 0;
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) wsum[wid] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; w++) { const int c = wsum[w]; wsum[w] = t; t += c; }
        base_s = t ? atomicAdd(count, t) : 0;
    }
    __syncthreads();
    if (v) list[base_s + wsum[wid] + __popc(m & ((1u << lane) - 1u))] = idx;
}

}  // namespace svgir

using namespace svgir;

extern "C" {

const char* svgir_last_error(void) { return g_err; }
int svgir_version(void) { return 100; }

void svgir_timing_enable(int on) {
    g_timing = on != 0;
}

// Sums the elapsed device time (ms) and launch count of every recorded launch whose kernel name
// equals `name` (NULL = all), then forgets the records if `reset` is set. Synchronises the device.
int svgir_timing_collect(const char* name, double* total_ms, int* launches, int reset) {
    cudaDeviceSynchronize();
    double t = 0; int n = 0;
    for (int i = 0; i < g_nrec; i++) {
        if (name && strcmp(name, g_recs[i].name) != 0) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, g_recs[i].e0, g_recs[i].e1) == cudaSuccess) { t += ms; n++; }
    }
    if (total_ms) *total_ms = t;
    if (launches) *launches = n;
    if (reset) {
        for (int i = 0; i < g_nrec; i++) { cudaEventDestroy(g_recs[i].e0); cudaEventDestroy(g_recs[i].e1); }
        g_nrec = 0;
    }
    return SVGIR_OK;
}

// Number of kernel launches issued by this library since the last reset (the bench's gpu_launches).
long long svgir_launch_count(int reset) {
    long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

int svgir_raster_preprocess(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                            svgir_raster_state* st, svgir_raster_out* out, void* stream) {
    int rc = validate(cfg, in, false);  // the per-surfel preprocess does not read features / vfeatures
    if (rc) return rc;
    if (!st || !out || !st->rec || !st->tile_count || !st->num_rendered || !out->radii) {
        set_error("state/out buffers missing");
        return SVGIR_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (cfg->P == 0) return SVGIR_OK;
    rc = launch_preprocess(*cfg, *in, *st, *out, s);
    if (rc) return rc;
    if (st->vis_list) {
        if (!st->vis_count) { set_error("vis_list needs vis_count"); return SVGIR_ERR_INVALID; }
        if (cudaMemsetAsync(st->vis_count, 0, sizeof(int32_t), s) != cudaSuccess) { set_error("memset vis_count failed"); return SVGIR_ERR_CUDA; }
        { TimedScope ts_("compact_visible", s);
          compact_visible_kernel<<<(cfg->P + 255) / 256, 256, 0, s>>>(cfg->P, out->radii, st->vis_list, st->vis_count); }
        rc = check_launch("compact_visible", cfg->debug, s);
        if (rc) return rc;
    }
    return launch_tile_scan(*cfg, *st, s);
}

int svgir_raster_render(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                        svgir_raster_state* st, svgir_raster_out* out, void* stream) {
    int rc = validate(cfg, in);
    if (rc) return rc;
    if (!st || !out || !st->keys || !st->point_list || !st->final_T || !out->color) {
        set_error("state/out buffers missing");
        return SVGIR_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (cfg->P == 0) return SVGIR_OK;
    rc = launch_binning(*cfg, *st, out->radii, s);
    if (rc) return rc;
    rc = launch_composite_fwd(*cfg, *in, *st, *out, s);
    if (rc) return rc;
    if (cfg->variant == SVGIR_VARIANT_RGSS && cfg->computer_pseudo_normal) {
        if (!out->pseudo_normal || !out->surface_xyz) { set_error("pseudo_normal/surface_xyz buffers missing"); return SVGIR_ERR_INVALID; }
        rc = launch_pseudo_normal(*cfg, *out, s);
    }
    return rc;
}

int svgir_raster_bin(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                     svgir_raster_state* st, svgir_raster_out* out, void* stream) {
    int rc = validate(cfg, in, false);   // binning reads the geometry only
    if (rc) return rc;
    if (!st || !out || !st->keys || !st->point_list || !out->radii) { set_error("state/out buffers missing"); return SVGIR_ERR_INVALID; }
    if (cfg->P == 0) return SVGIR_OK;
    return launch_binning(*cfg, *st, out->radii, (cudaStream_t)stream);
}

int svgir_raster_composite(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                           svgir_raster_state* st, svgir_raster_out* out, void* stream) {
    int rc = validate(cfg, in);
    if (rc) return rc;
    if (!st || !out || !st->point_list || !st->final_T || !out->color) { set_error("state/out buffers missing"); return SVGIR_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    if (cfg->P == 0) return SVGIR_OK;
    rc = launch_composite_fwd(*cfg, *in, *st, *out, s);
    if (rc) return rc;
    if (cfg->variant == SVGIR_VARIANT_RGSS && cfg->computer_pseudo_normal) {
        if (!out->pseudo_normal || !out->surface_xyz) { set_error("pseudo_normal/surface_xyz buffers missing"); return SVGIR_ERR_INVALID; }
        rc = launch_pseudo_normal(*cfg, *out, s);
    }
    return rc;
}

int svgir_raster_backward_composite(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                                    const svgir_raster_state* st, svgir_raster_grads* g, void* stream) {
    int rc = validate(cfg, in);
    if (rc) return rc;
    if (!st || !g || !g->geo_grad || !g->dL_dcolor) { set_error("state/grad buffers missing"); return SVGIR_ERR_INVALID; }
    if (cfg->P == 0) return SVGIR_OK;
    return launch_composite_bwd(*cfg, *in, *st, *g, (cudaStream_t)stream);
}

int svgir_raster_backward_params(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                                 const svgir_raster_state* st, const float* geo_grad,
                                 const svgir_param_grads* pg, void* stream) {
    int rc = validate(cfg, in, false);
    if (rc) return rc;
    if (!st || !pg || !geo_grad || !st->vis_list || !st->vis_count || !st->num_rendered || !st->cov3D || !st->clamped) {
        set_error("backward_params: needs geo_grad, the visible-surfel list of the forward and the forward's state");
        return SVGIR_ERR_INVALID;
    }
    if ((in->shs && !pg->d_sh) || (in->scales && (!pg->d_scales || !pg->d_rotations)) || !pg->d_means3D || !pg->d_opacities ||
        ((uintptr_t)pg->d_rotations & 15)) {
        set_error("backward_params: d_means3D, d_opacities, d_sh (with shs), d_scales + d_rotations (with scales; d_rotations 16-byte aligned) are required");
        return SVGIR_ERR_INVALID;
    }
    if (cfg->P == 0) return SVGIR_OK;
    return launch_preprocess_bwd_params(*cfg, *in, *st, geo_grad, *pg, (cudaStream_t)stream);
}

int svgir_raster_backward(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                          const svgir_raster_state* st, const int32_t* radii,
                          svgir_raster_grads* g, void* stream) {
    int rc = validate(cfg, in);
    if (rc) return rc;
    if (!st || !g || !radii || !g->geo_grad || !g->dL_dcolor) {
        set_error("state/grad buffers missing");
        return SVGIR_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (cfg->P == 0) return SVGIR_OK;
    rc = launch_composite_bwd(*cfg, *in, *st, *g, s);
    if (rc) return rc;
    return launch_preprocess_bwd(*cfg, *in, *st, radii, *g, s);
}

int svgir_mark_visible(int variant, int P, const float* means3D, const float* viewmatrix,
                       const float* projmatrix, uint8_t* present, void* stream) {
    (void)projmatrix;
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return SVGIR_OK;
    if (variant == SVGIR_VARIANT_SVGSS) {
        // the stage-2 reference kernel body is commented out: all false (rasterizer_impl.cu:54-66)
        if (cudaMemsetAsync(present, 0, (size_t)P, s) != cudaSuccess) { set_error("memset failed"); return SVGIR_ERR_CUDA; }
        return SVGIR_OK;
    }
    { TimedScope ts_("mark_visible", s); mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, viewmatrix, present); }
    return check_launch("mark_visible", false, s);
}

}  // extern "C"
