/* svgir_b200 -- C ABI of the B200-native SVG-IR hot path (libsvgir_b200.so).
 *
 * Drop-in boundary: these entry points are what the reference's Python/C++ seam for the
 * splatting + shading path binds.  Each one cites the reference interface it replaces
 * (paths relative to learner-shx/SVG-IR @96dd9a5).  All pointers are DEVICE pointers unless a
 * parameter is documented as host; all tensors are contiguous fp32 unless noted; every call is
 * asynchronous on `stream` (a cudaStream_t passed as void*) and returns 0 on success or a
 * negative svgir_status (message via svgir_last_error()).  No torch types cross this boundary;
 * the caller (Python host code under svg-ir_b200/) owns every buffer.
 */
#ifndef SVGIR_B200_H_
#define SVGIR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum svgir_status {
    SVGIR_OK = 0,
    SVGIR_ERR_INVALID = -1,   /* bad argument (shape limits, null pointer, misalignment) */
    SVGIR_ERR_CUDA = -2,      /* a CUDA runtime call or kernel launch failed */
    SVGIR_ERR_CAPACITY = -3,  /* a caller-provided buffer is too small */
};

#define SVGIR_VARIANT_SVGSS 0 /* stage 2: svgss_rasterization (SV materials)          */
#define SVGIR_VARIANT_RGSS 1  /* stage 1: rgss-rasterization (flat features, config={1,1,1}) */

#define SVGIR_REC_FLOATS 24   /* packed per-surfel record, see svgir_raster_state.rec   */
#define SVGIR_GEO_GRAD_FLOATS 16
#define SVGIR_MAX_S 64        /* reference limits: S<=50, VS/4<=20 (forward.cu:483-486)   */
#define SVGIR_MAX_NV 32

const char* svgir_last_error(void);
int svgir_version(void);

/* Measurement hooks (no reference counterpart; the reference has only commented-out timers,
 * SURVEY.md section 5). With timing enabled every kernel launch of this library is bracketed by
 * CUDA events on the launching stream; svgir_timing_collect sums elapsed ms / launch count for one
 * kernel name (NULL = all) and synchronises the device. svgir_launch_count returns the number of
 * kernels this library has launched (bench.py's gpu_launches). */
void svgir_timing_enable(int on);
int svgir_timing_collect(const char* kernel_name, double* total_ms, int* launches, int reset);
long long svgir_launch_count(int reset);

/* Per-view constants. Mirrors GaussianRasterizationSettings
 * (gaussian_renderer/svgss_rasterization.py:331-346, rgss_rasterization.py:189-205) plus the
 * sizes RasterizeGaussiansCUDA derives (svgss_rasterization/rasterize_points.cu:69-73,102-106). */
typedef struct svgir_raster_cfg {
    int32_t P, S, VS, sh_degree, M;  /* surfels, flat features, SV floats (VS%4==0), SH degree, SH coeffs */
    int32_t W, H;
    int32_t variant;                 /* SVGIR_VARIANT_* */
    float tan_fovx, tan_fovy, scale_modifier;
    int32_t prefiltered, debug;      /* debug!=0: synchronise + check after every launch (auxiliary.h:425-432) */
    int32_t n_config;                /* #floats behind `config` (reference model passes 3; [3] treated as 0) */
    int32_t backward_geometry;       /* rgss only (rgss backward.cu:646-649) */
    int32_t computer_pseudo_normal;  /* rgss only */
    float cx, cy;                    /* rgss only */
    const float* bg;          /* [3]  */
    const float* viewmatrix;  /* [16] row-vector convention, read column-major (auxiliary.h:65-84) */
    const float* projmatrix;  /* [16] */
    const float* campos;      /* [3]  */
    const float* patch_bbox;  /* [4] (h0,w0,h1,w1); svgss only */
    const float* config;      /* [n_config] svgss only: [0]>0 surface [1]>0 normalize_depth [2]>0 per_pixel_depth */
} svgir_raster_cfg;

/* Per-surfel inputs (rasterize_points.cu:36-63). Absent optional = NULL. */
typedef struct svgir_raster_in {
    const float* means3D;        /* [P,3] */
    const float* opacities;      /* [P,1] */
    const float* scales;         /* [P,3] or NULL when cov3D_precomp */
    const float* rotations;      /* [P,4] */
    const float* cov3D_precomp;  /* [P,6] or NULL */
    const float* shs;            /* [P,M,3] or NULL */
    const float* colors_precomp; /* [P,3] or NULL */
    const float* features;       /* [P,S]  or NULL when S==0 */
    const float* vfeatures;      /* [P,VS] or NULL when VS==0 */
} svgir_raster_in;

/* State carried from forward to backward. Replaces the three opaque byte buffers
 * geomBuffer/binningBuffer/imgBuffer (rasterizer_impl.h:32-75) with typed arrays the caller
 * allocates. rec[P][24] = { mx,my,conic.x,conic.y | conic.z,opacity,depth,su | J0,J1,J2,J3 |
 * J6,J9,sv,radius | r,g,b,n.x | n.y,n.z,lambda.x,lambda.y } with su,sv = 0.5/(0.5*lambda+0.1). */
typedef struct svgir_raster_state {
    float* rec;               /* [P,24] written for visible surfels only */
    float* cov3D;             /* [P,6]  */
    uint8_t* clamped;         /* [P] bit c set when SH colour channel c was clamped */
    uint16_t* rect;           /* [P,4] tile rect (x0,y0,x1,y1) */
    uint32_t* tiles_touched;  /* [P] */
    uint32_t* tile_count;     /* [T] */
    uint32_t* tile_cursor;    /* [T] scratch */
    uint32_t* ranges;         /* [T,2] (start,end), (0,0) for empty tiles */
    uint32_t* big_tiles;      /* [3*T+4] scratch: work lists of the medium / large tile sorters, then the tiles'
                                 compositing order (heaviest first) at [2+2T, 2+3T) */
    int32_t* num_rendered;    /* [2] device: R, overflow flag */
    uint64_t* keys;           /* [cap_R] scratch, (depth_bits<<32)|surfel */
    uint32_t* point_list;     /* [cap_R] sorted surfel ids */
    uint64_t* sorted_keys;    /* [cap_R] optional debug output (tile<<32)|depth_bits, or NULL */
    int64_t cap_R;            /* capacity of keys / point_list */
    float* final_T;           /* [H*W] */
    float* final_D;           /* [H*W] */
    uint32_t* n_contrib;      /* [H*W] */
    int32_t* vis_list;        /* [P] optional: indices of the surfels with radii > 0 (unordered), or NULL */
    int32_t* vis_count;       /* [1] optional: length of vis_list (device) */
} svgir_raster_state;

/* Outputs of the forward pass; caller zero-fills out_weights, everything else is fully written.
 * (rasterize_points.cu:78-88,144; rgss: rgss-rasterization/rasterize_points.cu:77-141) */
typedef struct svgir_raster_out {
    float* color;    /* [3,H,W] */
    float* normal;   /* [3,H,W] */
    float* depth;    /* [1,H,W] */
    float* opacity;  /* [1,H,W] */
    float* feature;  /* [S,H,W] */
    float* vfeature; /* [VS/4,H,W] svgss */
    float* weights;  /* [P,1] accumulated blend weight (forward.cu:653) */
    int32_t* radii;  /* [P] */
    float* pseudo_normal; /* [3,H,W] rgss, zero-filled by caller */
    float* surface_xyz;   /* [3,H,W] rgss */
} svgir_raster_out;

/* Forward, part 1: per-surfel preprocess (forward.cu:230-396), per-tile counting and the tile
 * scan (replaces InclusiveSum + D2H, rasterizer_impl.cu:307-311). Leaves R in
 * state->num_rendered[0]; needs rec..num_rendered of `state`. Does not read in->features / in->vfeatures
 * (they may still be NULL: the shading that produces them can run between part 1 and part 2, restricted
 * to state->vis_list, which this call fills when the pointer is set). */
int svgir_raster_preprocess(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                            svgir_raster_state* state, svgir_raster_out* out, void* stream);

/* Forward, part 2: duplicateWithKeys + tile|depth radix sort + identifyTileRanges
 * (rasterizer_impl.cu:70-138,319-348) and the forward compositing kernel (forward.cu:402-750).
 * If R exceeds state->cap_R the overflow flag num_rendered[1] is set and nothing is rendered. */
int svgir_raster_render(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                        svgir_raster_state* state, svgir_raster_out* out, void* stream);

/* The two halves of svgir_raster_render as separate calls, so that a caller can run the binning on one stream while
 * the shading that produces in->features / in->vfeatures runs on another (binning needs the geometry only):
 * svgir_raster_bin = duplicateWithKeys + sort + ranges; svgir_raster_composite = the forward compositing kernel
 * (+ the rgss screen-space kernels). render == bin followed by composite. */
int svgir_raster_bin(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                     svgir_raster_state* state, svgir_raster_out* out, void* stream);
int svgir_raster_composite(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                           svgir_raster_state* state, svgir_raster_out* out, void* stream);

/* Pixel gradients in, per-surfel gradients out (RasterizeGaussiansBackwardCUDA,
 * rasterize_points.cu:147-265; Rasterizer::backward rasterizer_impl.cu:386-523). */
typedef struct svgir_raster_grads {
    const float* dL_dcolor;    /* [3,H,W] */
    const float* dL_dnormal;   /* [3,H,W] */
    const float* dL_ddepth;    /* [1,H,W] */
    const float* dL_dopacity;  /* [1,H,W] */
    const float* dL_dfeature;  /* [S,H,W]  or NULL */
    const float* dL_dvfeature; /* [VS/4,H,W] or NULL */
    /* accumulators, zero-filled by the caller */
    float* geo_grad;           /* [P,16]: mean2D.xy, conic.xyw, opacity, colour rgb, normal xyz, depth, pad */
    float* dL_dfeatures;       /* [P,S]  (final output) */
    float* dL_dvfeatures;      /* [P,VS] (final output) */
    /* outputs fully written by the backward-preprocess kernel */
    float* dL_dmeans2D;        /* [P,3] */
    float* dL_dcolors;         /* [P,3] */
    float* dL_dopacities;      /* [P,1] */
    float* dL_dmeans3D;        /* [P,3] */
    float* dL_dcov3D;          /* [P,6] */
    float* dL_dsh;             /* [P,M,3] */
    float* dL_dscales;         /* [P,3] */
    float* dL_drotations;      /* [P,4] */
    float* dL_dconic;          /* [P,4] optional debug copy or NULL */
    float* dL_dnormal3;        /* [P,3] optional debug copy or NULL */
    float* dL_ddepths;         /* [P]   optional debug copy or NULL */
} svgir_raster_grads;

int svgir_raster_backward(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                          const svgir_raster_state* state, const int32_t* radii,
                          svgir_raster_grads* g, void* stream);

/* The two halves of svgir_raster_backward for a caller that owns the step (no autograd in between):
 * svgir_raster_backward_composite runs the backward compositing kernel only (fills geo_grad / dL_dfeatures /
 * dL_dvfeatures, which the caller zero-filled); svgir_raster_backward_params turns geo_grad into PARAMETER
 * gradients for the surfels of state->vis_list only (culled surfels are skipped, not zero-filled) and ADDS them
 * into the caller's gradient buffers -- the .grad tensors of the optimiser, or one rank's slice of the flat
 * data-parallel bucket -- so that a step rendering several views accumulates without extra kernels. d_means3D is
 * updated with fp32 atomics (the shading backward adds the view-direction term into the same rows, possibly at
 * the same time on another stream). NULL pointers are skipped. Does nothing when the forward's bins overflowed
 * (state->num_rendered[1] != 0): an overflowed step contributes exactly zero and is re-run by its owner. */
typedef struct svgir_param_grads {
    float* d_means3D;    /* [P,3]   += (atomic) */
    float* d_opacities;  /* [P,1]   += */
    float* d_scales;     /* [P,3]   += */
    float* d_rotations;  /* [P,4]   += */
    float* d_sh;         /* [P,M,3] += */
    float* d_means2D;    /* [P,3]   += screen-space gradient (densification statistic, gaussian_model.py:1270-1276) */
} svgir_param_grads;
int svgir_raster_backward_composite(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                                    const svgir_raster_state* state, svgir_raster_grads* g, void* stream);
int svgir_raster_backward_params(const svgir_raster_cfg* cfg, const svgir_raster_in* in,
                                 const svgir_raster_state* state, const float* geo_grad,
                                 const svgir_param_grads* pg, void* stream);

/* mark_visible (rasterize_points.cu:267-286). svgss: all false (rasterizer_impl.cu:54-66);
 * rgss: frustum test z > 0.2 (rgss auxiliary.h:146-171). present is uint8/bool [P]. */
int svgir_mark_visible(int variant, int P, const float* means3D, const float* viewmatrix,
                       const float* projmatrix, uint8_t* present, void* stream);


/* ---------------------------------------------------------------------------------------------
 * PBR render_equation (shading).  Replaces the live torch path rendering_equation4 +
 * GGX_specular4 (gaussian_renderer/svgss.py:537-631) fused with the env-map lookup
 * DirectLightMap.direct_light / EnvLight.direct_light (scene/direct_light_map.py:70-106,
 * scene/envmap.py:54-72); the optional per-vertex metallic follows the legacy CUDA
 * render_equation (rgss-rasterization/render_equation.cu:55-190: f_d=(1-m)base/pi,
 * F0=0.04(1-m)+base*m).  All [N,12] tensors are channel-major (R v0..v3, G v0..v3, B v0..v3). */
#define SVGIR_SHADE_ENV_COPIES 16   /* replicas of the env-gradient accumulator (same-address reductions serialise) */
typedef struct svgir_shade_cfg {
    int32_t N, Ns;            /* surfels, light samples per surfel */
    int32_t env_h, env_w;     /* lat-long env map size */
    int32_t env_mode;         /* 0: learnable map -- softplus(param), result x2 (direct_light_map.py:83,106);
                                 1: fixed linear map as given (EnvLight after its 32x64 resize), x1 */
    int32_t debug;
    int32_t flags;            /* SVGIR_SHADE_* bits */
    int32_t reserved_;
} svgir_shade_cfg;
#define SVGIR_SHADE_ENV_READY 1   /* env_act_scratch already holds the activated map (a forward call of this step
                                     wrote it): skip the activation kernel */
#define SVGIR_SHADE_ACCUMULATE 2  /* backward: ADD into d_base_color / d_roughness / d_metallic / d_normals instead
                                     of overwriting them (caller-zeroed .grad buffers, multi-view accumulation) */
#define SVGIR_SHADE_VIEW_4X4 4    /* in->view3x3 points at the 4x4 world-view matrix (row stride 4) instead of a
                                     contiguous [3,3] copy of its rotation block */

typedef struct svgir_shade_in {
    const float* base_color;     /* [N,12] */
    const float* roughness;      /* [N,4]  */
    const float* metallic;       /* [N,4] or NULL (= 0, the reference's formula) */
    const float* normals;        /* [N,4,3] shading normals */
    const float* viewdirs;       /* [N,3]  */
    const float* radiance;       /* [N,Ns,3] cached indirect radiance */
    const float* visibility;     /* [N,Ns,1] */
    const float* incident_dirs;  /* [N,Ns,3] */
    const float* incident_areas; /* [N,Ns,1] */
    const float* env;            /* [env_h,env_w,3] parameter (mode 0) or map (mode 1) */
    const float* env_transform;  /* [3,3] or NULL (envmap.py:58-61) */
    float* env_act_scratch;      /* [env_h,env_w,3] scratch for the activated map */
    const float* view3x3;        /* [3,3] = viewmatrix[:3,:3] (row-vector convention), only for `pack` */
    /* Optional work list (device): only surfels surfel_list[0 .. *surfel_count) are shaded / differentiated;
     * the rows of every other surfel are left untouched in all outputs (the caller zero-fills them).
     * svgir_raster_preprocess produces the list of surfels that survive culling (state.vis_list), so
     * that culled surfels -- which no pixel reads and whose gradients are zero -- cost nothing. */
    const int32_t* surfel_list;  /* [<=N] surfel indices, or NULL = all N surfels */
    const int32_t* surfel_count; /* [1] device-side length of surfel_list */
    /* Fused view directions: with viewdirs == NULL the kernels evaluate normalize(campos - means3D[n]) themselves
     * (gaussian_renderer/svgss.py:95: F.normalize(camera_center - means3D)), and the backward adds the resulting
     * position gradient into svgir_shade_grads.d_means3D instead of writing d_viewdirs. */
    const float* means3D;        /* [N,3] or NULL */
    const float* campos;         /* [3] device, or NULL */
    const int32_t* skip_flag;    /* optional device flag: the backward does nothing when *skip_flag != 0 (the
                                    rasteriser's binning-overflow flag: an overflowed step contributes zero) */
    const float* env_taps;       /* optional [N,Ns,3] from svgir_env_taps for THESE incident_dirs / env size / transform:
                                    the lat-long texel corner and bilinear weights of every sample, so that the kernels
                                    skip acos / atan2 (the directions are fixed between two update_radiace calls and
                                    across the views and env maps of a relight sweep). Same values, bit for bit. */
} svgir_shade_in;

/* Env-map tap cache: taps[i] = { bits(x0 | y0 << 16) (signed 16-bit halves), wx1, wy1 } of the bilinear lat-long lookup
 * (scene/direct_light_map.py:70-83) of direction dirs[i] (after the optional 3x3 transform) in an env_h x env_w map. */
int svgir_env_taps(long long n, int env_h, int env_w, const float* transform, const float* dirs, float* taps, void* stream);

typedef struct svgir_shade_out {
    float* pbr; float* diffuse_light; float* specular; float* direct; float* indirect; /* [N,12]; any may be NULL */
    float* mean_visibility;  /* [N,1] sample means that render_view packs into `features` */
    float* mean_local;       /* [N,3]  (svgss.py:149-156); any of the four may be NULL */
    float* mean_incident;    /* [N,3] */
    float* mean_global;      /* [N,3] */
    /* Optional fused packing of render_view (gaussian_renderer/svgss.py:141-166): the five [N,12]
     * outputs may be column blocks of one `vfeatures` tensor (row_stride = VS floats between rows,
     * 0 = dense 12) and the means column blocks of `features` (mean_vis_stride / mean_stride, 0 = dense
     * 1 / 3). `pack`, if set, points at the first pass-through column and receives
     * base_color[12] | view-space shading normals[12, channel-major] | roughness[4]. */
    float* pack;
    /* Saved for svgir_shade_backward (replaces autograd's saved [N,Ns,12] tensors): per (vertex,channel)
     * mean_s max(n.w,0)*area*L for the env light and the cached radiance. With sum_indirect NULL the
     * kernel runs un-split and sum_direct receives the total (enough when no direct/indirect gradient
     * will be requested). */
    float* sum_direct;    /* [N,12] */
    float* sum_indirect;  /* [N,12] or NULL */
    int32_t row_stride, mean_vis_stride, mean_stride, reserved_;
} svgir_shade_out;

typedef struct svgir_shade_grads {
    const float* g_pbr; const float* g_diffuse_light; const float* g_specular; const float* g_direct;
    const float* g_indirect;            /* [N,12] upstream gradients, each may be NULL */
    const float* g_mean_visibility;     /* [N,1] or NULL */
    const float* g_mean_local;          /* [N,3] or NULL */
    const float* g_mean_incident;       /* [N,3] or NULL */
    const float* g_mean_global;         /* [N,3] or NULL */
    float* d_base_color;                /* [N,12] */
    float* d_roughness;                 /* [N,4]  */
    float* d_metallic;                  /* [N,4] or NULL */
    float* d_normals;                   /* [N,4,3] */
    float* d_viewdirs;                  /* [N,3] */
    float* d_radiance;                  /* [N,Ns,3] or NULL */
    float* d_visibility;                /* [N,Ns,1] or NULL */
    float* d_env;                       /* [env_h,env_w,3] zero-filled by the caller (accumulated into), or NULL */
    /* packed upstream gradients (see svgir_shade_out): strides of the g_* rows, and the gradient of
     * the pass-through columns, which is added into d_base_color / d_normals / d_roughness */
    const float* g_pack;
    const float* sum_direct;    /* [N,12] written by svgir_shade_forward (required) */
    const float* sum_indirect;  /* [N,12] or NULL; required when g_direct / g_indirect are given */
    float* d_env_scratch;       /* [SVGIR_SHADE_ENV_COPIES,env_h,env_w,4] scratch (library zeroes it); required with d_env */
    int32_t g_row_stride, g_mean_vis_stride, g_mean_stride, reserved_;
    float* d_means3D;           /* [N,3] += (atomic) -d/d(campos - means3D); required when in->viewdirs is NULL */
} svgir_shade_grads;

int svgir_shade_forward(const svgir_shade_cfg* cfg, const svgir_shade_in* in, const svgir_shade_out* out,
                        void* stream);
int svgir_shade_backward(const svgir_shade_cfg* cfg, const svgir_shade_in* in, const svgir_shade_grads* g,
                         void* stream);

/* Stand-alone env lookup for arbitrary directions [n,3] -> rgb [n,3]
 * (DirectLightMap.direct_light, scene/direct_light_map.py:70-83; EnvLight.direct_light,
 * scene/envmap.py:54-72 after its resize). Backward scatters into d_env (zero-filled by caller). */
int svgir_direct_light_forward(int n, int env_h, int env_w, int env_mode, const float* env,
                               float* env_act_scratch, const float* transform, const float* dirs,
                               float* out, void* stream);
int svgir_direct_light_backward(int n, int env_h, int env_w, int env_mode, const float* env,
                                const float* transform, const float* dirs, const float* g_out,
                                float* d_env, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LBVH over the surfels + visibility trace.  Replaces bvh_tracing._C (submodules/bvh/src/bindings.cpp:9-11):
 * create_bvh (src/bvh.cu:9-27 -> construct_bvh, src/construct.cu:147-265) and trace_bvh_opacity
 * (src/bvh.cu:89-116 -> trace_bvh_opacity_cuda, src/trace.cu:196-286), plus the torch code of
 * RayTracer.__init__ that prepares the leaf boxes (submodules/bvh/__init__.py:29-57). */
typedef struct svgir_bvh {
    int32_t P;               /* leaves (surfels) */
    int32_t reserved_;
    int32_t* nodes;          /* [2P-1,5] (parent, left, right, object, count) -- the reference layout */
    float* aabbs;            /* [2P-1,6] (lower xyz, upper xyz); rows P-1.. hold the leaf boxes on entry to
                                svgir_bvh_build and the Morton-sorted leaf boxes on return */
    uint64_t* morton;        /* [P] (code30 << 31) | object, ascending */
    float* packed;           /* [max(P-1,1),16] traversal records written by svgir_bvh_build (16-B aligned) */
    void* workspace;         /* svgir_bvh_workspace_bytes(P) bytes, 256-B aligned; scratch of the build only */
    size_t workspace_bytes;
} svgir_bvh;

size_t svgir_bvh_workspace_bytes(int P);

/* RayTracer.__init__ (submodules/bvh/__init__.py:31-57): initialises nodes (-1, counts 0/1) and aabbs
 * (+-100000) and writes the 8-corner leaf boxes mu +- 3 s R e_k into rows P-1.. ; bit-identical to the
 * reference's torch evaluation. rotations [P,4] un-normalised quaternions (r,x,y,z), 16-B aligned. */
int svgir_bvh_leaf_aabbs(int P, const float* means3D, const float* scales, const float* rotations,
                         int32_t* nodes, float* aabbs, void* stream);

/* construct_bvh: Morton codes of the leaf-box centroids, stable sort, Karras hierarchy, bottom-up boxes.
 * nodes / aabbs / morton come out bit-identical to the reference. */
int svgir_bvh_build(const svgir_bvh* bvh, void* stream);

/* Packs the per-surfel arguments of trace_bvh_opacity (means3D [P,3], symm_inv [P,6] = upper triangle of
 * Sigma^-1, opacity [P], normals [P,3]) into one 64-byte record per surfel (leaf_records [P,16]). */
int svgir_bvh_pack_leaves(int P, const float* means3D, const float* symm_inv, const float* opacity,
                          const float* normals, float* leaf_records, void* stream);

/* trace_bvh_opacity: ray r has direction rays_d[r] and origin rays_o[r / rays_per_origin] + origin_offset *
 * rays_d[r] (rays_per_origin = 1 and origin_offset = 0 give the reference call; RayTracer.trace_visibility
 * adds 0.05 d, submodules/bvh/__init__.py:63). Writes contributes [n_rays] int32 (0 for early-out rays) and
 * visibility [n_rays] (transmittance, or 0 once it drops below 0.9). */
int svgir_bvh_trace_opacity(const svgir_bvh* bvh, long long n_rays, const float* rays_o, const float* rays_d,
                            int rays_per_origin, float origin_offset, const float* leaf_records,
                            int32_t* contributes, float* visibility, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Incident-ray sampling.  Replaces sample_incident_rays (scene/gaussian_model.py:23-31) ->
 * fibonacci_sphere_sampling (utils/graphics_utils.py:9-37) + rotation_between_z (utils/sh_utils.py:36-68).
 * normals [N,3]; rand_u [N] = the per-surfel torch.rand offset in [0,1) used when training, or NULL
 * (random_rotate=False); incident_dirs [N,Ns,3]; incident_areas [N,Ns,1] (= 2*pi) or NULL. */
int svgir_sample_incident_rays(int N, int Ns, const float* normals, const float* rand_u, float* incident_dirs,
                               float* incident_areas, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SH-lit render_equation with per-surfel scalar materials (metallic workflow).  Replaces the R3DG-style
 * operators the reference declares in rgss-rasterization/render_equation.h:7-46 and defines in
 * rgss-rasterization/render_equation.cu (RenderEquationForwardCUDA :555-729, _complex :55-277,
 * RenderEquationBackwardCUDA :280-550).  base_color/normals/viewdirs [P,3], roughness/metallic [P,1],
 * incidents_shs [P,S_incident,3], direct_shs [1,S_direct,3], visibility_shs [P,S_vis,1]; S_* <= 16. */
typedef struct svgir_req_sh_cfg {
    int32_t P, S_incident, S_direct, S_vis, sample_num;
    int32_t is_training;   /* simple forward only: add rand_float*2*pi to the sample azimuth (:583) */
    int32_t legacy_exact;  /* backward: 1 = the reference's arithmetic including its known slips (dL_dn_d_i
                              overwritten :406, incident-SH loop over S_direct :453, no clamp masks);
                              0 = analytic gradient of the forward */
    int32_t debug;
} svgir_req_sh_cfg;

typedef struct svgir_req_sh_in {
    const float* base_color; const float* roughness; const float* metallic; const float* normals;
    const float* viewdirs; const float* incidents_shs; const float* direct_shs; const float* visibility_shs;
    const float* rand_float;  /* [P,sample_num,1] uniform randoms (torch::rand in the reference, :705) or NULL */
} svgir_req_sh_in;

typedef struct svgir_req_sh_out {
    float* pbr;            /* [P,3] */
    float* incident_dirs;  /* [P,sample_num,3] */
    float* diffuse_light;  /* [P,3] */
    /* the eight extra results of RenderEquationForwardCUDA_complex: all NULL (simple forward) or all set */
    float* incident_lights; float* local_incident_lights; float* global_incident_lights; /* [P,sample_num,3] */
    float* incident_visibility;   /* [P,sample_num,1] */
    float* local_diffuse_light;   /* [P,3] */
    float* accum;                 /* [P,1] */
    float* rgb_d; float* rgb_s;   /* [P,3] */
} svgir_req_sh_out;

typedef struct svgir_req_sh_grads {
    const float* incident_dirs;      /* [P,sample_num,3] as returned by the forward */
    const float* dL_dpbr;            /* [P,3] */
    const float* dL_ddiffuse_light;  /* [P,3] */
    float* dL_dbase_color; float* dL_droughness; float* dL_dmetallic; float* dL_dnormals; float* dL_dviewdirs;
    float* dL_dincidents_shs;        /* [P,S_incident,3] */
    float* dL_ddirect_shs;           /* [1,S_direct,3]; zeroed by the library, exact sum (the reference races here) */
    float* dL_dvisibility_shs;       /* [P,S_vis,1] */
} svgir_req_sh_grads;

int svgir_render_equation_sh_forward(const svgir_req_sh_cfg* cfg, const svgir_req_sh_in* in,
                                     const svgir_req_sh_out* out, void* stream);
int svgir_render_equation_sh_backward(const svgir_req_sh_cfg* cfg, const svgir_req_sh_in* in,
                                      const svgir_req_sh_grads* g, void* stream);

/* ---- fused G-buffer resolve + image loss (tail of a stage-2 training iteration) -----------------
 * Replaces the torch code between the rasteriser's forward and backward: gaussian_renderer/svgss.py:187-233
 * (divide feature/vfeature by opacity.clamp_min(1e-5), split, "pbr" = rgb_to_srgb(pbr*o + (1-o)*bg),
 * utils/graphics_utils.py:198-213) and the L1 terms of calculate_loss (svgss.py:280-294,
 * utils/loss_utils.py:33-34) plus the 0.02-weighted surface term (svgss.py:300-313):
 *   loss = mean|color-gt| + lambda_pbr*mean|pbr_srgb-gt| + lambda_normal*L_n,
 *   normal_mode 1 (the reference's term): L_n = cos_loss(n_shade, depth2normal(depth, mask, camera))
 *       = mean over the pixels with cos < 1 of (1 - <n_shade, d2n>)   (utils/loss_utils.py:117-119),
 *       d2n = normalised sum of the four cross products of the back-projected depth differences to the upper / left /
 *       lower / right neighbour, replicate padding, masked (utils/image_utils.py:61-125); its gradient flows into the
 *       shading normal AND into the rendered depth;
 *   normal_mode 0 (round-1 stand-in, kept for A/B): L_n = mean(1 - <n_shade, geo_normal>) over all pixels.
 * All images are channel-major [C,H,W] fp32 device pointers (the rasteriser's raw outputs). */
typedef struct svgir_train_loss_cfg {
    int32_t W, H, S, NV;            /* raw G-buffer: feature [S,H,W], vfeature [NV,H,W] (NV = VS/4) */
    int32_t pbr_ch, normal_ch;      /* first vfeature channel of pbr (0) and of the shading normal (6), svgss.py:210-213 */
    float lambda_pbr, lambda_normal;
    const float* bg;                /* [3] device */
    int32_t normal_mode;            /* see above */
    /* depth2normal camera terms (image_utils.py:73-81): x is divided by focal_x = fov2focal(FoVy, H) and y by
     * focal_y = fov2focal(FoVx, W) -- the reference's own pairing --, the principal point is prcppoint * (W, H) */
    float inv_focal_x, inv_focal_y, cx, cy;
    int32_t reserved_;
} svgir_train_loss_cfg;

typedef struct svgir_train_loss_in {
    const float* color;       /* [3,H,W] "render" */
    const float* geo_normal;  /* [3,H,W] rasteriser's normal output (raw) */
    const float* opacity;     /* [1,H,W] */
    const float* vfeature;    /* [NV,H,W] raw (opacity-premultiplied) */
    const float* gt;          /* [3,H,W] */
    const float* depth;       /* [1,H,W] rasteriser's depth output; required with normal_mode 1 */
    const float* mask;        /* [1,H,W] image mask (nonzero = inside), or NULL = all ones */
} svgir_train_loss_in;

typedef struct svgir_train_loss_grads {   /* dL/d(rasteriser outputs); every element of a non-NULL image is written */
    float* color;       /* [3,H,W] */
    float* geo_normal;  /* [3,H,W] */
    float* depth;       /* [1,H,W] (zeros in normal_mode 0, may then be NULL; required with normal_mode 1) */
    float* opacity;     /* [1,H,W] */
    float* feature;     /* [S,H,W] (zeros; may be NULL) */
    float* vfeature;    /* [NV,H,W] */
} svgir_train_loss_grads;

int svgir_train_loss_blocks(int W, int H);
/* loss[8] = total, l1, l1_pbr, normal term, number of pixels in the normal term's mean, 3 unused; partials:
 * 4*svgir_train_loss_blocks floats of scratch; counter: one zero-initialised u32 (left at zero on return).
 * Deterministic summation order. */
int svgir_train_loss_forward(const svgir_train_loss_cfg* cfg, const svgir_train_loss_in* in, float* loss,
                             float* partials, unsigned int* counter, void* stream);
/* grad_loss: device scalar dL/dloss (NULL = 1). loss: the buffer svgir_train_loss_forward filled (normal_mode 1
 * reads the pixel count of the cos_loss mean from loss[4]; may be NULL in normal_mode 0). */
int svgir_train_loss_backward(const svgir_train_loss_cfg* cfg, const svgir_train_loss_in* in,
                              const float* grad_loss, const float* loss, const svgir_train_loss_grads* g, void* stream);

/* ---- smoothness terms of calculate_loss (gaussian_renderer/svgss.py:366-390) --------------------------------
 * first_order_edge_aware_loss(data * mask, img * mask) (utils/loss_utils.py:103-104): with the normalised Sobel
 * gradient of kornia.filters.spatial_gradient (order 1; kernels / 8, replicate padding),
 *   loss = mean over [C,H,W] of |d/dx data| exp(-|d/dx img|) + |d/dy data| exp(-|d/dy img|).
 * data, img [C,H,W]; mask [1,H,W] or NULL. forward: loss_out[0]; partials = svgir_edge_aware_blocks floats,
 * counter one zero-initialised u32. backward: d_data [C,H,W] = grad_out[0] (NULL = 1) * dloss/ddata (img: no gradient;
 * the library zero-fills d_data, then scatters). */
int svgir_edge_aware_blocks(int C, int H, int W);
int svgir_edge_aware_forward(int C, int H, int W, const float* data, const float* img, const float* mask,
                             float* loss_out, float* partials, unsigned int* counter, void* stream);
int svgir_edge_aware_backward(int C, int H, int W, const float* data, const float* img, const float* mask,
                              const float* grad_out, float* d_data, void* stream);
/* tv_loss (utils/loss_utils.py:112-116) of x[c][h][w] addressed with element strides (sc, sh, sw), so the env map
 * [He,We,3] is read in place as env.permute(2,0,1) (svgss.py:388): mean((x[h+1]-x[h])^2) + mean((x[w+1]-x[w])^2).
 * One launch computes loss_out[0] and, if d_x is not NULL, d_x (same strides) = grad_out[0] (NULL = 1) * dloss/dx. */
int svgir_tv_loss(int C, int H, int W, long long sc, long long sh, long long sw, const float* x, const float* grad_out,
                  float* loss_out, float* d_x, void* stream);

/* ---- G-buffer resolve of an evaluation / relighting frame -----------------------------------------
 * The torch tail of render_view's eval branch (gaussian_renderer/svgss.py:187-262 with is_training=False:
 * divide by opacity.clamp_min(1e-5), split, rgb_to_srgb (utils/graphics_utils.py:198-213), opacity filter
 * r*o + (1-o)*bg) as ONE kernel instead of ~45 elementwise launches. Inputs are the rasteriser's raw
 * outputs for the eval G-buffer: feature [7,H,W] = light 3 | local light 3 | visibility 1; vfeature [16,H,W] =
 * pbr 3 | base colour 3 | shading normal 3 | roughness 1 | direct 3 | indirect 3 (svgss.py:150-166).
 * Every output is [3,H,W] (roughness / visibility are broadcast against the 3-channel background exactly as
 * the reference's opacity_filter does); NULL outputs are skipped. */
typedef struct svgir_resolve_eval_out {
    float* pbr;          /* srgb(pbr*o + (1-o)*bg) */
    float* normal;       /* shading normal / o (no filter) */
    float* base_color;   /* filter(srgb(base)) */
    float* roughness;    /* filter(rough), 3 channels */
    float* lights;       /* filter(srgb(light)) */
    float* local_lights; /* filter(srgb(local)) */
    float* visibility;   /* filter(vis), 3 channels */
    float* direct;       /* srgb(direct) */
    float* indirect;     /* srgb(indirect) */
} svgir_resolve_eval_out;
int svgir_resolve_eval(int W, int H, const float* bg, const float* opacity, const float* feature,
                       const float* vfeature, const svgir_resolve_eval_out* out, void* stream);

/* ---- SSIM and its gradient (verified on B200 against goldens from the reference's own ssim) ----
 * utils/loss_utils.py:21-62 `ssim(img1, img2)`: 11-tap Gaussian window (sigma 1.5), zero padding, C1 = 0.01^2,
 * C2 = 0.03^2, mean over [C,H,W]; called on the splatted colour and on the PBR image (svgss.py:282-293).
 * forward: ssim_out[0] = ssim; gmaps [3,C,H,W] (optional) receives the per-pixel partials the backward needs;
 * partials: svgir_ssim_blocks(C,H,W) floats of scratch; counter: one zero-initialised u32 (left at zero).
 * backward: d_img1 [C,H,W] = grad_out[0] (NULL = 1) * d ssim / d img1; img2 is the ground truth (no gradient). */
int svgir_ssim_blocks(int C, int H, int W);
int svgir_ssim_forward(int C, int H, int W, const float* img1, const float* img2, float* ssim_out, float* gmaps,
                       float* partials, unsigned int* counter, void* stream);
int svgir_ssim_backward(int C, int H, int W, const float* img1, const float* img2, const float* gmaps,
                        const float* grad_out, float* d_img1, void* stream);

/* ---- optimiser step and densification ----------------------------------------------------
 * Fused Adam: all parameter groups of torch.optim.Adam(l, eps=1e-15) (scene/gaussian_model.py:737-773) in one launch,
 * gradients read in place from the flat gradient bucket, NaN gradients replaced on the fly as
 * replace_nangrad_to_zero does (:775-795; nan_fix != 0: NaN -> nan_value). step = 1, 2, ... (bias correction). */
#define SVGIR_ADAM_MAX_GROUPS 16
typedef struct svgir_adam_group {
    float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
    long long numel;
    float lr;
    int32_t nan_fix;
    float nan_value;
    int32_t reserved_;
} svgir_adam_group;
int svgir_adam_step(const svgir_adam_group* groups /* host */, int n_groups, double beta1, double beta2, double eps, int step,
                    void* stream);

/* add_densification_stats (gaussian_model.py:1270-1276): weights_accum += weights; for radii > 0:
 * xyz_grad_accum += |viewspace_grad.xy|, denom += 1 (and max_radii2D = max(max_radii2D, radii) if given, train.py). */
int svgir_densify_stats(int P, const float* viewspace_grad /* [P,3] */, const int32_t* radii, const float* weights,
                        float* weights_accum, float* xyz_grad_accum, float* denom, float* max_radii2D, void* stream);

/* densify_and_prune (gaussian_model.py:1136-1250) as device-side compaction. decide: per-surfel flags and output
 * counts (keep / clone / split, each after the final prune mask); the caller scans the three count arrays
 * (inclusive, int64); index: source row + kind of every row of the new model laid out as the reference leaves it --
 * surviving originals, clones, first split copies, second split copies; gather_rows: any [P,K] fp32 tensor -> [P',K]
 * (new rows zero / constant filled on request: Adam moments, statistics); split: positions and scales of the split
 * children from caller-drawn N(0,1) samples [2 nC, 3]. */
typedef struct svgir_densify_cfg {
    int32_t P;
    float grad_threshold, grad_normal_threshold, percent_dense, extent, min_opacity, weights_threshold;
    int32_t use_screen_size;   /* max_screen_size given: also prune max(scaling) > 0.1 extent (:1237-1239) */
} svgir_densify_cfg;
int svgir_densify_decide(const svgir_densify_cfg* cfg, const float* xyz_grad_accum, const float* normal_grad_accum /* or NULL */,
                         const float* denom, const float* scaling_raw /* [P,3] log-scales */, const float* opacity_raw /* [P] logits */,
                         const float* weights_accum, uint8_t* flags, int32_t* keep, int32_t* nclone, int32_t* nsplit, void* stream);
int svgir_densify_index(int P, const int32_t* keep, const int32_t* nclone, const int32_t* nsplit, const int64_t* keep_scan,
                        const int64_t* clone_scan, const int64_t* split_scan, long long nA, long long nB, long long nC,
                        int32_t* src /* [nA+nB+2nC] */, uint8_t* kind, void* stream);
int svgir_gather_rows(long long n_rows, int K, const float* src_rows, const int32_t* index, const uint8_t* kind /* or NULL */,
                      int zero_new, float new_value, float* dst, void* stream);
int svgir_densify_split(long long n_new, long long first_split, const int32_t* src, const uint8_t* kind, const float* xyz_old,
                        const float* scaling_old, const float* rotation_old, const float* normal_samples, float* xyz_new,
                        float* scaling_new, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Radiance cache and radiance-consistency loss.  Replaces, for stage 2:
 *   GaussianModel.update_radiace (scene/gaussian_model.py:469-522) -> Renderer.render_radiance_with_sampling_SH
 *     (pbgi/renderer.py:596-615) -> the Slang kernels render_radiance_with_sampling_SH / gs_bvh_hit / ellipse_hit
 *     (pbgi/bvhworkers/intersect_test.slang:1879-1991, :251-443, :94-149), and
 *   GaussianModel.get_radiance_loss (scene/gaussian_model.py:544-575) -> Renderer.render_irradiance_sample
 *     (pbgi/renderer.py:181-227, :748-751) -> render_irradiance_sample (intersect_test.slang:1143-1378) with
 *     shading_brdf_simple (pbgi/bvhworkers/pbr.slang:283-330) and DirectLightMap.direct_light.
 * The reference builds a second LBVH for this (pbgi/renderer.py:586-590, cube leaf boxes of 3*max|scale|,
 * get_elements.slang:58-82); here the closest-hit query walks the SAME tree svgir_bvh_build made for the visibility
 * trace (its leaf boxes contain every point with disM <= 9, the only points a hit can lie on).
 * Where the reference's kernels are racy or depend on the traversal order the intended value is computed; every such
 * place is listed in DESIGN.md Appendix C (R1-R5). */
#define SVGIR_RADIANCE_RECORD_FLOATS 32
#define SVGIR_RADIANCE_SCRATCH_FLOATS 2048
#define SVGIR_RADIANCE_ENV_COPIES 32

/* One 128-byte record per surfel for the closest-hit query: centre, opacity, the rotation columns
 * (matrixFromRotationQuaternions, intersect_test.slang:224-249; rotations [P,4] raw (r,x,y,z)), the first two scales
 * (scales [P,scale_stride], scale_stride >= 2), the normal the facing test uses (normals [P,3] = proxy_normals) and
 * Sigma^-1 (symm_inv [P,6], upper triangle as strip_symmetric orders it). records [P,32], 16-B aligned. */
int svgir_radiance_pack_surfels(int P, const float* means3D, const float* scales, int scale_stride, const float* rotations,
                                const float* normals, const float* opacity, const float* symm_inv, float* records,
                                void* stream);

/* render_radiance_with_sampling_SH: ray (n,s) starts at origins[n] along normalize(dirs[n,s]) and is advanced from hit
 * to hit (closest facing surfel with alpha >= 1/255 in [t_min, 0.2), t_min 0.042 then 0.01) while T > 0.001, summing
 * eval_sh(shs[hit], centre - origin) * alpha * T. radiance [N,S,3] (clamped to [0,10]), visibility [N,S] (T, or 0 once
 * T < 0.2), hit_index [N,S] (first hit or -1), uv [N,S,2] (of the first hit). shs [P,16,3].
 * Which surfel a ray ignores:
 *   self_mod == 0  the ray ignores its own surfel, index first_index + n (what the kernel means to do);
 *   self_mod  > 0  the reference's behaviour when update_radiace feeds it chunks of self_mod surfels: the kernel
 *                  compares the hit with the CHUNK-LOCAL index (intersect_test.slang:1932), i.e. ignores surfel
 *                  (first_index + n) % self_mod. */
int svgir_radiance_cache_build(const svgir_bvh* bvh, int N, int S, int first_index, int self_mod, const float* origins,
                               const float* dirs, const float* records, const float* shs, float* radiance,
                               float* visibility, int32_t* hit_index, float* uv, void* stream);

#define SVGIR_RADIANCE_ENV_READY 1      /* env_act_scratch already holds the activated env map */
#define SVGIR_RADIANCE_BWD_REFERENCE_GRID 2 /* backward: only secondary sample 0 carries gradient, S times over -- what
                                               the reference's backward launch computes (grid (N/256,1,S) with the
                                               sample index read from y, pbgi/renderer.py:223) */
#define SVGIR_RADIANCE_NORMALS_VERTEX_MAJOR 4 /* normals is [P,4,3] (get_shading_normal as it is, element 3*v + c) instead of
                                                the transposed [P,12] the reference hands to its kernel (:551) */
typedef struct svgir_radiance_loss_cfg {
    int32_t P, S, env_h, env_w;
    int32_t env_mode;       /* as svgir_shade_cfg: 0 = learnable map (softplus, x2), 1 = fixed map */
    int32_t flags;
    int32_t rough_stride;   /* floats per row of roughness / d_roughness (column 0 is used, intersect_test.slang:1277) */
    int32_t reserved_;
} svgir_radiance_loss_cfg;

typedef struct svgir_radiance_loss_in {
    const float* means3D;        /* [P,3] */
    const float* campos;         /* [3] device */
    const float* geo_normal;     /* [P,3] */
    const float* incident_dirs;  /* [P,S,3]  _incident_dirs */
    const float* incident_areas; /* [P,S]    _incident_areas */
    const float* visibility;     /* [P,S]    _visibility_tracing */
    const int32_t* hit_index;    /* [P,S]    hemi_index_buffers */
    const float* uv;             /* [P,S,2]  uv_buffers */
    const float* radiances;      /* [P,S,3]  _radiances */
    const float* radiance_ratio; /* device scalar or NULL (= 1) */
    const float* normals;        /* [P,12] shading normals, element 4*c + v */
    const float* albedo;         /* [P,12] element 4*c + v */
    const float* roughness;      /* [P,rough_stride] */
    const float* env;            /* [env_h,env_w,3] raw parameter */
    float* env_act_scratch;      /* [env_h*env_w*3] */
    const int32_t* skip_flag;    /* optional device flag: nonzero = the backward adds nothing (the fused training step
                                    passes its binning-overflow flag: an overflowed step contributes no gradient) */
    const float* env_taps;       /* optional [P,S,3] from svgir_env_taps(incident_dirs, env_h, env_w): skips acos / atan2 */
} svgir_radiance_loss_in;

/* loss [1] = mean |irradiance - nan_to_num(radiances[n, sel[n]] * ratio)| over [P,3]. Written: irradiance [P,3];
 * sample_index [P] (= max_idx, gaussian_model.py:565; may be NULL); saved [P,8] (16-B aligned: per surfel the view
 * vector of the selected sample, the surfel it hit, the target colour -- what the backward needs). scratch
 * [SVGIR_RADIANCE_SCRATCH_FLOATS]. The sum over secondary samples is complete and fixed in order (the reference adds
 * with a non-atomic read-modify-write from S threads, intersect_test.slang:1354-1356). */
int svgir_radiance_loss_forward(const svgir_radiance_loss_cfg* cfg, const svgir_radiance_loss_in* in, float* loss,
                                float* irradiance, int32_t* sample_index, float* saved, float* scratch, void* stream);

/* Adds grad_loss * dloss/d{albedo, roughness[:,0], env} into d_albedo [P,12], d_roughness [P,rough_stride] (atomic;
 * zero them first) and d_env [env_h,env_w,3] (+=, through d_env_scratch [SVGIR_RADIANCE_ENV_COPIES*env_h*env_w*4]). grad_loss: device scalar or
 * NULL (= 1). irradiance / saved: as the forward wrote them. Any of the three destinations may be NULL. */
int svgir_radiance_loss_backward(const svgir_radiance_loss_cfg* cfg, const svgir_radiance_loss_in* in,
                                 const float* grad_loss, const float* irradiance, const float* saved,
                                 float* d_albedo, float* d_roughness, float* d_env, float* d_env_scratch, void* stream);

/* Both in one pass, for a caller that knows the upstream gradient beforehand (the training step: grad_loss =
 * lambda_radiance): the gradients of a surfel are accumulated right after its irradiance, while its rows are in cache.
 * With in->skip_flag set (nonzero on the device) the loss is still evaluated and the gradients are left untouched. */
int svgir_radiance_loss_forward_backward(const svgir_radiance_loss_cfg* cfg, const svgir_radiance_loss_in* in,
                                         const float* grad_loss, float* loss, float* irradiance, int32_t* sample_index,
                                         float* saved, float* scratch, float* d_albedo, float* d_roughness, float* d_env,
                                         float* d_env_scratch, void* stream);

/* ---- per-surfel gradient all-reduce over NVLink peer memory (view-sharded data parallelism) ------
 * New: the reference is single-process / single-GPU (train.py:108-143). Every rank keeps
 * its flat gradient buffer in a symmetric allocation that is peer-mapped into all ranks of the box; the
 * backward kernels write their gradients straight into it, and ONE kernel per rank then sums the buffers
 * in place: a flag barrier over peer memory, rank r reduces slice r of all `world` buffers (NVSwitch
 * multicast `multimem.ld_reduce` when `multicast` is set, otherwise 128-bit peer loads), writes the sum
 * back into every rank's buffer (`multimem.st` / peer stores), flag barrier. No staging copy, no NCCL.
 * `flags[i]` is rank i's flag area: SVGIR_PEER_FLAG_WORDS zero-initialised u32, never touched by the host
 * afterwards (the put/wait protocol leaves it zeroed, so the launch can be replayed from a CUDA graph).
 * All ranks must launch with the same numel; numel*4 bytes must be 16-byte aligned per slice (the kernel
 * rounds slices to 4 floats). The grid is at most one CTA per SM so all CTAs are co-resident. */
#define SVGIR_MAX_PEERS 8
#define SVGIR_PEER_BLOCKS 128
#define SVGIR_PEER_BANKS 4   /* independent flag banks: one per all-reduce that may be in flight at the same time */
#define SVGIR_PEER_FLAG_WORDS (SVGIR_PEER_BANKS * 2 * SVGIR_PEER_BLOCKS * SVGIR_MAX_PEERS)
typedef struct svgir_peer_comm {
    int32_t world, rank;
    float* bufs[SVGIR_MAX_PEERS];            /* bufs[i]: rank i's buffer as mapped in THIS process */
    unsigned int* flags[SVGIR_MAX_PEERS];    /* flags[i]: rank i's flag area as mapped in this process */
    float* multicast;                        /* multicast mapping of the buffers (NVLS), or NULL */
} svgir_peer_comm;
int svgir_peer_allreduce(const svgir_peer_comm* comm, long long numel, void* stream);
/* The same on the sub-range [offset, offset+numel) of the buffers (both multiples of 4 floats), with flag bank
 * `bank` (< SVGIR_PEER_BANKS; two launches that can overlap in time must use different banks) and `grid` CTAs
 * (0 = default; the NVLink ports saturate at ~16). Used to sum the rasteriser-side gradients on a side stream while
 * the shading backward still runs: svgir_shade_reserve_sms(n) makes the shading kernels leave n SMs to it. */
int svgir_peer_allreduce_range(const svgir_peer_comm* comm, long long offset, long long numel, int bank, int grid,
                               void* stream);
void svgir_shade_reserve_sms(int n);

#ifdef __cplusplus
}
#endif
#endif /* SVGIR_B200_H_ */
