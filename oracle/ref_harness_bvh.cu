// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI harness around the reference LBVH kernels: construct_bvh
// (/root/reference/submodules/bvh/src/construct.cu:147-265) and trace_bvh_opacity_cuda
// (src/trace.cu:196-286). The reference sources are compiled where they lie (oracle/Makefile; the
// two extended lambdas at construct.cu:166,186 get trailing return types via sed on a temporary
// copy because CUDA 12.9's CCCL cannot deduce them -- behaviour-neutral); this file only supplies
// what the torch glue src/bvh.cu:9-27,89-116 supplies, with raw device pointers instead of tensors.
#include <cstdint>
#include <cuda_runtime.h>
#include "construct.cuh"   // reference headers (found through -I, not copied)
#include "trace.cuh"

extern "C" {

// nodes [2P-1,5] int32 and aabbs [2P-1,6] float come initialised as RayTracer.__init__ does
// (submodules/bvh/__init__.py:31-57); morton [P] uint64 is written.
int ref_bvh_build(int P, const float* means3D, const float* scales, const float* rotations,
                  int32_t* nodes, float* aabbs, uint64_t* morton) {
    try {
        construct_bvh(P, means3D, scales, rotations, nodes, aabbs, morton);
    } catch (...) { return -1; }
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

// contributes zero-filled, visibility one-filled by the caller (bvh.cu:103-104).
int ref_bvh_trace_opacity(int num_rays, int32_t* nodes, float* aabbs, float* rays_o, float* rays_d,
                          float* means3D, float* covs3D, float* opacities, float* normals,
                          int32_t* contributes, float* visibility) {
    try {
        trace_bvh_opacity_cuda(num_rays, nodes, aabbs, (float3*)rays_o, (float3*)rays_d, (float3*)means3D,
                               covs3D, opacities, (float3*)normals, contributes, visibility);
    } catch (...) { return -1; }
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
