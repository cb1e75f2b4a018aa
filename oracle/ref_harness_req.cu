// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI harness around the reference's (unbuilt) SH render_equation kernels
// (/root/reference/rgss-rasterization/render_equation.cu). That file mixes the kernels, their
// plain launchers (`render_equation_forward_cuda` etc., glm::vec3 pointers) and torch glue in one
// translation unit, so it is compiled where it lies with the torch headers on the include path
// (oracle/Makefile) and this harness calls the plain launchers with raw device pointers.
#include <cuda_runtime.h>
#include <glm/glm.hpp>

void render_equation_forward_cuda(
    const int P, const int S_incident, const int S_direct, const int S_vis, const glm::vec3* base_color,
    const float* roughness, const float* metallic, const glm::vec3* normals, const glm::vec3* viewdirs,
    const glm::vec3* incidents_shs, const glm::vec3* direct_shs, const float* visibility_shs, const int sample_num,
    const bool is_training, const float* rand_float, glm::vec3* incident_dirs, glm::vec3* out_pbr,
    glm::vec3* out_diffuse_light);

void render_equation_forward_complex_cuda(
    const int P, const int S_incident, const int S_direct, const int S_vis, const glm::vec3* base_color,
    const float* roughness, const float* metallic, const glm::vec3* normals, const glm::vec3* viewdirs,
    const glm::vec3* incidents_shs, const glm::vec3* direct_shs, const float* visibility_shs, const int sample_num,
    glm::vec3* incident_dirs, glm::vec3* out_pbr, glm::vec3* incident_lights, glm::vec3* local_incident_lights,
    glm::vec3* global_incident_lights, float* incident_visibility, glm::vec3* diffuse_light,
    glm::vec3* local_diffuse_light, float* accum, glm::vec3* rgb_d, glm::vec3* rgb_s);

void render_equation_backward_cuda(
    const int P, const int S_incident, const int S_direct, const int S_vis, const glm::vec3* base_color,
    const float* roughness, const float* metallic, const glm::vec3* normals, const glm::vec3* viewdirs,
    const glm::vec3* incidents_shs, const glm::vec3* direct_shs, const float* visibility_shs, const int sample_num,
    const glm::vec3* incident_dirs, const glm::vec3* dL_dpbrs, const glm::vec3* dL_ddiffuse_light,
    glm::vec3* dL_dbase_color, float* dL_droughness, float* dL_dmetallic, glm::vec3* dL_dnormals,
    glm::vec3* dL_dviewdirs, glm::vec3* dL_dincidents_shs, glm::vec3* dL_ddirect_shs, float* dL_dvisibility_shs);

#define V3(p) ((glm::vec3*)(p))
#define CV3(p) ((const glm::vec3*)(p))

extern "C" {

int ref_req_forward(int P, int S_incident, int S_direct, int S_vis, const float* base_color, const float* roughness,
                    const float* metallic, const float* normals, const float* viewdirs, const float* incidents_shs,
                    const float* direct_shs, const float* visibility_shs, int sample_num, int is_training,
                    const float* rand_float, float* incident_dirs, float* pbr, float* diffuse_light) {
    render_equation_forward_cuda(P, S_incident, S_direct, S_vis, CV3(base_color), roughness, metallic, CV3(normals),
                                 CV3(viewdirs), CV3(incidents_shs), CV3(direct_shs), visibility_shs, sample_num,
                                 is_training != 0, rand_float, V3(incident_dirs), V3(pbr), V3(diffuse_light));
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

int ref_req_forward_complex(int P, int S_incident, int S_direct, int S_vis, const float* base_color,
                            const float* roughness, const float* metallic, const float* normals, const float* viewdirs,
                            const float* incidents_shs, const float* direct_shs, const float* visibility_shs,
                            int sample_num, float* incident_dirs, float* pbr, float* incident_lights,
                            float* local_incident_lights, float* global_incident_lights, float* incident_visibility,
                            float* diffuse_light, float* local_diffuse_light, float* accum, float* rgb_d, float* rgb_s) {
    render_equation_forward_complex_cuda(P, S_incident, S_direct, S_vis, CV3(base_color), roughness, metallic, CV3(normals),
                                         CV3(viewdirs), CV3(incidents_shs), CV3(direct_shs), visibility_shs, sample_num,
                                         V3(incident_dirs), V3(pbr), V3(incident_lights), V3(local_incident_lights),
                                         V3(global_incident_lights), incident_visibility, V3(diffuse_light),
                                         V3(local_diffuse_light), accum, V3(rgb_d), V3(rgb_s));
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

// Gradient outputs must be zero-filled by the caller (RenderEquationBackwardCUDA uses torch::zeros, :501-508).
int ref_req_backward(int P, int S_incident, int S_direct, int S_vis, const float* base_color, const float* roughness,
                     const float* metallic, const float* normals, const float* viewdirs, const float* incidents_shs,
                     const float* direct_shs, const float* visibility_shs, int sample_num, const float* incident_dirs,
                     const float* dL_dpbrs, const float* dL_ddiffuse_light, float* dL_dbase_color, float* dL_droughness,
                     float* dL_dmetallic, float* dL_dnormals, float* dL_dviewdirs, float* dL_dincidents_shs,
                     float* dL_ddirect_shs, float* dL_dvisibility_shs) {
    render_equation_backward_cuda(P, S_incident, S_direct, S_vis, CV3(base_color), roughness, metallic, CV3(normals),
                                  CV3(viewdirs), CV3(incidents_shs), CV3(direct_shs), visibility_shs, sample_num,
                                  CV3(incident_dirs), CV3(dL_dpbrs), CV3(dL_ddiffuse_light), V3(dL_dbase_color),
                                  dL_droughness, dL_dmetallic, V3(dL_dnormals), V3(dL_dviewdirs), V3(dL_dincidents_shs),
                                  V3(dL_ddirect_shs), dL_dvisibility_shs);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
