/*
 * oracle.h -- CPU restatement of Realtime-Vulkan-Hair's guide-strand physics step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it, and there only as the checker or the reported CPU baseline.  The product
 * path (realtime-vulkan-hair_b200/csrc) never links, loads or calls this.
 *
 * What it restates (all file:line relative to the reference tree):
 *   src/shaders/compute.comp:124-303  main(): integrate + collide + FTL, corrected
 *                                     velocity + grid splat, grid gather + friction
 *   src/shaders/compute.comp:64-79    ellipsoid helpers
 *   src/shaders/compute.comp:83-121   random / noise / fbm (the commented-out wind)
 *   src/Scene.h:23-39                 Collider ctor (T*Rz*Ry*Rx*S, inverse, transpose)
 *   src/Scene.cpp:110-120             translateSphere
 *   src/main.cpp:229-237              the six colliders of the shipped scene
 *   src/Strand.cpp:157-175            initial strand state from follicle + normal
 *   src/Renderer.cpp:2063-2070        clear grid, then one invocation per strand
 * Third-party arithmetic restated: glm 0.9.9.0 (vendored in the reference under
 * external/glm): translate/rotate/scale, mat4*mat4, mat4*vec4, inverse, transpose,
 * dot/length/distance/normalize/mix -- same operation ORDER as glm so that this file
 * is bit-identical (built with -ffp-contract=off) to the reference shader source
 * compiled as C++ against that glm (oracle/_ref, see oracle/ref_build/).
 *
 * Parity pin: oracle/_ref (the reference's own compute.comp text and Scene.h /
 * Strand.cpp compiled here, see oracle/ref_build/README) -- tests/test_oracle_vs_ref.py
 * checks bit-equality in this container and freezes the outputs as tests/golden/.
 *
 * Semantics chosen where the shader is undefined:
 *   - barrier() between splat and gather is treated as a global barrier;
 *   - only S invocations run (the shader has no idx<S guard);
 *   - grid accumulators are int64; ORC_GRID_INT32_WRAP reads them back through
 *     their low 32 bits, which equals int32 atomics with wrap-around.
 */
#ifndef RVH_ORACLE_H
#define RVH_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_GRID_ON         = 1,   /* phases P2 splat + P3 gather (compute.comp:211-298)  */
    ORC_WIND_A          = 2,   /* compute.comp:151 (commented out in the reference)   */
    ORC_WIND_B          = 4,   /* compute.comp:152 (commented out in the reference)   */
    ORC_GRID_INT32_WRAP = 8,   /* emulate the reference's int32 GridCell (Scene.h:42) */
    /* Extensions named by the north star that the reference does NOT contain; these modes are the
     * definition the CUDA path is checked against (include/rvh.h RVH_SDF_ON / RVH_REPULSION_ON). */
    ORC_SDF_ON          = 64,  /* colliders 1..n replaced by a sampled head SDF (orc_set_head_sdf) */
    ORC_REPULSION_ON    = 128  /* v -= repulsion * h*grad(rho)/sum(D) with the friction gather   */
};

typedef struct {
    int   num_strands;      /* S                                                      */
    int   num_points;       /* N, root included                  compute.comp:5       */
    float rest_length;      /* "radius" = 2.5f/(N-1)             compute.comp:139-140 */
    float gravity_y;        /* -9.8f                             compute.comp:150     */
    float damping;          /* 0.998f                            compute.comp:7       */
    float vmax;             /* 10.f                              compute.comp:198     */
    float penalty_k;        /* 1900.f                            compute.comp:163,173 */
    float sphere_radius;    /* 1.f                               compute.comp:161     */
    int   grid_dim;         /* 64                                compute.comp:9       */
    float grid_extent;      /* 7.f                               compute.comp:10      */
    float grid_origin[3];   /* -3,-2,-5                          compute.comp:206     */
    float grid_scale;       /* 1e6f                              compute.comp:11      */
    float friction;         /* 0.08f                             compute.comp:296     */
    int   flags;
    int   num_colliders;    /* 6; index 0 is the sphere          compute.comp:8,160   */
    float repulsion;        /* 0.2f; ORC_REPULSION_ON only (extension)                */
} orc_params;

/* Fill every field with the reference constant for (S, N). */
void orc_default_params(orc_params* p, int num_strands, int num_points);

/* Collider = 3 column-major mat4 (transform, inv, invTrans) = 48 floats (Scene.h:23-26). */
void orc_collider_build(const float trans[3], const float rot_deg[3], const float scale[3],
                        float out48[48]);
void orc_collider_translate(float c48[48], const float translation[3]);  /* Scene.cpp:110-120 */
void orc_default_colliders(float out[6 * 48]);                           /* main.cpp:229-237  */

float orc_fbm_time(float total_time);  /* fbm(vec2(sin T, cos T)), compute.comp:83-121,152 */

/* Strand state is the reference's AoS, N generalised: float [S][3][N][4]
 * (curvePoints, curveVels, correctionVecs; Strand.h:11-15).  grid is int64 [G^3][4]
 * (vx, vy, vz, density).  orc_step clears the grid first (Renderer.cpp:2063). */
void orc_step(const orc_params* p, const float* colliders48, float dt, float total_time,
              float* strands, int64_t* grid);

/* The three phases separately, for targeted tests (same code orc_step runs). */
void orc_phase_integrate(const orc_params* p, const float* colliders48, float dt,
                         float total_time, float* strands);           /* compute.comp:144-202 */
void orc_phase_splat(const orc_params* p, float dt, float* strands, int64_t* grid);
                                                                      /* compute.comp:211-253 */
void orc_phase_gather(const orc_params* p, float* strands, const int64_t* grid);
                                                                      /* compute.comp:257-298 */

/* OpenMP build of orc_step (per-thread grids summed before the gather): the CPU
 * baseline bench.py reports.  Bit-identical to orc_step (integer grid). */
void orc_step_parallel(const orc_params* p, const float* colliders48, float dt,
                       float total_time, float* strands, int64_t* grid, int num_threads);
int  orc_max_threads(void);

/* ---- extensions (not in the reference) ------------------------------------------------ */
/* Head SDF: node values [nz][ny][nx], x fastest, node (i,j,k) at origin + cell*(i,j,k), negative
 * inside.  The pointer is kept (not copied) until the next call; NULL clears it. */
void orc_set_head_sdf(const float* sdf, const int dim[3], const float origin[3], float cell);
/* Trilinear sample exactly as the step uses it; returns 0 when the point's cell is outside. */
int  orc_sdf_sample(const float p[3], float* d, float grad[3]);
/* Bakes (CPU twins of k_sdf_bake_colliders / k_sdf_bake_mesh): */
void orc_sdf_bake_colliders(const float* colliders48, int num_colliders, const int dim[3],
                            const float origin[3], float cell, float* out);
void orc_sdf_bake_mesh(const float* verts, const int* tris, int ntris, const int dim[3],
                       const float origin[3], float cell, float* out);

/* ---- guide strand -> render strands: hair.tesc:19-20 + hair.tese -------
 * CPU restatement of what the tessellator + hair.tese emit for every guide strand: `isolines`
 * line strips of divisions+1 vertices.  pos_width / tangent_u are float [S][isolines][divisions+1][4].
 * Stated choices (the product makes the same ones, include/rvh.h): sine of the fract(sin) hash in
 * double precision; last vertex (v = 1) = last curve point; model matrix = identity. */
void orc_expand_strands(const float* strands, int S, int N, int isolines, int divisions,
                        float* pos_width, float* tangent_u);

/* Hair::Hair initial state (Strand.cpp:157-175) from follicle roots + normals. */
void orc_init_strands_reference(int S, int N, const float* roots3, const float* normals3,
                                float* strands);

/* collider decisions per point (test hook, see oracle.c) */
void orc_hit_masks(const orc_params* p, const float* colliders, const float* strands, unsigned char* out);

#ifdef __cplusplus
}
#endif
#endif
