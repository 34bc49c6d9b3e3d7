/*
 * rvh.h -- C ABI of the B200-native guide-strand physics step (librvh.so).
 *
 * Drop-in boundary for ONE path of clach/Realtime-Vulkan-Hair: the per-frame
 * Follow-the-Leader / PBD compute pass (src/shaders/compute.comp).  The reference has no
 * FFI seam; its physics is reached only through Vulkan objects.  Each entry point below
 * names the reference call site it replaces (file:line relative to the reference tree).
 * The C++ host mirror that keeps the reference's Hair / Scene / Renderer signatures on
 * top of this ABI lives in realtime-vulkan-hair_b200/host/; the binding a maintainer
 * would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every function returns 0 (RVH_OK) or a negative rvh_status;
 *     nothing throws across the boundary (reference convention is
 *     `throw std::runtime_error`, e.g. Renderer.cpp:2317-2319 -- the C++ mirror converts).
 *   - the caller owns every host pointer; data is copied during the call.
 *   - one context = one CUDA device + one stream; calls on a context are not re-entrant
 *     (the reference is single-threaded with one frame in flight, main.cpp:257-285).
 *   - there is NO CPU fallback: without a CUDA device rvh_create fails.
 *
 * Buffer layouts (bit-for-bit the reference's):
 *   Strand[S]           float [S][3][N][4]: curvePoints, curveVels, correctionVecs
 *                       (Strand.h:11-15; compute.comp:44-48), N generalised from 10
 *   Collider[n]         3 column-major mat4: transform, inv, invTrans = 192 B
 *                       (Scene.h:23-26; compute.comp:25-33); index 0 is the sphere
 *   StrandDrawIndirect  4 x uint32 (Strand.h:53-58)
 *   grid download       int64 [G^3][4] = (vel.x, vel.y, vel.z, density), fixed point
 *                       x grid_scale; with RVH_GRID_INT32_WRAP: int32 [G^3][4], the
 *                       reference's GridCell (Scene.h:42-49; compute.comp:35-42)
 */
#ifndef RVH_H
#define RVH_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rvh_ctx rvh_ctx;

typedef enum {
    RVH_OK = 0,
    RVH_ERR_INVALID = -1,   /* bad argument / size mismatch                 */
    RVH_ERR_CUDA = -2,      /* CUDA runtime error (message in last_error)   */
    RVH_ERR_NCCL = -3,      /* NCCL not loadable or a collective failed     */
    RVH_ERR_STATE = -4      /* call order (e.g. step before upload)         */
} rvh_status;

enum {
    RVH_GRID_ON         = 1,   /* hair-hair friction through the voxel grid, compute.comp:211-298 */
    RVH_WIND_A          = 2,   /* compute.comp:151 (commented out in the reference)              */
    RVH_WIND_B          = 4,   /* compute.comp:152 (commented out in the reference)              */
    RVH_GRID_INT32_WRAP = 8,   /* read the grid back through its low 32 bits = reference int32   */
    RVH_KEEP_CORRECTION = 16,  /* also store correctionVecs (dead across steps; download only)   */
    RVH_KEEP_ORDER      = 32,  /* no internal Morton reordering of strands                       */
    /* North-star extensions that do NOT exist in the reference (SURVEY.md top table, section 8 f3); default off,
     * each with its own oracle mode (oracle/oracle.c ORC_SDF_ON / ORC_REPULSION_ON). */
    RVH_SDF_ON          = 64,  /* colliders 1..n (the analytic ellipsoids of compute.comp:170-179) are replaced by a
                                  sampled signed-distance volume of the head (rvh_set_head_sdf / rvh_bake_head_sdf_*);
                                  collider 0, the movable sphere, stays analytic                     */
    RVH_REPULSION_ON    = 128, /* hair-hair repulsion: v -= repulsion * h*grad(rho)/sum(D) from the same voxel grid (each component bounded by `repulsion`),
                                  applied with the friction gather (needs RVH_GRID_ON)              */
    RVH_SDF_TMA         = 256  /* stage the SDF through TMA (warp-private 8x4x4-node shared-memory tiles, one row ahead)
                                  instead of plain cached loads.  Bit-identical results.  Measured on B200 the staging
                                  costs more than it saves (L1 already captures the reuse between neighbouring strands;
                                  DESIGN.md section 8), so plain loads are the default and this is the A/B switch.  */
};

typedef struct {              /* every field defaults to the reference constant     */
    int   device;             /* CUDA ordinal                                       */
    int   num_strands;        /* S (strands owned by this context)                  */
    int   num_points;         /* N >= 2, root included      compute.comp:5          */
    float rest_length;        /* "radius" 2.5f/(N-1)        compute.comp:139-140    */
    float gravity_y;          /* -9.8f                      compute.comp:150        */
    float damping;            /* 0.998f                     compute.comp:7          */
    float vmax;               /* 10.f                       compute.comp:198        */
    float penalty_k;          /* 1900.f                     compute.comp:163,173    */
    float sphere_radius;      /* 1.f                        compute.comp:161        */
    int   grid_dim;           /* 64                         compute.comp:9          */
    float grid_extent;        /* 7.f                        compute.comp:10         */
    float grid_origin[3];     /* -3,-2,-5                   compute.comp:206        */
    float grid_scale;         /* 1e6f                       compute.comp:11         */
    float friction;           /* 0.08f                      compute.comp:296        */
    int   flags;              /* RVH_* bits; default RVH_GRID_ON                    */
    int   strands_per_thread; /* 0 = auto; 1 or 2 (tuning, results identical; 4 = 2) */
    float repulsion;          /* 0.2f (velocity units); RVH_REPULSION_ON only (extension)       */
} rvh_config;

/* Fill cfg with the reference constants for S strands of N points. */
void rvh_default_config(rvh_config* cfg, int num_strands, int num_points);

/* Replaces Renderer::Create{Time,Colliders,Grid,Compute}DescriptorSetLayout
 * (Renderer.cpp:402-497), the matching descriptor sets (836-997), CreateComputePipeline
 * (1748-1788) and RecordComputeCommandBuffer (2022-2077); allocates the grid that
 * Scene::Scene uploads (Scene.cpp:16-20). */
int rvh_create(rvh_ctx** out, const rvh_config* cfg);

/* Multi-GPU: this context owns one contiguous shard of the strands; only the voxel grid is
 * exchanged, once per step.  With one process per GPU on an NVLink/NVSwitch box the exchange is
 * ONE fused kernel per rank over CUDA-IPC peer memory (pull-reduce the rank's cell slice from all
 * peers, finalize, push the float cells to all peers); otherwise an in-place ncclAllReduce of the
 * int64 accumulators.  Both give bit-identical results for any rank count (integer sums).
 * nccl_unique_id is the 128-byte ncclUniqueId from rvh_nccl_unique_id() on rank 0, distributed by
 * the launcher; NCCL is also what carries the IPC handles at creation.  Every rank must call
 * rvh_step (and rvh_download_grid, which all-reduces on demand) in lockstep. */
int rvh_nccl_unique_id(void* out128);
int rvh_create_sharded(rvh_ctx** out, const rvh_config* cfg, int rank, int nranks, const void* nccl_unique_id);
int rvh_exchange_mode(rvh_ctx* ctx);   /* 0 = single rank, 1 = NCCL all-reduce, 2 = fused peer-memory exchange */

/* Collider UBO write: Scene::Scene memcpy (Scene.cpp:10-13) and Scene::translateSphere
 * (Scene.cpp:133).  colliders = n x 192 bytes, n <= 8. */
int rvh_set_colliders(rvh_ctx* ctx, const void* colliders, int n);

/* Hair::Hair's strands upload (Strand.cpp:188).  bytes must be S*48*N. */
int rvh_upload_strands_aos(rvh_ctx* ctx, const void* strands, size_t bytes);

/* GPU scene init (no host traffic): the seeded synthetic head of the benchmarks -- roots on the top hemisphere
 * of collider 1 (the head ellipsoid, main.cpp:231), counter-based splitmix64 keyed by first_strand + local
 * index so any rank generates exactly its shard, points at rest spacing strand_length/(N-1), velocity (0,0,-1)
 * as Strand.cpp:166.  Host twin: realtime-vulkan-hair_b200/scenes.py synthetic_head.  Needs rvh_set_colliders first. */
int rvh_init_synthetic_head(rvh_ctx* ctx, unsigned long long first_strand, float strand_length, unsigned long long seed);

/* GPU scene init from a triangle mesh: follicles drawn on the device with the area weighting
 * that Strand.cpp:92 leaves as a TODO (triangle from the area CDF, then the reference's folded barycentric sample,
 * Strand.cpp:96-110), counter-based RNG keyed by first_strand + local index so every rank generates exactly its shard.
 * tri_pos / tri_nrm = ntris x 3 corners x 3 floats (tri_nrm may be NULL: face normals).  Strands leave the surface at rest
 * spacing strand_length/(N-1).  Host twin: realtime-vulkan-hair_b200/scenes.py mesh_head.  (The reference's own uniform
 * srand(8) placement is reproduced bit for bit by the host mirror, rvh_host::Hair.) */
int rvh_init_from_mesh(rvh_ctx* ctx, const float* tri_pos, const float* tri_nrm, int ntris, unsigned long long first_strand,
                       float strand_length, unsigned long long seed);

/* ---- head SDF collision (north-star extension; the reference has analytic ellipsoids only) ----------------
 * The volume is a dense float array of signed distances at the NODES of a regular lattice, node (i,j,k) at
 * origin + cell*(i,j,k), stored x-fastest [nz][ny][nx], negative inside.  With RVH_SDF_ON a point x whose lattice
 * cell lies inside the volume and whose trilinearly interpolated distance d is < 0 receives the penalty
 * penalty_k * (-d) * normalize(grad d) in place of the ellipsoid terms of compute.comp:170-179 (hit counting and
 * the division by the number of colliders hit, compute.comp:182-184, are unchanged; the sphere stays analytic).
 * With RVH_SDF_TMA k_ftl_step reads the volume through TMA: per row of a warp's 32-64 neighbouring strands one
 * 8x4x4-node box is staged in shared memory by cp.async.bulk.tensor.3d one row ahead of its use. */
int rvh_set_head_sdf(rvh_ctx* ctx, const float* sdf, const int dim[3], const float origin[3], float cell);
/* GPU bake from the ellipsoid colliders 1..n currently set: signed radial distance |x - T*normalize(inv*x)|, the
 * penetration depth compute.comp:171-172 uses, negative inside, min over the ellipsoids. */
int rvh_bake_head_sdf_from_colliders(rvh_ctx* ctx, const int dim[3], const float origin[3], float cell);
/* GPU bake from a triangle mesh (e.g. models/mannequin.obj, main.cpp:222): exact distance to the closest triangle,
 * sign from the generalized winding number (robust to the open neck of the mesh).  verts = nverts x 3 floats,
 * tris = ntris x 3 vertex indices. */
int rvh_bake_head_sdf_from_mesh(rvh_ctx* ctx, const float* verts, int nverts, const int* tris, int ntris,
                                const int dim[3], const float origin[3], float cell);
int rvh_download_head_sdf(rvh_ctx* ctx, float* sdf, size_t bytes);     /* nx*ny*nz floats */
/* Test hook: the collider candidate mask k_ftl_step consults in steady-state stepping: one byte per box
 * of 2x2x2 grid cells, x fastest, bit j set when ellipsoid j (collider j+1) can contain a point of the box.  *dim receives
 * grid_dim/2, or 0 when no mask is in use (then nothing is written). */
int rvh_download_collider_mask(rvh_ctx* ctx, unsigned char* out, size_t bytes, int* dim);
int rvh_sdf_mode(rvh_ctx* ctx);   /* 0 = no volume, 1 = plain loads (default), 2 = TMA-staged tiles (RVH_SDF_TMA) */

/* ---- guide strands -> render strands -------------------------------------------------
 * What the reference's tessellator does with the simulated buffer: every guide strand is one patch, expanded into
 * `isolines` x `divisions` line segments (12 x 42, hair.tesc:19-20) whose vertices hair.tese places by Bezier
 * interpolation along the guide plus a sideways deviation (hair.tese:34-80, 226-304).  This produces the same
 * vertices headless, as line strips: two float4 per vertex, arrays [S][isolines][divisions+1] in the caller's strand
 * order -- pos_width = (position, strand width hair.tese:313-315), tangent_u = (unit tangent of the guide segment
 * hair.tese:267, isoline coordinate u).  Host pointers may be NULL (results stay on the device,
 * rvh_expand_device_buffers); ms_out, if given, receives the kernel time (CUDA events).  Deviations: the fract(sin)
 * hashes take the sine in double precision so that CPU and GPU agree; at v = 1 the shader reads one curve point past
 * the end (hair.tese:165-166), here the last vertex is the guide's last point; model matrix = identity as Hair::Hair
 * sets it (Strand.cpp:183-185). */
int rvh_expand_strands(rvh_ctx* ctx, int isolines, int divisions, float* pos_width, float* tangent_u, size_t bytes_each, float* ms_out);
int rvh_expand_device_buffers(rvh_ctx* ctx, void** pos_width, void** tangent_u, size_t* vertices);

/* Vulkan interop: map the exported strands VkBuffer (VK_KHR_external_memory_fd) and
 * keep it updated after every step in the reference's AoS vertex-buffer layout
 * (Renderer.cpp:2153-2161 binds it).  Needs a Vulkan device on the caller's side. */
int rvh_import_strands_fd(rvh_ctx* ctx, int fd, size_t bytes);
/* ... and the indirect-args VkBuffer vkCmdDrawIndirect reads (Hair::GetNumStrandsBuffer; Renderer.cpp:2240-2254 puts a buffer
 * barrier on it): every step writes StrandDrawIndirect {S, 1, 0, 0} into it (the shader's own count, compute.comp:126-130, 302,
 * ends at 32*ceil(S/32): include/rvh.h rvh_draw_indirect). */
int rvh_import_indirect_fd(rvh_ctx* ctx, int fd, size_t bytes);
/* A binary VkSemaphore exported with VK_KHR_external_semaphore_fd: every step signals it on the context's stream after its
 * last write into the imported buffers, and the graphics submit waits on it.  This is the compute -> graphics ordering the
 * reference does not have: Renderer::Frame submits the compute and the graphics command buffers to two queues with no
 * semaphore between them (Renderer.cpp:2311-2345).  Without it the caller must rvh_sync() before the raster submit.
 * The three import entry points need a Vulkan device on the caller's side; this image has none, so only their failure
 * paths and -- through the hook below -- everything behind them are exercised by the tests. */
int rvh_import_semaphore_fd(rvh_ctx* ctx, int fd);
/* Test hook: stand-ins for the imported buffers.  strands_dev / indirect_dev are DEVICE pointers the caller owns (>= S*48*N and
 * 16 bytes; NULL = none): every step then packs Strand[S] and writes the draw arguments into them exactly as it would into
 * imported Vulkan memory. */
int rvh_debug_set_interop_device_buffers(rvh_ctx* ctx, void* strands_dev, size_t strands_bytes, void* indirect_dev);

/* vkQueueSubmit(Compute, computeCommandBuffer) in Renderer::Frame (Renderer.cpp:2311-2319):
 * grid clear + one pass of compute.comp.  dt / total_time are Scene::UpdateTime's values
 * (Scene.cpp:78-87).  Asynchronous on the context's stream. */
int rvh_step(rvh_ctx* ctx, float dt, float total_time);

/* n back-to-back steps, total_time advancing by dt (float additions, exactly as n calls would); if ms_out != NULL the whole
 * batch is timed with CUDA events on the context's stream and the call synchronises.  Same results as n calls of rvh_step,
 * bit for bit; small scenes take a faster route to them: without the grid up to 32 steps ride in ONE kernel launch (the strands
 * never interact; small ones as a wavefront over the steps, k_ftl_wave: several lanes per strand, each a step ahead of the next); with the grid, scenes of up to ~9.5K strands run whole steps (up to 32) inside ONE persistent cooperative launch
 * (k_scene_step: FTL | splat | gather between grid-wide barriers; rvh_step takes the same kernel, one launch per step); larger
 * ones without wind capture the step once as a CUDA graph and replay it (nothing in its kernel parameters depends on time).
 * Per-kernel profiling (rvh_profile_enable) and RVH_NO_FAST_STEP_N=1 select plain stepping (RVH_SCENE_CTAS=0: no k_scene_step). */
int rvh_step_n(rvh_ctx* ctx, int n, float dt, float total_time0, float* ms_out);

/* Host round trip in one call: upload Strand[S], one step, download Strand[S].  Only curvePoints and
 * curveVels cross PCIe (correctionVecs are dead on input; on output they are written only with
 * RVH_KEEP_CORRECTION, otherwise that third of the host buffer is left untouched).
 * From 128K strands up (and without RVH_KEEP_CORRECTION / imported buffers) the call is pipelined over 16 strand chunks on three
 * streams: the upload of chunk c+1 runs beside the kernels of chunk c, and the POSITIONS of chunk c travel back while later
 * chunks are still arriving (the gather changes velocities only); the velocities follow once the grid is complete.  Same
 * bytes in the host buffer as upload + rvh_step + download.  Use pinned host memory, or the copies serialise. */
int rvh_step_host(rvh_ctx* ctx, void* strands_inout, size_t bytes, float dt, float total_time);

/* Read-backs (synchronise).  The reference never reads these back; tests do. */
int rvh_download_strands_aos(rvh_ctx* ctx, void* strands, size_t bytes);
int rvh_download_grid(rvh_ctx* ctx, void* cells, size_t bytes);
int rvh_draw_indirect(rvh_ctx* ctx, uint32_t out[4]);   /* {S,1,0,0}, compute.comp:126-130,302 */

/* Debug/test hooks: run the step split at the shader's barriers.  phase bits:
 * 1 = integrate+FTL+corrected velocity (+ splat and all-reduce when the grid is on),
 * 2 = grid finalize + gather. */
int rvh_step_phases(rvh_ctx* ctx, float dt, float total_time, int phases);

/* Test hook: the collider decisions (compute.comp:162 sphere, :66 ellipsoids; with RVH_SDF_ON bit 1 = head volume) this path
 * takes for the positions it currently holds, one byte per point, [S][N] in the caller's strand order, bit 0 = sphere, bit j =
 * collider j.  The penalty force is continuous across a collider surface, but the division by the number of colliders hit
 * (compute.comp:182-184) is not, so tests count how often a decision differs from the oracle's (tests/test_parity_full_gpu.py). */
int rvh_debug_hit_masks(rvh_ctx* ctx, unsigned char* out, size_t bytes);

/* Per-kernel CUDA-event timing (for the roofline figure).  When enabled every step
 * brackets its kernels with events; rvh_profile_read returns accumulated milliseconds
 * and launch counts since the last call: [0] ftl_step, [1] grid_gather, [2] grid
 * all-reduce, [3] grid clear, [4] grid_splat, [5] grid_finalize. */
int rvh_profile_enable(rvh_ctx* ctx, int on);   /* 0 off, 1 every kernel, 2 only ftl_step and only in every 4th step (events between kernels cost ~3 us each) */
int rvh_profile_read(rvh_ctx* ctx, float ms[6], int launches[6]);

int rvh_sync(rvh_ctx* ctx);
float rvh_last_step_ms(rvh_ctx* ctx);
long long rvh_kernel_launches(rvh_ctx* ctx);            /* kernels launched so far */
const char* rvh_last_error(rvh_ctx* ctx);               /* ctx may be NULL: create errors */
void rvh_destroy(rvh_ctx* ctx);

/* Host-side helpers mirroring reference host code (no GPU needed). */
void  rvh_collider_build(const float trans[3], const float rot_deg[3], const float scale[3],
                         float out48[48]);                       /* Scene.h:28-38      */
void  rvh_collider_translate(float collider48[48], const float translation[3]); /* Scene.cpp:110-120 */
float rvh_wind_fbm(float total_time);                            /* compute.comp:83-121,152 */
int   rvh_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
