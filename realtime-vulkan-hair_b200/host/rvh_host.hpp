// rvh_host.hpp -- host-side mirror of the reference's Hair / Scene / Renderer classes for the
// ONE path this repository replaces: the per-frame compute pass (src/shaders/compute.comp).
//
// The reference has no plugin seam: physics is reached through Vulkan objects owned by these three
// classes.  This mirror keeps their public signatures, struct byte layouts and call order
// (main.cpp:226-251 scene assembly, main.cpp:281-283 `UpdateTime -> Frame -> moveSphere`), and
// routes what used to be descriptor sets + vkQueueSubmit(compute) through the C ABI in
// include/rvh.h.  Everything Vulkan-only (raster pipelines, swap chain, textures) is out of scope
// and appears here only as opaque pointer types so the constructor signatures stay intact.
//
//   reference                                         here
//   Strand.h:11-15   struct Strand                    rvh_host::Strand (static_assert 480 B at N=10)
//   Strand.h:53-58   struct StrandDrawIndirect        rvh_host::StrandDrawIndirect
//   Scene.h:16-19    struct Time                      rvh_host::Time
//   Scene.h:23-39    struct Collider                  rvh_host::Collider  (rvh_collider_build)
//   Scene.h:42-49    struct GridCell                  rvh_host::GridCell
//   Strand.h:61-80   class Hair                       rvh_host::Hair      (own OBJ reader, same sampling)
//   Scene.h:52-105   class Scene                      rvh_host::Scene
//   Renderer.h:11-66 class Renderer (compute half)    rvh_host::Renderer  (rvh_create / rvh_step)
//
// Errors: the reference throws std::runtime_error on every Vulkan failure (e.g. Renderer.cpp:2317-2319);
// so does this layer on any non-zero rvh_status.  OBJ load failure is exit(1) in the reference
// (Strand.cpp:45-47); here it is a std::runtime_error (a library must not end the process).
#pragma once
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rvh.h"

#ifdef RVH_HOST_WITH_VULKAN
#include <vulkan/vulkan.h>
#else
typedef struct VkBuffer_T* VkBuffer;                 // opaque, never dereferenced in the headless build
typedef struct VkCommandPool_T* VkCommandPool;
#endif
class Device;      // reference classes that only travel through constructors here
class SwapChain;
class Camera;
class Model;

namespace rvh_host {

constexpr unsigned int NUM_STRANDS = 900;            // Strand.h:8
constexpr unsigned int NUM_CURVE_POINTS = 10;        // Strand.h:9

struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };
struct ivec3 { int x, y, z; };
struct mat4 { float m[16]; };                        // column-major, like glm::mat4

struct Strand {                                      // Strand.h:11-15 at the reference's N = 10
    vec4 curvePoints[NUM_CURVE_POINTS];
    vec4 curveVels[NUM_CURVE_POINTS];
    vec4 correctionVecs[NUM_CURVE_POINTS];
};
static_assert(sizeof(Strand) == 480, "Strand must stay 48*N bytes (vertex-buffer stride, Strand.h:20)");

struct StrandDrawIndirect { uint32_t vertexCount, instanceCount, firstVertex, firstInstance; };
static_assert(sizeof(StrandDrawIndirect) == 16, "vkCmdDrawIndirect argument block");

struct Time { float deltaTime = 0.0f; float totalTime = 0.0f; };

struct Collider {                                    // Scene.h:23-39
    mat4 transform, inv, invTrans;
    Collider(vec3 trans, vec3 rot, vec3 scale);      // rotation in degrees, transform = T*Rz*Ry*Rx*S
};
static_assert(sizeof(Collider) == 192, "Collider is three mat4 (std140 UBO element, compute.comp:25-33)");

struct GridCell {                                    // Scene.h:42-49
    ivec3 velocity; int density;
    GridCell(ivec3 vel, int density) : velocity(vel), density(density) {}
};
static_assert(sizeof(GridCell) == 16, "GridCell is ivec3 + int");

// Follicle sampling of Strand.cpp:26-146: triangulated OBJ, srand(8), triangle = rand() % T,
// (u, v) folded into the triangle, normal of the triangle's first corner.
int GeneratePointsOnMesh(const std::string& filename, int numStrands, std::vector<vec3>& points, std::vector<vec3>& pointNormals);

class Hair {
public:
    Hair(Device* device, VkCommandPool commandPool, std::string objFilename);                      // Strand.h:74
    Hair(Device* device, VkCommandPool commandPool, std::string objFilename, int numStrands, int numCurvePoints);
    // Pre-built state (synthetic heads): Strand[S] AoS, float [S][3][N][4].
    Hair(Device* device, VkCommandPool commandPool, std::vector<float> strandsAos, int numStrands, int numCurvePoints);
    VkBuffer GetStrandsBuffer() const { return strandsBuffer; }        // exported VkBuffer in the interop build
    VkBuffer GetNumStrandsBuffer() const { return numStrandsBuffer; }
    VkBuffer GetModelBuffer() const { return modelBuffer; }
    int GetNumStrands() const { return numStrands; }
    int GetNumCurvePoints() const { return numCurvePoints; }
    const std::vector<float>& GetInitialStrands() const { return strands; }   // what Hair::Hair uploads (Strand.cpp:188)
    const StrandDrawIndirect& GetIndirectDraw() const { return indirectDraw; }
    // Interop build: the caller exports strandsBuffer's memory (VK_KHR_external_memory_fd) and hands the fd over.
    void SetExportedStrandsMemory(VkBuffer buffer, int fd, size_t bytes) { strandsBuffer = buffer; exportedFd = fd; exportedBytes = bytes; }
    // ... likewise the indirect-args buffer (Renderer.cpp:2240-2254 barriers it before vkCmdDrawIndirect) and a binary semaphore
    // (VK_KHR_external_semaphore_fd) that the compute step signals and the graphics submit waits on: the compute -> graphics
    // ordering the reference leaves out (Renderer.cpp:2311-2345 submits both queues with no semaphore in between)
    void SetExportedIndirectMemory(VkBuffer buffer, int fd, size_t bytes) { numStrandsBuffer = buffer; exportedIndirectFd = fd; exportedIndirectBytes = bytes; }
    void SetExportedSemaphore(int fd) { exportedSemaphoreFd = fd; }
    int exportedFd = -1; size_t exportedBytes = 0;
    int exportedIndirectFd = -1; size_t exportedIndirectBytes = 0;
    int exportedSemaphoreFd = -1;
private:
    void buildFromFollicles(const std::vector<vec3>& roots, const std::vector<vec3>& normals);
    VkBuffer strandsBuffer = nullptr, numStrandsBuffer = nullptr, modelBuffer = nullptr;
    int numStrands = 0, numCurvePoints = 0;
    std::vector<float> strands;
    StrandDrawIndirect indirectDraw{};
};

class Scene {
public:
    Scene() = delete;
    Scene(Device* device, VkCommandPool commandPool, std::vector<Collider> colliders, std::vector<Model*> models);   // Scene.h:83
    const std::vector<Model*>& GetModels() const { return models; }                // Scene.h:88
    const std::vector<Hair*>& GetHair() const { return hair; }
    const std::vector<Collider>& GetColliders() const { return colliders; }
    // Scene.h:91: the host copy of the grid the reference uploads once, all zero (Scene.cpp:16-20); the live grid is device memory
    const std::vector<GridCell>& GetGrid() const { return grid; }
    void AddModel(Model* m) { models.push_back(m); }                                // Scene.h:94
    void AddHair(Hair* h) { hair.push_back(h); }
    void AddCollider(Collider c) { colliders.push_back(c); }
    // Scene.h:98-101.  In the reference these are the VkBuffers behind descriptor sets 1-3 of the compute pipeline
    // (Renderer.cpp:836-997).  Here the time and collider UBOs are arguments of rvh_step / rvh_set_colliders and the grid is
    // owned by the rvh context, so the handles stay VK_NULL_HANDLE unless an interop build registers its own (SetVulkanBuffers);
    // nothing in the compute path reads them.
    VkBuffer GetTimeBuffer() const { return timeBuffer; }
    VkBuffer GetCollidersBuffer() const { return collidersBuffer; }
    VkBuffer GetGridBuffer() const { return gridBuffer; }
    VkBuffer GetModelBuffer() const { return modelBuffer; }
    void SetVulkanBuffers(VkBuffer time, VkBuffer colliders, VkBuffer grid, VkBuffer model) { timeBuffer = time; collidersBuffer = colliders; gridBuffer = grid; modelBuffer = model; }
    void UpdateTime();                                  // Scene.cpp:78-87: wall clock unless a fixed step is set
    void translateSphere(vec3 translation);             // Scene.cpp:110-136 (collider half)
    const Time& GetTime() const { return time; }
    // Harness extension: deterministic dt instead of the wall clock (the reference's dt is not reproducible).
    void SetFixedDeltaTime(float dt) { fixedDt = dt; }
private:
    Device* device;
    Time time;
    std::vector<Model*> models;
    std::vector<Hair*> hair;
    std::vector<Collider> colliders;
    std::vector<GridCell> grid;
    VkBuffer timeBuffer = nullptr, collidersBuffer = nullptr, gridBuffer = nullptr, modelBuffer = nullptr;
    float fixedDt = 0.0f;
    std::chrono::high_resolution_clock::time_point startTime = std::chrono::high_resolution_clock::now();
};

class Renderer {
public:
    Renderer() = delete;
    // Renderer.h:14.  Does what the compute half of the reference constructor does (Renderer.cpp:18-68):
    // descriptor-set layouts/sets for time, colliders, grid, strands (402-497, 836-997), the compute pipeline
    // (1748-1788) and the pre-recorded command buffer (2022-2077) become one rvh context per Hair.
    Renderer(Device* device, SwapChain* swapChain, Scene* scene, Camera* camera, Camera* shadowCamera);
    Renderer(Device* device, SwapChain* swapChain, Scene* scene, Camera* camera, Camera* shadowCamera, int flags, int cudaDevice);
    ~Renderer();
    Scene* scene;
    void CreateComputePipeline();            // Renderer.cpp:1748-1788 -> rvh_create + upload
    void RecordComputeCommandBuffer();       // Renderer.cpp:2022-2077 -> nothing to pre-record; validates state
    void Frame();                            // Renderer.cpp:2309-2350, compute submit only (2311-2319)
    // Harness extensions (the reference never reads simulation state back):
    void WaitIdle();
    void DownloadStrands(size_t hairIndex, std::vector<float>& out);
    StrandDrawIndirect ReadIndirectDraw(size_t hairIndex);
    rvh_ctx* GetContext(size_t hairIndex) const { return contexts.at(hairIndex); }
private:
    void check(int status, rvh_ctx* ctx, const char* what) const;
    Device* device; SwapChain* swapChain; Camera* camera; Camera* shadowCamera;
    int flags, cudaDevice;
    std::vector<rvh_ctx*> contexts;          // one per Hair (the reference dispatches once per Hair, Renderer.cpp:2065-2071)
};

}  // namespace rvh_host
