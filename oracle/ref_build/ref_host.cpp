/*
 * ref_host.cpp -- link-time stubs + C entry points around the reference's OWN host
 * sources, compiled where they lie under /root/reference (never copied):
 *   - src/Strand.cpp   (Hair::Hair -> GeneratePointsOnMesh: follicles + initial state)
 *   - src/Scene.h      (Collider ctor: T*Rz*Ry*Rx*S with the vendored glm 0.9.9.0)
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  Output goes to oracle/_ref/.
 *
 * The Vulkan side of those classes is stubbed out: Model's ctor/dtor and
 * BufferUtils::CreateBufferFromData are defined here so the link closes, and the
 * "upload" simply captures the bytes Hair::Hair would have sent to the GPU.
 */
#include <cstring>
#include <vector>
#include <cstdlib>
#include "Scene.h"       /* reference header: Collider, Time, GridCell */
#include "BufferUtils.h" /* reference header */

static std::vector<std::vector<unsigned char>> g_uploads;

Model::Model(Device* device, VkCommandPool, const std::vector<Vertex>& v, const std::vector<uint32_t>& i, glm::mat4)
    : device(device), vertices(v), indices(i) {}
Model::~Model() {}
VkDevice Device::GetVkDevice() { return nullptr; }
void vkDestroyBuffer(VkDevice, VkBuffer, const VkAllocationCallbacks*) {}
void vkFreeMemory(VkDevice, VkDeviceMemory, const VkAllocationCallbacks*) {}

void BufferUtils::CreateBufferFromData(Device*, VkCommandPool, void* data, VkDeviceSize size, VkBufferUsageFlags,
                                       VkBuffer& buffer, VkDeviceMemory& memory) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    g_uploads.emplace_back(p, p + size);
    buffer = nullptr;
    memory = nullptr;
}

extern "C" {

int ref_num_strands(void) { return (int)NUM_STRANDS; }
int ref_num_curve_points(void) { return (int)NUM_CURVE_POINTS; }
int ref_sizeof_strand(void) { return (int)sizeof(Strand); }
int ref_sizeof_collider(void) { return (int)sizeof(Collider); }
int ref_sizeof_gridcell(void) { return (int)sizeof(GridCell); }
int ref_sizeof_indirect(void) { return (int)sizeof(StrandDrawIndirect); }

/* Runs the reference's Hair::Hair(device, pool, objFilename) (Strand.cpp:149-191) and
 * returns what it uploaded: Strand[NUM_STRANDS] and the StrandDrawIndirect. */
int ref_hair_init(const char* obj_path, float* strands_out, size_t strands_bytes, uint32_t indirect_out[4]) {
    g_uploads.clear();
    Hair* hair = new Hair(nullptr, nullptr, std::string(obj_path));
    int n = hair->GetNumStrands();
    if (g_uploads.size() < 2 || g_uploads[0].size() != strands_bytes) { delete hair; return -1; }
    std::memcpy(strands_out, g_uploads[0].data(), strands_bytes);
    std::memcpy(indirect_out, g_uploads[1].data(), 16);
    delete hair;
    return n;
}

/* Collider(trans, rot, scale) (Scene.h:28-38) -> 48 floats (transform, inv, invTrans). */
void ref_collider_build(const float t[3], const float r[3], const float s[3], float out48[48]) {
    Collider c(glm::vec3(t[0], t[1], t[2]), glm::vec3(r[0], r[1], r[2]), glm::vec3(s[0], s[1], s[2]));
    std::memcpy(out48, &c, sizeof(Collider));
}

/* The collider half of Scene::translateSphere (Scene.cpp:112-119), same glm calls. */
void ref_collider_translate(float c48[48], const float tr[3]) {
    Collider c(glm::vec3(0), glm::vec3(0), glm::vec3(1));
    std::memcpy(&c, c48, sizeof(Collider));
    glm::mat4 currTransform = c.transform;
    glm::mat4 newTransform = glm::translate(currTransform, glm::vec3(tr[0], tr[1], tr[2]));
    c.transform = newTransform;
    glm::mat4 inverse = glm::inverse(newTransform);
    c.inv = inverse;
    c.invTrans = glm::transpose(inverse);
    std::memcpy(c48, &c, sizeof(Collider));
}

} /* extern "C" */
