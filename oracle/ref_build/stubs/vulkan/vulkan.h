/*
 * Minimal stand-in for <vulkan/vulkan.h> -- this image has no Vulkan SDK.
 * It declares only the handle types, enums and structs that the reference's HEADERS
 * (Strand.h, Scene.h, Model.h, Vertex.h, Device.h, SwapChain.h, BufferUtils.h) mention,
 * so that the reference's own Strand.cpp / Scene.h can be compiled unmodified into
 * oracle/_ref/ as a checker.  Nothing here does anything; no Vulkan call is ever made.
 * Written for this repo; not derived from the Khronos header.
 */
#pragma once
#include <cstdint>
#include <cstddef>
#define VK_NULL_HANDLE nullptr
#define RVH_STUB_HANDLE(name) typedef struct name##_T* name
RVH_STUB_HANDLE(VkInstance); RVH_STUB_HANDLE(VkPhysicalDevice); RVH_STUB_HANDLE(VkDevice);
RVH_STUB_HANDLE(VkQueue); RVH_STUB_HANDLE(VkCommandPool); RVH_STUB_HANDLE(VkBuffer);
RVH_STUB_HANDLE(VkDeviceMemory); RVH_STUB_HANDLE(VkImage); RVH_STUB_HANDLE(VkImageView);
RVH_STUB_HANDLE(VkSampler); RVH_STUB_HANDLE(VkSurfaceKHR); RVH_STUB_HANDLE(VkSwapchainKHR);
RVH_STUB_HANDLE(VkSemaphore);
typedef uint64_t VkDeviceSize;
typedef uint32_t VkFlags;
typedef VkFlags VkBufferUsageFlags;
typedef VkFlags VkMemoryPropertyFlags;
struct VkAllocationCallbacks;
enum VkFormat { VK_FORMAT_UNDEFINED = 0, VK_FORMAT_R32G32_SFLOAT = 103, VK_FORMAT_R32G32B32_SFLOAT = 106, VK_FORMAT_R32G32B32A32_SFLOAT = 109 };
enum VkVertexInputRate { VK_VERTEX_INPUT_RATE_VERTEX = 0, VK_VERTEX_INPUT_RATE_INSTANCE = 1 };
enum { VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT = 0x10, VK_BUFFER_USAGE_STORAGE_BUFFER_BIT = 0x20,
       VK_BUFFER_USAGE_VERTEX_BUFFER_BIT = 0x80, VK_BUFFER_USAGE_INDIRECT_BUFFER_BIT = 0x100 };
enum { VK_MEMORY_PROPERTY_HOST_VISIBLE_BIT = 0x2, VK_MEMORY_PROPERTY_HOST_COHERENT_BIT = 0x4 };
struct VkExtent2D { uint32_t width, height; };
struct VkVertexInputBindingDescription { uint32_t binding, stride; VkVertexInputRate inputRate; };
struct VkVertexInputAttributeDescription { uint32_t location, binding; VkFormat format; uint32_t offset; };
/* called by Hair::~Hair (Strand.cpp:213-222); defined as no-ops in ref_harness.cpp */
void vkDestroyBuffer(VkDevice, VkBuffer, const VkAllocationCallbacks*);
void vkFreeMemory(VkDevice, VkDeviceMemory, const VkAllocationCallbacks*);

/* The reference's Strand.h:32,38,45 use offsetof() with a run-time array index, which MSVC
 * (the authors' compiler) accepts and g++'s __builtin_offsetof rejects.  Use the classic
 * null-pointer form MSVC's own <cstddef> expands to, for these reference headers only. */
#include <array>
#include <vector>
#include <string>
#include <iostream>
#include <chrono>
#include <bitset>
#include <unordered_map>
#include <functional>
#include <cmath>
#include <cstring>
#include <cstdlib>
#define GLM_ENABLE_EXPERIMENTAL
#include <glm/glm.hpp>
#include <glm/gtx/transform.hpp>
#include <glm/gtx/hash.hpp>
#undef offsetof
#define offsetof(type, member) ((size_t)&reinterpret_cast<char const volatile&>((((type*)0)->member)))

/* MSVC lets `friend class Instance;` (Device.h:8) introduce the name; g++ does not. */
class Instance;
class Device;
class SwapChain;
