/*
 * ref_compute_tu.cpp -- runs the reference's OWN compute shader text on the CPU.
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * src/shaders/compute.comp is GLSL; this image has no Vulkan/lavapipe/glslang.  GLSL's
 * vector language is what the reference's vendored glm 0.9.9.0 mirrors, so
 * build_ref.py applies a short list of purely textual substitutions to the shader
 * (listed in build_ref.py; no statement of main() is added, removed or reordered) and
 * writes the result to oracle/_ref/gen/compute_comp.gen.inc, which is #included below
 * and compiled against that glm.  What this file adds is the execution model:
 *   - one invocation per strand, each on its own ucontext fiber;
 *   - barrier() yields to the scheduler, which resumes every invocation in turn, so a
 *     barrier is GLOBAL across workgroups (the "intended" semantics);
 *   - atomicAdd is a plain add (fibers are cooperative, one OS thread);
 *   - vkCmdFillBuffer(grid, 0) before the dispatch (Renderer.cpp:2063).
 *   - an optional barrier hook (ref_set_barrier_hook) is called once every invocation has reached barrier k (k = 1, 2, 3
 *     in compute.comp order: :130, :208, :255) with the grid buffer, so that several PROCESSES -- one per host core, each
 *     dispatching its own strand range of one head -- can sum their grids before the gather: bench.py --impl reference;
 *   - fiber stacks live in a grow-only arena, so a steady-state dispatch pays no page faults for them.
 * Invocations: exactly S by default.  The shader has no `idx < S` guard and the
 * reference dispatches 32*ceil((S+31)/32) (Renderer.cpp:2070); emulate_oob=1 runs those
 * extra invocations on a zero-filled tail, which is one possible outcome of that
 * out-of-bounds access, to show its effect on the grid / vertexCount.
 */
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <ucontext.h>
#include <cstring>
#include <cstdlib>
#include <cstdint>
#include <vector>

using namespace glm;

static uvec3 gl_GlobalInvocationID;
static void barrier();
static inline int atomicAdd(int& mem, int v) { int old = mem; mem += v; return old; }
static inline uint atomicAdd(uint& mem, uint v) { uint old = mem; mem += v; return old; }

#include "compute_comp.gen.inc"

/* ---- fiber scheduler ---------------------------------------------------------------- */
static ucontext_t g_sched;
static ucontext_t* g_current = nullptr;
static bool g_done_flag = false;

static void (*g_barrier_hook)(int, void*, size_t) = nullptr;
static char* g_arena = nullptr;
static size_t g_arena_bytes = 0;

static void barrier() { swapcontext(g_current, &g_sched); }

static void fiber_entry() {
    shader_main();
    g_done_flag = true;
    /* uc_link returns to the scheduler */
}

extern "C" {

int ref_shader_num_curve_points(void) { return NUM_CURVE_POINTS; }
int ref_shader_grid_dim(void) { return GRID_DIM; }
void ref_set_barrier_hook(void (*fn)(int, void*, size_t)) { g_barrier_hook = fn; }

/* strands: Strand[S] (in/out); colliders: Collider[NUM_COLLIDERS]; grid_out: GridCell[GRID_DIM^3]
 * (int32 x4) after the dispatch; indirect_out: {vertexCount, instanceCount, firstVertex, firstInstance}. */
int ref_compute_dispatch(int S, float* strands, const float* colliders48, float dt, float total_time,
                         int32_t* grid_out, uint32_t* indirect_out, int emulate_oob) {
    const int inv = emulate_oob ? ((S + 31) / 32) * 32 : S;
    std::vector<Strand> buf((size_t)inv);
    std::memset((void*)buf.data(), 0, sizeof(Strand) * (size_t)inv);
    std::memcpy((void*)buf.data(), strands, sizeof(Strand) * (size_t)S);
    inStrands = buf.data();
    std::memcpy((void*)colliders, colliders48, sizeof(Collider) * NUM_COLLIDERS);
    deltaTime = dt;
    totalTime = total_time;
    std::memset((void*)&grid, 0, sizeof(grid));              /* vkCmdFillBuffer, Renderer.cpp:2063 */
    numStrands.vertexCount = (uint)S; numStrands.instanceCount = 1; numStrands.firstVertex = 0; numStrands.firstInstance = 0;

    const size_t stack_bytes = 32 * 1024 + 8 * sizeof(Strand);
    std::vector<ucontext_t> ctx((size_t)inv);
    if (g_arena_bytes < stack_bytes * (size_t)inv) {
        std::free(g_arena);
        g_arena_bytes = stack_bytes * (size_t)inv;
        g_arena = (char*)std::malloc(g_arena_bytes);
        if (!g_arena) { g_arena_bytes = 0; return -1; }
    }
    char* stacks = g_arena;
    std::vector<char> done((size_t)inv, 0);
    for (int i = 0; i < inv; ++i) {
        getcontext(&ctx[i]);
        ctx[i].uc_stack.ss_sp = stacks + stack_bytes * (size_t)i;
        ctx[i].uc_stack.ss_size = stack_bytes;
        ctx[i].uc_link = &g_sched;
        makecontext(&ctx[i], fiber_entry, 0);
    }
    int remaining = inv, pass = 0;
    while (remaining > 0) {
        for (int i = 0; i < inv; ++i) {
            if (done[i]) continue;
            gl_GlobalInvocationID = uvec3((uint)i, 0u, 0u);
            g_current = &ctx[i];
            g_done_flag = false;
            swapcontext(&g_sched, &ctx[i]);
            if (g_done_flag) { done[i] = 1; --remaining; }
        }
        ++pass;                                               /* every live invocation now waits at barrier number `pass` */
        if (remaining > 0 && g_barrier_hook) g_barrier_hook(pass, (void*)&grid, sizeof(grid));
    }
    std::memcpy(strands, buf.data(), sizeof(Strand) * (size_t)S);
    if (grid_out) std::memcpy(grid_out, &grid, sizeof(grid));
    if (indirect_out) std::memcpy(indirect_out, &numStrands, 16);
    inStrands = nullptr;
    return inv;
}

} /* extern "C" */
