// rvh_host_capi.cpp -- C entry points over the C++ host mirror, for the Python test/bench harness
// and as the worked example of how an application drives the mirror (it is main.cpp:226-251 +
// the loop at main.cpp:281-283, headless).
#include "rvh_host.hpp"

#include <cstring>
#include <string>

using namespace rvh_host;

namespace { thread_local std::string g_err; }

extern "C" {

const char* rvhh_last_error(void) { return g_err.c_str(); }

int rvhh_sizeof(int what) {
    switch (what) {
        case 0: return (int)sizeof(Strand);
        case 1: return (int)sizeof(Collider);
        case 2: return (int)sizeof(GridCell);
        case 3: return (int)sizeof(StrandDrawIndirect);
        case 4: return (int)sizeof(Time);
        default: return -1;
    }
}

/* Hair::Hair(device, pool, objFilename[, S, N]) -> the bytes it would upload.  No GPU needed. */
int rvhh_hair_init(const char* obj_path, int S, int N, float* strands_out, size_t bytes, uint32_t indirect_out[4]) {
    try {
        Hair hair(nullptr, nullptr, std::string(obj_path), S, N);
        const std::vector<float>& st = hair.GetInitialStrands();
        if (bytes != st.size() * sizeof(float)) { g_err = "strands_out must be S*48*N bytes"; return -1; }
        std::memcpy(strands_out, st.data(), bytes);
        std::memcpy(indirect_out, &hair.GetIndirectDraw(), 16);
        return hair.GetNumStrands();
    } catch (const std::exception& e) { g_err = e.what(); return -2; }
}

/* The reference application, headless: scene of main.cpp:226-251 (follicles from obj_path, the six
 * colliders), then `frames` iterations of  scene->UpdateTime(); renderer->Frame(); moveSphere()
 * (main.cpp:281-283) with a fixed dt.  sphere_moves: [frames][3] translations applied AFTER each
 * frame like moveSphere (main.cpp:133-157), or NULL.  If strands_in != NULL it replaces the OBJ
 * initial state (synthetic heads).  Needs a GPU. */
int rvhh_run_scene(const char* obj_path, const float* strands_in, int S, int N, int flags, int frames, float fixed_dt,
                   const float* sphere_moves, float* strands_out, size_t bytes, uint32_t indirect_out[4], float* total_time_out) {
    try {
        Hair* hair = strands_in ? new Hair(nullptr, nullptr, std::vector<float>(strands_in, strands_in + (size_t)S * 3 * N * 4), S, N)
                                : new Hair(nullptr, nullptr, std::string(obj_path), S, N);
        std::vector<Collider> colliders = {                                       // main.cpp:229-237
            Collider({ 2.0f, 0.0f, 1.0f }, { 0.0f, 0.0f, 0.0f }, { 1.0f, 1.0f, 1.0f }),
            Collider({ 0.0f, 2.64f, 0.08f }, { -38.270f, 0.0f, 0.0f }, { 0.817f, 1.158f, 1.01f }),
            Collider({ 0.0f, 1.35f, -0.288f }, { 18.301f, 0.0f, 0.0f }, { 0.457f, 1.0f, 0.538f }),
            Collider({ 0.0f, -0.380f, -0.116f }, { -17.260f, 0.0f, 0.0f }, { 1.078f, 1.683f, 0.974f }),
            Collider({ -0.698f, 0.087f, -0.36f }, { -20.254f, 13.144f, 34.5f }, { 0.721f, 1.0f, 0.724f }),
            Collider({ 0.698f, 0.087f, -0.36f }, { -20.254f, 13.144f, -34.5f }, { 0.721f, 1.0f, 0.724f }),
        };
        Scene* scene = new Scene(nullptr, nullptr, colliders, {});
        scene->AddHair(hair);
        scene->SetFixedDeltaTime(fixed_dt);
        int rc = 0;
        {
            Renderer renderer(nullptr, nullptr, scene, nullptr, nullptr, flags, 0);
            for (int f = 0; f < frames; ++f) {
                scene->UpdateTime();
                renderer.Frame();
                if (sphere_moves) scene->translateSphere(vec3{ sphere_moves[3 * f], sphere_moves[3 * f + 1], sphere_moves[3 * f + 2] });
            }
            std::vector<float> out;
            renderer.DownloadStrands(0, out);
            if (bytes != out.size() * sizeof(float)) { g_err = "strands_out must be S*48*N bytes"; rc = -1; }
            else std::memcpy(strands_out, out.data(), bytes);
            const StrandDrawIndirect ind = renderer.ReadIndirectDraw(0);
            if (indirect_out) std::memcpy(indirect_out, &ind, 16);
            if (total_time_out) *total_time_out = scene->GetTime().totalTime;
        }
        delete scene;
        delete hair;
        return rc;
    } catch (const std::exception& e) { g_err = e.what(); return -2; }
}

}  // extern "C"
