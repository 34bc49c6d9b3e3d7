// rvh_host.cpp -- see rvh_host.hpp.  Host logic only; every device operation goes through include/rvh.h.
#include "rvh_host.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rvh_host {

// ---- Collider (Scene.h:28-38) ---------------------------------------------------------------------
Collider::Collider(vec3 trans, vec3 rot, vec3 scale) {
    const float t[3] = { trans.x, trans.y, trans.z }, r[3] = { rot.x, rot.y, rot.z }, s[3] = { scale.x, scale.y, scale.z };
    float out[48];
    rvh_collider_build(t, r, s, out);
    std::memcpy(transform.m, out, 64); std::memcpy(inv.m, out + 16, 64); std::memcpy(invTrans.m, out + 32, 64);
}

// ---- OBJ reading + triangulation --------------------------------------------------------------------
// The reference loads the follicle surface with tiny_obj_loader 2.0.0 (vendored there as
// src/tiny_obj_loader.h), triangulate = true.  That library is a third-party dependency; what matters
// for this path is the ORDER of the resulting triangle corners, because follicles index into it
// (Strand.cpp:95-100).  This reader keeps faces in file order and clips ears starting at the face's
// first corner, which yields (0,1,2),(0,2,3) for every convex quad -- the same order that library
// produces.  tests/test_host_mirror.py pins the result against the reference's own Hair::Hair output.
namespace {

struct Corner { int v = -1, vn = -1; };

int fix_index(long idx, size_t count) {      // OBJ indices are 1-based; negative = relative to the end
    if (idx > 0) return (int)idx - 1;
    if (idx < 0) return (int)count + (int)idx;
    return -1;
}

bool parse_corner(const std::string& tok, size_t nv, size_t nvn, Corner& c) {
    // v, v/vt, v//vn, v/vt/vn
    const char* p = tok.c_str();
    char* end = nullptr;
    long a = std::strtol(p, &end, 10);
    if (end == p) return false;
    c.v = fix_index(a, nv);
    if (*end != '/') return true;
    p = end + 1;
    if (*p != '/') { std::strtol(p, &end, 10); p = end; }     // texture index, unused
    if (*p != '/') return true;
    ++p;
    long n = std::strtol(p, &end, 10);
    if (end != p) c.vn = fix_index(n, nvn);
    return true;
}

bool inside_triangle(const float px[3], const float py[3], float x, float y) {
    // even-odd crossing test over the three edges
    bool in = false;
    for (int i = 0, j = 2; i < 3; j = i++) {
        if (((py[i] > y) != (py[j] > y)) && (x < (px[j] - px[i]) * (y - py[i]) / (py[j] - py[i]) + px[i])) in = !in;
    }
    return in;
}

void triangulate_face(const std::vector<Corner>& face, const std::vector<float>& V, std::vector<Corner>& out) {
    size_t n = face.size();
    if (n < 3) return;
    if (n == 3) { out.insert(out.end(), face.begin(), face.end()); return; }
    // projection plane: drop the axis along which the first non-degenerate corner's normal is largest
    int ax0 = 1, ax1 = 2;
    for (size_t k = 0; k < n; ++k) {
        const float* a = &V[3 * face[k].v]; const float* b = &V[3 * face[(k + 1) % n].v]; const float* c = &V[3 * face[(k + 2) % n].v];
        const float e0[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] }, e1[3] = { c[0] - b[0], c[1] - b[1], c[2] - b[2] };
        const float cx = std::fabs(e0[1] * e1[2] - e0[2] * e1[1]), cy = std::fabs(e0[2] * e1[0] - e0[0] * e1[2]), cz = std::fabs(e0[0] * e1[1] - e0[1] * e1[0]);
        const float eps = 1.1920929e-7f;
        if (cx > eps || cy > eps || cz > eps) {
            if (!(cx > cy && cx > cz)) { ax0 = 0; if (cz > cx && cz > cy) ax1 = 1; }
            break;
        }
    }
    float area = 0.f;
    for (size_t k = 0; k < n; ++k) {
        const float* a = &V[3 * face[k].v]; const float* b = &V[3 * face[(k + 1) % n].v];
        area += (a[ax0] * b[ax1] - a[ax1] * b[ax0]) * 0.5f;
    }
    std::vector<Corner> rest = face;
    size_t guess = 0, budget = rest.size(), prev = rest.size();
    while (rest.size() > 3 && budget > 0) {
        n = rest.size();
        if (guess >= n) guess -= n;
        if (prev != n) { prev = n; budget = n; } else { --budget; }
        Corner tri[3]; float px[3], py[3];
        for (int k = 0; k < 3; ++k) { tri[k] = rest[(guess + k) % n]; px[k] = V[3 * tri[k].v + ax0]; py[k] = V[3 * tri[k].v + ax1]; }
        const float cross = (px[1] - px[0]) * (py[2] - py[1]) - (py[1] - py[0]) * (px[2] - px[1]);
        if (cross * area < 0.f) { ++guess; continue; }                     // reflex corner: not an ear
        bool blocked = false;
        for (size_t o = 3; o < n && !blocked; ++o) {
            const Corner& q = rest[(guess + o) % n];
            blocked = inside_triangle(px, py, V[3 * q.v + ax0], V[3 * q.v + ax1]);
        }
        if (blocked) { ++guess; continue; }
        out.push_back(tri[0]); out.push_back(tri[1]); out.push_back(tri[2]);
        rest.erase(rest.begin() + (long)((guess + 1) % n));                // clip the ear's middle corner
    }
    if (rest.size() == 3) out.insert(out.end(), rest.begin(), rest.end());
}

}  // namespace

int GeneratePointsOnMesh(const std::string& filename, int numStrands, std::vector<vec3>& points, std::vector<vec3>& pointNormals) {
    std::ifstream in(filename);
    if (!in) throw std::runtime_error("Cannot open OBJ file " + filename);
    std::vector<float> V, VN;
    std::vector<Corner> corners;       // 3 per triangle, faces in file order
    std::string line;
    std::vector<Corner> face;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string tag;
        if (!(ss >> tag)) continue;
        if (tag == "v") { float x, y, z; if (ss >> x >> y >> z) { V.push_back(x); V.push_back(y); V.push_back(z); } }
        else if (tag == "vn") { float x, y, z; if (ss >> x >> y >> z) { VN.push_back(x); VN.push_back(y); VN.push_back(z); } }
        else if (tag == "f") {
            face.clear();
            std::string tok;
            while (ss >> tok) {
                Corner c;
                if (!parse_corner(tok, V.size() / 3, VN.size() / 3, c)) break;
                if (c.v < 0 || (size_t)c.v >= V.size() / 3) throw std::runtime_error("OBJ face references a missing vertex in " + filename);
                face.push_back(c);
            }
            triangulate_face(face, V, corners);
        }
    }
    const int numTriangles = (int)(corners.size() / 3);
    if (numTriangles == 0) throw std::runtime_error("OBJ file has no faces: " + filename);
    for (const Corner& c : corners)
        if (c.vn < 0 || (size_t)c.vn >= VN.size() / 3) throw std::runtime_error("follicle surface needs per-corner normals: " + filename);

    srand(8);                                                         // Strand.cpp:88
    for (int i = 0; i < numStrands; ++i) {
        const int triangle = rand() % numTriangles;                   // Strand.cpp:17-22,93 (the [min,max) helper ignores min)
        const Corner* c = &corners[3 * (size_t)triangle];
        const float* p1 = &V[3 * c[0].v]; const float* p2 = &V[3 * c[1].v]; const float* p3 = &V[3 * c[2].v];
        const float* n = &VN[3 * c[0].vn];                            // first corner's normal for the whole face (:99)
        float u = rand() / (float)RAND_MAX;
        float v = rand() / (float)RAND_MAX;
        if (u + v >= 1.f) { u = 1 - u; v = 1 - v; }                   // :103-106
        const float w = 1.f - u - v;
        vec3 p;
        p.x = (p1[0] * u + p2[0] * v) + p3[0] * w;
        p.y = (p1[1] * u + p2[1] * v) + p3[1] * w;
        p.z = (p1[2] * u + p2[2] * v) + p3[2] * w;
        points.push_back(p);
        pointNormals.push_back(vec3{ n[0], n[1], n[2] });
    }
    return numStrands;
}

// ---- Hair (Strand.cpp:149-191) ------------------------------------------------------------------------
void Hair::buildFromFollicles(const std::vector<vec3>& roots, const std::vector<vec3>& normals) {
    const int N = numCurvePoints;
    strands.assign((size_t)numStrands * 3 * N * 4, 0.0f);
    const float length = 2.5f;
    const float seg = (float)(length / (N - 1.0));                    // Strand.cpp:172: double division, then float
    for (int i = 0; i < numStrands; ++i) {
        float* cp = &strands[(size_t)i * 3 * N * 4];
        float* cv = cp + (size_t)N * 4;
        vec3 cur = roots[i];
        vec3 dir = normals[i];
        dir.z -= 2.0f; dir.y += 5.0f; dir.x += 0.05f;                 // :168-171, not normalised
        for (int j = 0; j < N; ++j) {
            cp[4 * j + 0] = cur.x; cp[4 * j + 1] = cur.y; cp[4 * j + 2] = cur.z; cp[4 * j + 3] = 1.0f;
            cv[4 * j + 0] = 0.0f; cv[4 * j + 1] = 0.0f; cv[4 * j + 2] = -1.0f; cv[4 * j + 3] = 0.0f;
            cur.x += seg * dir.x; cur.y += seg * dir.y; cur.z += seg * dir.z;
        }
    }
    indirectDraw = StrandDrawIndirect{ (uint32_t)numStrands, 1u, 0u, 0u };   // :178-182
}

Hair::Hair(Device* device, VkCommandPool commandPool, std::string objFilename)
    : Hair(device, commandPool, std::move(objFilename), (int)NUM_STRANDS, (int)NUM_CURVE_POINTS) {}

Hair::Hair(Device*, VkCommandPool, std::string objFilename, int S, int N) : numStrands(S), numCurvePoints(N) {
    if (S < 1 || N < 2) throw std::runtime_error("Hair needs numStrands >= 1 and numCurvePoints >= 2");
    std::vector<vec3> roots, normals;
    GeneratePointsOnMesh(objFilename, S, roots, normals);
    buildFromFollicles(roots, normals);
}

Hair::Hair(Device*, VkCommandPool, std::vector<float> aos, int S, int N) : numStrands(S), numCurvePoints(N), strands(std::move(aos)) {
    if (S < 1 || N < 2 || strands.size() != (size_t)S * 3 * N * 4) throw std::runtime_error("Hair: strands must be float[S][3][N][4]");
    indirectDraw = StrandDrawIndirect{ (uint32_t)S, 1u, 0u, 0u };
}

// ---- Scene ---------------------------------------------------------------------------------------------
Scene::Scene(Device* device, VkCommandPool, std::vector<Collider> colliders, std::vector<Model*> models)
    : device(device), models(std::move(models)), colliders(std::move(colliders)) {
    grid.assign((size_t)64 * 64 * 64, GridCell(ivec3{ 0, 0, 0 }, 0));                  // Scene.cpp:16-20: GRID_DIM^3 zero cells
}

void Scene::UpdateTime() {
    if (fixedDt > 0.0f) {
        time.deltaTime = fixedDt;
    } else {                                                          // Scene.cpp:79-83
        const auto now = std::chrono::high_resolution_clock::now();
        time.deltaTime = std::chrono::duration_cast<std::chrono::duration<float>>(now - startTime).count();
        startTime = now;
    }
    time.totalTime += time.deltaTime;
}

void Scene::translateSphere(vec3 translation) {
    if (colliders.empty()) return;                                    // Scene.cpp:112
    float c48[48];
    std::memcpy(c48, &colliders[0], sizeof(Collider));
    const float t[3] = { translation.x, translation.y, translation.z };
    rvh_collider_translate(c48, t);
    std::memcpy(&colliders[0], c48, sizeof(Collider));
}

// ---- Renderer (compute half) ------------------------------------------------------------------------------
void Renderer::check(int status, rvh_ctx* ctx, const char* what) const {
    if (status != RVH_OK) {
        const char* msg = rvh_last_error(ctx);
        throw std::runtime_error(std::string(what) + ": " + (msg ? msg : "unknown error"));
    }
}

Renderer::Renderer(Device* device, SwapChain* swapChain, Scene* scene, Camera* camera, Camera* shadowCamera)
    : Renderer(device, swapChain, scene, camera, shadowCamera, RVH_GRID_ON, 0) {}

Renderer::Renderer(Device* device, SwapChain* swapChain, Scene* scene, Camera* camera, Camera* shadowCamera, int flags, int cudaDevice)
    : scene(scene), device(device), swapChain(swapChain), camera(camera), shadowCamera(shadowCamera), flags(flags), cudaDevice(cudaDevice) {
    if (!scene) throw std::runtime_error("Renderer needs a Scene");
    CreateComputePipeline();
    RecordComputeCommandBuffer();
}

Renderer::~Renderer() {
    for (rvh_ctx* c : contexts) rvh_destroy(c);
}

void Renderer::CreateComputePipeline() {
    for (rvh_ctx* c : contexts) rvh_destroy(c);
    contexts.clear();
    for (Hair* h : scene->GetHair()) {
        rvh_config cfg;
        rvh_default_config(&cfg, h->GetNumStrands(), h->GetNumCurvePoints());
        cfg.flags = flags;
        cfg.device = cudaDevice;
        rvh_ctx* ctx = nullptr;
        check(rvh_create(&ctx, &cfg), nullptr, "Failed to create compute pipeline");          // Renderer.cpp:1784-1786
        contexts.push_back(ctx);
        const std::vector<float>& st = h->GetInitialStrands();
        check(rvh_upload_strands_aos(ctx, st.data(), st.size() * sizeof(float)), ctx, "Failed to upload strands");
        if (h->exportedFd >= 0) check(rvh_import_strands_fd(ctx, h->exportedFd, h->exportedBytes), ctx, "Failed to import strands buffer");
        if (h->exportedIndirectFd >= 0) check(rvh_import_indirect_fd(ctx, h->exportedIndirectFd, h->exportedIndirectBytes), ctx, "Failed to import indirect-args buffer");
        if (h->exportedSemaphoreFd >= 0) check(rvh_import_semaphore_fd(ctx, h->exportedSemaphoreFd), ctx, "Failed to import semaphore");
    }
}

void Renderer::RecordComputeCommandBuffer() {
    // Renderer.cpp:2022-2077 pre-records grid clear + dispatch; rvh_step issues both every frame, so
    // all that is left is the state check the recording would have failed on.
    if (contexts.size() != scene->GetHair().size()) throw std::runtime_error("Failed to record compute command buffer");
}

void Renderer::Frame() {
    const Time& t = scene->GetTime();
    const std::vector<Collider>& cols = scene->GetColliders();
    // Scene::UpdateTime's wall-clock delta can be zero (two frames inside one clock tick).  The shader would divide by it
    // (compute.comp:195, 214: NaN velocities); this path leaves the state untouched for such a frame instead of failing it.
    if (!(t.deltaTime > 0.0f)) return;
    for (rvh_ctx* ctx : contexts) {
        // the collider and time UBOs are persistently mapped in the reference (Scene.cpp:10-13,86,133):
        // whatever the host wrote last is what the dispatch reads
        check(rvh_set_colliders(ctx, cols.data(), (int)cols.size()), ctx, "Failed to update colliders");
        check(rvh_step(ctx, t.deltaTime, t.totalTime), ctx, "Failed to submit compute command buffer");   // Renderer.cpp:2317-2319
    }
}

void Renderer::WaitIdle() {
    for (rvh_ctx* ctx : contexts) check(rvh_sync(ctx), ctx, "Failed to wait for the compute queue");
}

void Renderer::DownloadStrands(size_t hairIndex, std::vector<float>& out) {
    rvh_ctx* ctx = contexts.at(hairIndex);
    Hair* h = scene->GetHair().at(hairIndex);
    out.resize((size_t)h->GetNumStrands() * 3 * h->GetNumCurvePoints() * 4);
    check(rvh_download_strands_aos(ctx, out.data(), out.size() * sizeof(float)), ctx, "Failed to read strands back");
}

StrandDrawIndirect Renderer::ReadIndirectDraw(size_t hairIndex) {
    rvh_ctx* ctx = contexts.at(hairIndex);
    uint32_t v[4];
    check(rvh_draw_indirect(ctx, v), ctx, "Failed to read indirect draw arguments");
    return StrandDrawIndirect{ v[0], v[1], v[2], v[3] };
}

}  // namespace rvh_host
