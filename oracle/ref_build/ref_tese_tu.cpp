/*
 * ref_tese_tu.cpp -- evaluates the reference's OWN tessellation-evaluation shader text on the CPU.
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * src/shaders/hair.tese places every vertex the fixed-function tessellator generates for a guide strand (one patch,
 * 12 isolines x 42 segments, hair.tesc:19-20).  build_ref.py turns the shader text into C++ against the reference's vendored
 * glm with the same purely textual substitutions as for compute.comp (plus the interface-variable ones listed there) and
 * writes it to oracle/_ref/gen/tese_N<N>/hair_tese.gen.inc, #included below.  This file adds what the pipeline provides:
 *   - gl_TessCoord for isolines with equal spacing: x = j / divisions along the line, y = k / isolines (Vulkan spec,
 *     "Isoline Tessellation"), z = 0;
 *   - in_curvePoints[0] = the guide's curve points (hair.vert:9-18 passes them through, hair.tesc copies them).  The shader
 *     also reads in_curvePoints[1], [2] (its multi-strand experiment) although a patch has ONE control point
 *     (Renderer.cpp:1612): those reads are out of bounds in the reference and feed only values that are overwritten before
 *     use (`pos = singleStrandPos`, hair.tese:304); here they see copies of the same strand.  At v = 1 the shader indexes
 *     curve point N (hair.tese:165-166, out of bounds): the buffer is padded and callers skip j = divisions;
 *   - camera matrices = identity (they only feed out_viewDir / out_lightDir, which the expansion does not produce).
 */
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <cstring>

using namespace glm;

static vec3 gl_TessCoord;
static vec4 gl_Position;

#include "hair_tese.gen.inc"

extern "C" {

int ref_tese_num_curve_points(void) { return NUM_CURVE_POINTS; }

/* points: N x 4 floats of ONE guide strand; out: pos[3], strandWidth, out_u[3] (unit tangent), uv.x (= u) */
void ref_tese_eval(const float* points, int isolines, int divisions, int k, int j, float* out8) {
    static vec4 buf[3][NUM_CURVE_POINTS + 2];
    for (int c = 0; c < 3; ++c) {
        std::memcpy((void*)buf[c], points, sizeof(vec4) * NUM_CURVE_POINTS);
        buf[c][NUM_CURVE_POINTS] = buf[c][NUM_CURVE_POINTS - 1];
        buf[c][NUM_CURVE_POINTS + 1] = buf[c][NUM_CURVE_POINTS - 1];
    }
    static vec4 view3[3][NUM_CURVE_POINTS];
    /* the shader's array type is vec4[][NUM_CURVE_POINTS]: rows must be exactly NUM_CURVE_POINTS apart; the v = 1 over-read of
     * row 0 then lands in row 1's first element (a copy of the root), which callers never compare */
    for (int c = 0; c < 3; ++c) std::memcpy((void*)view3[c], (void*)buf[c], sizeof(vec4) * NUM_CURVE_POINTS);
    in_curvePoints = view3;
    camera.view = mat4(1.0f); camera.proj = mat4(1.0f);
    shadowCamera.view = mat4(1.0f); shadowCamera.proj = mat4(1.0f);
    gl_TessCoord = vec3((float)j / (float)divisions, (float)k / (float)isolines, 0.0f);
    shader_main();
    out8[0] = gl_Position.x; out8[1] = gl_Position.y; out8[2] = gl_Position.z; out8[3] = out_strandWidth;
    out8[4] = out_u.x; out8[5] = out_u.y; out8[6] = out_u.z; out8[7] = out_uv.x;
}

} /* extern "C" */
