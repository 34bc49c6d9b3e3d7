#!/usr/bin/env python3
"""Build oracle/_ref/ from the reference's own sources (TEST INFRASTRUCTURE ONLY).

Nothing from /root/reference is copied into the repository: sources are compiled where
they lie and every output (generated include, .so) lands under oracle/_ref/, which is
git-ignored but travels to the GPU box.

Outputs
  oracle/_ref/libref_host.so          reference src/Strand.cpp + src/Scene.h (Collider) with
                                      the vendored glm / tiny_obj_loader, Vulkan stubbed
  oracle/_ref/libref_compute_N<N>[_wind<A|B>].so
                                      reference src/shaders/compute.comp compiled as C++
                                      against the vendored glm (see ref_compute_tu.cpp)

The shader is GLSL, so it goes through the textual substitutions below.  They are the
complete list; main()'s statements are untouched:
  R1  drop `#version`, `#extension` and the `layout(local_size_x...) in;` line
  R2  `layout(...) uniform|buffer NAME { ... } inst;`  ->  `struct NAME { ... } inst;`
      `layout(...) uniform|buffer NAME { ... };`       ->  the members become globals
  R3  unsized SSBO array `T name[];` -> `T* name;`
  R4  float literals get an `f` suffix (GLSL literals are 32-bit; C++ ones are double)
  R5  swizzles `.xyz` / `.xy` -> `.xyz()` / `.xy()` (glm's swizzle call syntax)
  R6  `void main()` -> `void shader_main()`
  R7  (only with --points N != 10) `#define NUM_CURVE_POINTS 10` -> N   [parameterised build]
  R8  (only with --wind A|B) un-comment the authors' wind line compute.comp:151 / :152
hair.tese (oracle/_ref/libref_tese_N<N>.so, see ref_tese_tu.cpp) additionally needs its interface variables declared:
  R9  drop `layout(isolines) in;`
  R10 `layout(location = k) in vec4[][N] name;` -> `static vec4 (*name)[N];`; `layout(location = k) out T name;` -> `static T name;`
  R11 parameter qualifier `out vec3 x` -> `vec3& x`
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
# -ffp-contract=off: no FMA fusion, so results are the plain IEEE evaluation of the text
CXXFLAGS = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-std=c++14", "-w"]


def transliterate(src: str, points: int, wind: str) -> str:
    lines = src.split("\n")
    out = []
    i = 0
    block_re = re.compile(r"^\s*layout\s*\(.*?\)\s*(uniform|buffer)\s+(\w+)\s*\{")
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("#version") or ln.startswith("#extension"):          # R1
            i += 1
            continue
        if re.match(r"^\s*layout\s*\(\s*local_size_x", ln):                    # R1
            i += 1
            continue
        m = block_re.match(ln)
        if m:                                                                  # R2
            name = m.group(2)
            j = i + 1
            body = []
            while not lines[j].lstrip().startswith("}"):
                body.append(lines[j])
                j += 1
            closing = lines[j].strip()
            inst = closing[1:].rstrip(";").strip()
            body = [re.sub(r"^(\s*)(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"\1\2* \3;", b) for b in body]  # R3
            if inst:
                out.append("static struct %s {" % name)
                out.extend(body)
                out.append("} %s;" % inst)
            else:
                out.extend("static " + b.strip() if b.strip() and not b.strip().startswith("//") else b for b in body)
            i = j + 1
            continue
        if wind == "A" and ln.lstrip().startswith("//") and "force += 10.0 * vec3(" in ln:   # R8
            ln = ln.replace("//", "", 1)
        if wind == "B" and ln.lstrip().startswith("//") and "force += 7.0 * fbm(" in ln:     # R8
            ln = ln.replace("//", "", 1)
        out.append(ln)
        i += 1
    text = "\n".join(out)
    if points != 10:                                                           # R7
        text, n = re.subn(r"#define NUM_CURVE_POINTS 10\b", "#define NUM_CURVE_POINTS %d" % points, text)
        assert n == 1
    text, n = re.subn(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", text)  # R6
    assert n == 1
    # R4/R5 must not touch comments
    def fix_code(code: str) -> str:
        code = re.sub(r"(?<![\w.])(\d+\.\d*|\.\d+)(?![\w.])", r"\1f", code)    # R4
        code = re.sub(r"\.(xyz|xy)\b(?!\s*\()", r".\1()", code)                # R5
        return code
    fixed = []
    for ln in text.split("\n"):
        if "//" in ln:
            code, comment = ln.split("//", 1)
            fixed.append(fix_code(code) + "//" + comment)
        else:
            fixed.append(fix_code(ln))
    return "\n".join(fixed) + "\n"


def transliterate_tese(src: str, points: int) -> str:
    lines = []
    for ln in src.split("\n"):
        if re.match(r"^\s*layout\s*\(\s*isolines\s*\)\s*in\s*;", ln):                                   # R9
            continue
        m = re.match(r"^\s*layout\s*\(\s*location\s*=\s*\d+\s*\)\s*in\s+vec4\s*\[\s*\]\s*\[\s*(\w+)\s*\]\s*(\w+)\s*;", ln)
        if m:                                                                                          # R10
            lines.append("static vec4 (*%s)[%s];" % (m.group(2), m.group(1)))
            continue
        m = re.match(r"^\s*layout\s*\(\s*location\s*=\s*\d+\s*\)\s*out\s+(\w+)\s+(\w+)\s*;", ln)
        if m:                                                                                          # R10
            lines.append("static %s %s;" % (m.group(1), m.group(2)))
            continue
        if ln.lstrip().startswith("//layout"):
            lines.append(ln)
            continue
        lines.append(re.sub(r"\bout\s+vec3\s+(\w+)", r"vec3& \1", ln))                                # R11
    return transliterate("\n".join(lines), points, "")


def run(cmd, out=None, deps=()):
    """Compile unless `out` is newer than every dependency."""
    if out and os.path.exists(out) and all(os.path.exists(d) and os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--points", type=int, nargs="*", default=[10, 16, 32, 64])
    ap.add_argument("--wind", nargs="*", default=["", "B", "A"])
    args = ap.parse_args()
    ref = args.reference
    if not os.path.isdir(os.path.join(ref, "src")):
        print("reference tree not present at %s: keeping any prebuilt oracle/_ref" % ref)
        return 0
    os.makedirs(os.path.join(OUT, "gen"), exist_ok=True)
    glm_inc = os.path.join(ref, "external", "glm")
    src_dir = os.path.join(ref, "src")

    # host pieces: the reference's Strand.cpp, compiled unmodified, + Scene.h's Collider
    run([CXX] + CXXFLAGS + ["-I", os.path.join(HERE, "stubs"), "-I", src_dir, "-I", glm_inc,
                            os.path.join(HERE, "ref_host.cpp"), os.path.join(src_dir, "Strand.cpp"),
                            "-o", os.path.join(OUT, "libref_host.so")],
        out=os.path.join(OUT, "libref_host.so"),
        deps=[os.path.join(HERE, "ref_host.cpp"), os.path.join(src_dir, "Strand.cpp"), os.path.join(src_dir, "Scene.h"), os.path.abspath(__file__)])

    shader = open(os.path.join(src_dir, "shaders", "compute.comp")).read()
    for n in args.points:
        for w in args.wind:
            if w and n != 10 and n != 32:
                continue
            tag = "N%d%s" % (n, ("_wind" + w) if w else "")
            gen_dir = os.path.join(OUT, "gen", tag)
            os.makedirs(gen_dir, exist_ok=True)
            gen = os.path.join(gen_dir, "compute_comp.gen.inc")
            text = transliterate(shader, n, w)
            if not os.path.exists(gen) or open(gen).read() != text:
                with open(gen, "w") as f:
                    f.write(text)
            so = os.path.join(OUT, "libref_compute_%s.so" % tag)
            run([CXX] + CXXFLAGS + ["-I", gen_dir, "-I", glm_inc, os.path.join(HERE, "ref_compute_tu.cpp"), "-o", so],
                out=so, deps=[gen, os.path.join(HERE, "ref_compute_tu.cpp")])
    tese = open(os.path.join(src_dir, "shaders", "hair.tese")).read()
    for n in (10, 16):
        gen_dir = os.path.join(OUT, "gen", "tese_N%d" % n)
        os.makedirs(gen_dir, exist_ok=True)
        gen = os.path.join(gen_dir, "hair_tese.gen.inc")
        text = transliterate_tese(tese, n)
        if not os.path.exists(gen) or open(gen).read() != text:
            with open(gen, "w") as f:
                f.write(text)
        so = os.path.join(OUT, "libref_tese_N%d.so" % n)
        run([CXX] + CXXFLAGS + ["-I", gen_dir, "-I", glm_inc, os.path.join(HERE, "ref_tese_tu.cpp"), "-o", so],
            out=so, deps=[gen, os.path.join(HERE, "ref_tese_tu.cpp")])
    return 0


if __name__ == "__main__":
    sys.exit(main())
