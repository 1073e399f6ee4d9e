#!/usr/bin/env python3
"""Build the UNMODIFIED-ALGORITHM reference renderer into oracle/_ref/ (test infrastructure only).

This is the "compiled oracle" of SURVEY.md section 8(c): the reference's own C++ sources are compiled
where they lie (a scratch copy under /tmp is patched, never the mount, and no source is copied into
this repository).  Outputs go ONLY to oracle/_ref/ (git-ignored, travels to the GPU box):

  oracle/_ref/raytracer        reference + P1 + P2 + P3            (timed CPU baseline, `--impl reference`)
  oracle/_ref/raytracer_probe  the same + ray counters + hit dump  (parity oracle; never timed)
  oracle/_ref/raytracer_dropin the reference's parser / Scene / Camera / main() with its render loop (main.cpp:164-192)
                               replaced by ONE call into libdorktracer.so: the reference sources + the flattener
                               advanced-cpu-raytracing_b200/dropin/dt_flatten_scene.cpp, patches P1-P4, Q1 and D1-D3 below.
                               The proof that the C ABI is a drop-in for the reference's own host code (needs a GPU to RUN).

Recorded patches (each is applied by exact-anchor replacement and asserted to match exactly once):
  P1  InstancedMesh::SetMaterial also sets Shape::material_id  (instancedMesh.cpp:11-13; without it any
      instanced scene with a light segfaults at raytracer.cpp:590)
  P2  MeshLight face pick uses uniform_int_distribution(0, faceCount-1)  (meshLight.h:22; out-of-bounds)
  P3  THREAD_COUNT (main.cpp:15) becomes a run-time value read from $DT_THREADS (default 8, as shipped)
  P4  Mesh::surfaceArea (mesh.hpp:19) is initialised to 0 in the constructor: parser.cpp:608 accumulates face areas into
      it without ever setting it, so MeshLight::getSample's selection weight (meshLight.h:31) depends on heap garbage
      (observed: radiance of a mesh-lit scene changes in the 6th digit once a PLY mesh is parsed before the LightMesh)
  Q1  (quiet) the one-line-per-PLY-face print in Scene::createFace (parser.cpp:813) is removed; it is
      pure stdout noise (10M lines on config 5) and does not touch the hot path.
Drop-in edits of main.cpp (raytracer_dropin):
  D1  after scene.loadFromXml(argv[1]) (main.cpp:136): dt_dropin_create(scene)  -- flatten the Scene, create the GPU scene
  D2  main.cpp:164-185 (thread spawn over renderThreadMain ... join) -> dt_dropin_render(gpu, cam, image, hdrImage)
  D3  main.cpp:190 cam.GetTonemappedImage(...) removed (dt_render tonemaps); stbi_write_hdr / stbi_write_png stay
Probe-only instrumentation (raytracer_probe):
  I1  thread-local closest/shadow ray counters (one per Raytracer::IntersectObjects / CastShadowRay call),
      summed and printed as "DT_RAYS closest=<n> shadow=<n>"
  I2  $DT_DUMP_HITS=<file>: per pixel primary hit (int32 shapeIdx, int32 faceIdx, float t), shapeIdx =
      index in scene.meshes or meshes.size()+index in scene.spheres, -1 on miss; faceIdx = post-build
      index into Mesh::faces (base mesh for instances), -1 for spheres.
  I3  $DT_DUMP_HDR=<file>: raw float32 W*H*3 of the resolved radiance before LDR clamp/tonemap.

Usage: python oracle/build_ref.py [--force]
"""
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
OUT = os.path.join(HERE, "_ref")

# The reference's own build line (src/Makefile:1-2): g++ *.cpp -o raytracer -std=c++11 -O3 -lpthread
CXXFLAGS = ["-std=c++11", "-O3", "-w"]


def patch(text, anchor, replacement, name):
    n = text.count(anchor)
    if n != 1:
        raise RuntimeError("patch %s: anchor matched %d times (expected 1)" % (name, n))
    return text.replace(anchor, replacement)


def edit(path, fn):
    with open(path, "r", encoding="utf-8", errors="surrogateescape") as f:
        t = f.read()
    t2 = fn(t)
    with open(path, "w", encoding="utf-8", errors="surrogateescape") as f:
        f.write(t2)


def apply_common(d):
    # P1
    edit(os.path.join(d, "instancedMesh.cpp"), lambda t: patch(
        t, "this->material_id = matId;", "this->material_id = matId; Shape::SetMaterial(matId);", "P1"))
    # P2
    edit(os.path.join(d, "meshLight.h"), lambda t: patch(
        t, "uniform_int_distribution<>(0, faceCount)", "uniform_int_distribution<>(0, faceCount-1)", "P2"))
    # P3
    edit(os.path.join(d, "main.cpp"), lambda t: patch(
        t, "#define THREAD_COUNT 8",
        "#include <cstdlib>\nstatic int dt_thread_count(){ const char* e = getenv(\"DT_THREADS\"); "
        "int n = e ? atoi(e) : 8; return n > 0 ? n : 8; }\n#define THREAD_COUNT (dt_thread_count())", "P3"))
    # P4
    edit(os.path.join(d, "mesh.cpp"), lambda t: patch(
        t, "this->vertexOffset = 0;\n    this->textureOffset = 0;", "this->vertexOffset = 0;\n    this->textureOffset = 0;\n    this->surfaceArea = 0;", "P4"))
    # Q1
    edit(os.path.join(d, "parser.cpp"), lambda t: patch(
        t, 'std::cout << " total area of mesh: " << mesh->surfaceArea << std::endl;\n\n    return face;',
        "return face;", "Q1"))


PROBE_GLOBALS = r'''
#include <atomic>
#include <cstdio>
#include <cstdlib>
namespace dtprobe {
    thread_local long long tl_closest = 0, tl_shadow = 0;
    thread_local int tl_lastFace = -1;
    std::atomic<long long> g_closest(0), g_shadow(0);
    int* g_hitShape = nullptr; int* g_hitFace = nullptr; float* g_hitT = nullptr;
    float* g_hdr = nullptr;
    int g_W = 0, g_H = 0;
    struct Flush { ~Flush(){ g_closest += tl_closest; g_shadow += tl_shadow; } };
    thread_local Flush tl_flush;
}
'''


def apply_probe(d):
    # I1 + I2 in raytracer.cpp
    def rt(t):
        t = patch(t, "using namespace DorkTracer;\n\nRaytracer::Raytracer(Scene& scene){",
                  "using namespace DorkTracer;\nnamespace dtprobe { extern thread_local long long tl_closest, tl_shadow;"
                  " extern thread_local int tl_lastFace; extern int* g_hitShape; extern int* g_hitFace; extern float* g_hitT;"
                  " extern int g_W, g_H; struct Flush{ ~Flush(); }; extern thread_local Flush tl_flush; }\n"
                  "\nRaytracer::Raytracer(Scene& scene){", "I1-decl")
        t = patch(t, "void Raytracer::IntersectObjects(Ray& ray)\n{",
                  "void Raytracer::IntersectObjects(Ray& ray)\n{\n    dtprobe::tl_closest++; (void)&dtprobe::tl_flush;", "I1-closest")
        t = patch(t, "bool Raytracer::CastShadowRay(Ray& shadowRay, float lightSourceT)\n{",
                  "bool Raytracer::CastShadowRay(Ray& shadowRay, float lightSourceT)\n{\n    dtprobe::tl_shadow++;", "I1-shadow")
        t = patch(t, "    IntersectObjects(ray);\n    \n    if(ray.hitInfo.hasHit)\n    {\n       return PerformShading(ray, cam.position, scene.max_recursion_depth);",
                  "    dtprobe::tl_lastFace = -1;\n    IntersectObjects(ray);\n"
                  "    if(dtprobe::g_hitShape && coordX >= 0 && coordX < dtprobe::g_W && coordY >= 0 && coordY < dtprobe::g_H){\n"
                  "        int pi = coordX + coordY * dtprobe::g_W; int si = -1; int fi = -1;\n"
                  "        if(ray.hitInfo.hasHit){\n"
                  "            for(int k = 0; k < (int)scene.meshes.size(); k++) if((Shape*)scene.meshes[k] == ray.hitInfo.hitShape){ si = k; fi = dtprobe::tl_lastFace; }\n"
                  "            for(int k = 0; k < (int)scene.spheres.size(); k++) if((Shape*)scene.spheres[k] == ray.hitInfo.hitShape){ si = (int)scene.meshes.size() + k; fi = -1; }\n"
                  "        }\n"
                  "        dtprobe::g_hitShape[pi] = si; dtprobe::g_hitFace[pi] = fi; dtprobe::g_hitT[pi] = ray.hitInfo.hasHit ? ray.hitInfo.minT : INFINITY;\n"
                  "    }\n"
                  "    if(ray.hitInfo.hasHit)\n    {\n       return PerformShading(ray, cam.position, scene.max_recursion_depth);", "I2-dump")
        return t
    edit(os.path.join(d, "raytracer.cpp"), rt)

    # I2: remember the face index of the last accepted face (mesh.cpp:198-200)
    def mesh(t):
        t = patch(t, "using namespace DorkTracer;\n\nDorkTracer::Mesh::Mesh(",
                  "using namespace DorkTracer;\nnamespace dtprobe { extern thread_local int tl_lastFace; }\n\nDorkTracer::Mesh::Mesh(", "I2-decl")
        t = patch(t, "    return IntersectFace(ray, this->faces[faceIdx]);",
                  "    bool dt_r = IntersectFace(ray, this->faces[faceIdx]); if(dt_r) dtprobe::tl_lastFace = (int)faceIdx; return dt_r;", "I2-face")
        return t
    edit(os.path.join(d, "mesh.cpp"), mesh)
    # a later-tested sphere that wins must reset the face index: handled in the dump (fi=-1 for spheres).

    def main(t):
        t = patch(t, "struct RenderThreadArgs{", PROBE_GLOBALS + "\nstruct RenderThreadArgs{", "I-globals")
        t = patch(t, "        std::vector<std::thread> renderThreads;",
                  "        if(getenv(\"DT_DUMP_HITS\")){ dtprobe::g_W = width; dtprobe::g_H = height;\n"
                  "            dtprobe::g_hitShape = new int[width*height]; dtprobe::g_hitFace = new int[width*height]; dtprobe::g_hitT = new float[width*height];\n"
                  "            for(int k = 0; k < width*height; k++){ dtprobe::g_hitShape[k] = -2; dtprobe::g_hitFace[k] = -2; dtprobe::g_hitT[k] = 0.f; } }\n"
                  "        if(getenv(\"DT_DUMP_HDR\")){ dtprobe::g_hdr = new float[(size_t)width*height*3](); }\n"
                  "        std::vector<std::thread> renderThreads;", "I2-alloc")
        # I3: raw radiance before clamp
        t = patch(t, "            uint32_t imgIdx = 3 * (x + y * width);",
                  "            uint32_t imgIdx = 3 * (x + y * width);\n"
                  "            if(dtprobe::g_hdr){ dtprobe::g_hdr[imgIdx] = color.x; dtprobe::g_hdr[imgIdx+1] = color.y; dtprobe::g_hdr[imgIdx+2] = color.z; }", "I3-hdr")
        t = patch(t, "        if(cam.hasTonemapper)\n        {\n            // Post process",
                  "        if(dtprobe::g_hitShape){ FILE* f = fopen(getenv(\"DT_DUMP_HITS\"), \"wb\");\n"
                  "            for(int k = 0; k < width*height; k++){ fwrite(&dtprobe::g_hitShape[k], 4, 1, f); fwrite(&dtprobe::g_hitFace[k], 4, 1, f); fwrite(&dtprobe::g_hitT[k], 4, 1, f); }\n"
                  "            fclose(f); }\n"
                  "        if(dtprobe::g_hdr){ FILE* f = fopen(getenv(\"DT_DUMP_HDR\"), \"wb\"); fwrite(dtprobe::g_hdr, 4, (size_t)width*height*3, f); fclose(f); }\n"
                  "        if(cam.hasTonemapper)\n        {\n            // Post process", "I2-write")
        t = patch(t, '    std::cout << "Rendering took: " << elapsed_seconds.count() << "s\\n";',
                  '    std::cout << "Rendering took: " << elapsed_seconds.count() << "s\\n";\n'
                  '    std::cout << "DT_RAYS closest=" << dtprobe::g_closest.load() << " shadow=" << dtprobe::g_shadow.load() << std::endl;', "I1-print")
        return t
    edit(os.path.join(d, "main.cpp"), main)


REPO = os.path.dirname(HERE)
PKG = os.path.join(REPO, "advanced-cpu-raytracing_b200")
DROPIN_SRC = os.path.join(PKG, "dropin", "dt_flatten_scene.cpp")


def apply_dropin(d):
    def main(t):
        t = patch(t, "int main(int argc, char* argv[])\n{",
                  "void* dt_dropin_create(DorkTracer::Scene& scene);\nint dt_dropin_render(void* gpu, DorkTracer::Camera& cam, unsigned char* image, float* hdrImage);\n"
                  "void dt_dropin_destroy(void* gpu);\n\nint main(int argc, char* argv[])\n{", "D0")
        t = patch(t, "    scene.loadFromXml(argv[1]);\n", "    scene.loadFromXml(argv[1]);\n    void* dt_gpu = dt_dropin_create(scene);\n    if(!dt_gpu) return 1;\n", "D1")
        a = t.index("        std::vector<std::thread> renderThreads;")
        b = t.index("        if(cam.hasTonemapper)\n        {\n            // Post process")
        t = t[:a] + "        if(dt_dropin_render(dt_gpu, cam, image, hdrImage) != 0) return 1;      // replaces main.cpp:164-185\n\n" + t[b:]      # D2
        t = patch(t, "            cam.GetTonemappedImage(width,height, hdrImage, image);\n", "", "D3")
        return t
    edit(os.path.join(d, "main.cpp"), main)


def compile_variant(srcdir, out_bin, dropin=False):
    cpps = sorted(f for f in os.listdir(srcdir) if f.endswith(".cpp"))
    objdir = os.path.join(srcdir, "_obj")
    os.makedirs(objdir, exist_ok=True)

    def cc(f):
        o = os.path.join(objdir, f[:-4] + ".o")
        subprocess.run(["g++"] + CXXFLAGS + ["-c", f, "-o", o], cwd=srcdir, check=True)
        return o
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(cc, cpps))
    if dropin:
        # the flattener reads private members of the reference's classes: -fno-access-control for THIS translation unit only
        o = os.path.join(objdir, "dt_flatten_scene.o")
        subprocess.run(["g++"] + CXXFLAGS + ["-fno-access-control", "-I" + os.path.join(REPO, "include"), "-I" + srcdir, "-c", DROPIN_SRC, "-o", o], cwd=srcdir, check=True)
        subprocess.run(["g++"] + objs + [o, "-o", out_bin, "-L" + PKG, "-ldorktracer", "-Wl,-rpath,$ORIGIN/../../advanced-cpu-raytracing_b200", "-lpthread"], check=True)
    else:
        subprocess.run(["g++"] + objs + ["-o", out_bin, "-lpthread"], check=True)


def copy_sources(dst):
    for f in os.listdir(REF_SRC):
        p = os.path.join(REF_SRC, f)
        if os.path.isfile(p) and f.rsplit(".", 1)[-1] in ("cpp", "h", "hpp"):
            shutil.copy(p, os.path.join(dst, f))


def build(force=False):
    want = [os.path.join(OUT, "raytracer"), os.path.join(OUT, "raytracer_probe"), os.path.join(OUT, "raytracer_dropin")]
    lib = os.path.join(PKG, "libdorktracer.so")
    abi = os.path.join(REPO, "include", "dorktracer.h")         # the drop-in embeds the ABI structs: a header change makes it stale too
    stale = os.path.exists(want[2]) and any(os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(want[2]) for f in (DROPIN_SRC, abi))
    if not force and not stale and all(os.path.exists(w) for w in want):
        return True
    if not os.path.isdir(REF_SRC):
        # GPU box: only the prebuilt files exist.
        return all(os.path.exists(w) for w in want)
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="dt_ref_build_")
    try:
        a = os.path.join(tmp, "plain"); os.makedirs(a)
        copy_sources(a); apply_common(a)
        compile_variant(a, want[0])
        b = os.path.join(tmp, "probe"); os.makedirs(b)
        copy_sources(b); apply_common(b); apply_probe(b)
        compile_variant(b, want[1])
        if os.path.exists(lib):                    # links against the CUDA library (built by `make cuda` first)
            c = os.path.join(tmp, "dropin"); os.makedirs(c)
            copy_sources(c); apply_common(c); apply_dropin(c)
            compile_variant(c, want[2], dropin=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ok" if ok else "unavailable")
    sys.exit(0 if ok else 1)
