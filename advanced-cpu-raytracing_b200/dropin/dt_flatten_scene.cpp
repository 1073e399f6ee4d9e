// dt_flatten_scene.cpp -- the ONE file a maintainer adds to the reference tree (src/) to run its render loop on the GPU.
//
// It walks the reference's own in-memory scene (DorkTracer::Scene, scene.h:32-89, filled by Scene::loadFromXml) once and
// emits the flat dt_scene_desc of include/dorktracer.h, then forwards each camera to dt_render().  Nothing here parses,
// builds or shades: the parser, Mesh::ConstructBVH, the light / material / texture classes stay the reference's.
// main.cpp changes in three places (oracle/build_ref.py applies exactly these to a scratch copy of the reference):
//     after  scene.loadFromXml(argv[1]);              ->  void* gpu = dt_dropin_create(scene);
//     main.cpp:164-185 (thread spawn ... join)         ->  dt_dropin_render(gpu, cam, image, hdrImage);
//     main.cpp:190 cam.GetTonemappedImage(...)         ->  removed (dt_render tonemaps when the camera has a tonemapper)
// stbi_write_hdr / stbi_write_png (main.cpp:191,195) stay.
//
// Several members the flattener reads are private in the reference (mesh.hpp:37-45, camera.hpp:43-46, areaLight.h:49-51,
// imageTexture.h:100-105, perlinTexture.h:137-143, spotLight.h:60-61, sphericalEnvironmentLight.h:11-12, LDRImage.h:34).  The
// proof build compiles THIS translation unit with g++ -fno-access-control; a maintainer would add const accessors instead
// (INTEGRATION.md section 2).  Host flags stay the reference's (no -march=native / -ffast-math): face normals, boxes and the BVH face
// permutation are inputs of the GPU path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "scene.h"
#include "camera.hpp"
#include "mesh.hpp"
#include "instancedMesh.hpp"
#include "sphere.hpp"
#include "material.hpp"
#include "brdfPhong.h"
#include "brdfBlinnPhong.h"
#include "brdfModifiedPhong.h"
#include "brdfModifiedBlinnPhong.h"
#include "brdfTorranceSparrow.h"
#include "pointLight.h"
#include "areaLight.h"
#include "directionalLight.h"
#include "spotLight.h"
#include "sphericalEnvironmentLight.h"
#include "meshLight.h"
#include "imageTexture.h"
#include "perlinTexture.h"
#include "LDRImage.h"
#include "HDRImage.h"

#include "dorktracer.h"

using namespace DorkTracer;

namespace {

struct DtFlat {                                   // owns every array the description points into
    std::vector<dt_material> materials;
    std::vector<dt_brdf> brdfs;
    std::vector<dt_point_light> point_lights;
    std::vector<dt_area_light> area_lights;
    std::vector<dt_directional_light> directional_lights;
    std::vector<dt_spot_light> spot_lights;
    std::vector<dt_env_light> env_lights;
    std::vector<dt_mesh_light> mesh_lights;
    std::vector<dt_image> images;
    std::vector<dt_texture> textures;
    std::vector<dt_mesh> meshes;
    std::vector<dt_shape> shapes;
    std::vector<std::vector<float>> verts, uvs;
    std::vector<std::vector<dt_face>> faces;
    std::vector<std::vector<dt_bvh2_node>> nodes;
    dt_scene_desc desc;
};

struct DtDropin {
    DtFlat flat;
    dt_scene* gpu = nullptr;       // one GPU
    dt_multi* multi = nullptr;     // DT_GPUS > 1: the library fans the frame out over several GPUs itself
};

void put3(float* d, const Vec3f& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
void put16(double* d, Matrix& m) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) d[i * 4 + j] = m[i][j]; }

template <class T>
int index_of(const std::vector<T*>& v, const void* p) {
    if (!p) return -1;
    for (size_t i = 0; i < v.size(); i++) if ((const void*)v[i] == p) return (int)i;
    return -1;
}

void fill_shape_common(Scene& sc, Shape* s, dt_shape& d) {
    d.id = s->id;
    d.mesh = -1; d.base_shape = -1;
    d.tex_diffuse = index_of(sc.textures, s->diffuseTex);
    d.tex_specular = index_of(sc.textures, s->specularTex);
    d.tex_normal = index_of(sc.textures, s->normalMap);
    d.tex_bump = index_of(sc.textures, s->bumpMap);
    d.tex_replace_all = index_of(sc.textures, s->replaceAll);
    d.has_motion_blur = s->hasMotionBlur ? 1 : 0;
    put3(d.motion_blur, s->motionBlurVector);
    put16(d.transform, s->transform);
    put16(d.inverse_transform, s->inverseTransform);
    put16(d.inverse_transpose_transform, s->inverseTransposeTransform);
}

bool flatten(Scene& sc, DtFlat& f) {
    memset(&f.desc, 0, sizeof f.desc);
    dt_scene_desc& D = f.desc;
    D.abi_version = DT_ABI_VERSION;
    D.background_color[0] = sc.background_color.x; D.background_color[1] = sc.background_color.y; D.background_color[2] = sc.background_color.z;
    D.bg_texture = index_of(sc.textures, sc.bgTexture);
    D.max_recursion_depth = sc.max_recursion_depth;
    D.shadow_ray_epsilon = Scene::shadow_ray_epsilon;
    put3(D.ambient_light, sc.ambient_light);

    // ---- BRDFs (brdf.h + the five subclasses): kind by dynamic type, flag = isEnergyConserving / kdFresnel
    for (BRDF* b : sc.brdfs) {
        dt_brdf d; d.exponent = b->exponent; d.flag = b->isEnergyConserving ? 1 : 0;
        if (dynamic_cast<BrdfTorranceSparrow*>(b)) { d.kind = DT_BRDF_TORRANCE_SPARROW; d.flag = ((BrdfTorranceSparrow*)b)->kdFresnel ? 1 : 0; }
        else if (dynamic_cast<BrdfModifiedBlinnPhong*>(b)) d.kind = DT_BRDF_MODIFIED_BLINN_PHONG;
        else if (dynamic_cast<BrdfModifiedPhong*>(b)) d.kind = DT_BRDF_MODIFIED_PHONG;
        else if (dynamic_cast<BrdfBlinnPhong*>(b)) d.kind = DT_BRDF_BLINN_PHONG;
        else d.kind = DT_BRDF_PHONG;
        f.brdfs.push_back(d);
    }
    // ---- materials (material.hpp:8-50), addressed by POSITION (materials[matId - 1], raytracer.cpp:73)
    for (Material& m : sc.materials) {
        dt_material d; memset(&d, 0, sizeof d);
        d.type = (int)m.type;                                           // same enum order (dorktracer.h)
        d.brdf = index_of(sc.brdfs, m.brdf);
        put3(d.ambient, m.ambient); put3(d.diffuse, m.diffuse); put3(d.specular, m.specular); put3(d.mirror, m.mirror);
        d.phong_exponent = m.phong_exponent; d.refractive_index = m.refractiveIndex;
        put3(d.absorption_coefficient, m.absorptionCoefficient);
        d.conductor_absorption_index = m.conductorAbsorptionIndex; d.roughness = m.roughness;
        put3(d.radiance, m.radiance);
        f.materials.push_back(d);
    }
    // ---- images (LDRImage.h / HDRImage.h) and textures (imageTexture.h, perlinTexture.h)
    for (Image* im : sc.images) {
        dt_image d; memset(&d, 0, sizeof d);
        d.width = im->width; d.height = im->height;
        if (HDRImage* h = dynamic_cast<HDRImage*>(im)) { d.channels = 3; d.is_hdr = 1; d.data = h->src.data(); }
        else { LDRImage* l = (LDRImage*)im; d.channels = l->channels; d.is_hdr = 0; d.data = l->image; }
        if (!d.data) { fprintf(stderr, "dt_flatten_scene: image %d failed to load\n", im->id); return false; }
        f.images.push_back(d);
    }
    for (Texture* t : sc.textures) {
        dt_texture d; memset(&d, 0, sizeof d);
        d.image = -1;
        // DecalMode is not kept by the Texture object (texture.h:33-38 folds it into type / operationMode); recover it
        switch (t->type) {
            case Texture::Textures::Diffuse: d.decal_mode = t->operationMode == Texture::Blend ? DT_DECAL_BLEND_KD : DT_DECAL_REPLACE_KD; break;
            case Texture::Textures::Specular: d.decal_mode = DT_DECAL_REPLACE_KS; break;
            case Texture::Textures::Normal: d.decal_mode = DT_DECAL_REPLACE_NORMAL; break;
            case Texture::Textures::Bump: d.decal_mode = DT_DECAL_BUMP_NORMAL; break;
            default: d.decal_mode = DT_DECAL_REPLACE_ALL; break;
        }
        if (t == sc.bgTexture) d.decal_mode = DT_DECAL_REPLACE_BG;      // "replace_background" leaves `type` unset (texture.h:57-79)
        if (ImageTexture* it = dynamic_cast<ImageTexture*>(t)) {
            d.kind = DT_TEX_IMAGE;
            d.image = index_of(sc.images, it->img);
            d.interpolation = it->interpolationMode == ImageTexture::Nearest ? DT_INTERP_NEAREST : DT_INTERP_BILINEAR;
            d.normalizer = it->normalizer; d.sample_multiplier = it->sampleMultiplier;
        } else {
            PerlinTexture* pt = (PerlinTexture*)t;
            d.kind = DT_TEX_PERLIN;
            d.normalizer = 1.0f; d.sample_multiplier = pt->bumpFactor; d.noise_scale = pt->scale;
            d.noise_conversion = pt->conversionType == PerlinTexture::AbsoluteVal ? DT_NOISE_ABSVAL : DT_NOISE_LINEAR;
        }
        f.textures.push_back(d);
    }
    // ---- lights
    for (PointLight& l : sc.point_lights) { dt_point_light d; put3(d.position, l.position); put3(d.intensity, l.intensity); f.point_lights.push_back(d); }
    for (AreaLight* l : sc.areaLights) {
        dt_area_light d; put3(d.position, l->position); put3(d.normal, l->normal); put3(d.radiance, l->radiance); d.extent = l->extent;
        put3(d.u, l->u); put3(d.v, l->v);
        f.area_lights.push_back(d);
    }
    for (DirectionalLight* l : sc.directionalLights) { dt_directional_light d; put3(d.dir, l->dir); put3(d.radiance, l->radiance); f.directional_lights.push_back(d); }
    for (SpotLight* l : sc.spotLights) {
        dt_spot_light d; put3(d.pos, l->pos); put3(d.dir, l->dir); put3(d.intensity, l->intensity);
        d.coverage_angle = l->coverageAngle; d.falloff_angle = l->falloffAngle;
        d.cos_half_falloff = l->cosHalfFalloff; d.cos_half_coverage = l->cosHalfCoverage;
        f.spot_lights.push_back(d);
    }
    for (SphericalEnvironmentLight* l : sc.sphericalEnvLights) {
        dt_env_light d; d.image = index_of(sc.images, l->image);
        if (d.image < 0) { fprintf(stderr, "dt_flatten_scene: environment light without image\n"); return false; }
        f.env_lights.push_back(d);
    }

    // ---- shapes in the scan order of Raytracer::IntersectObjects (raytracer.cpp:625-643) = the tie-break order:
    // scene.meshes (Mesh, LightMesh, MeshInstance, Triangle in push order), then scene.spheres
    std::map<const Mesh*, int> shape_of_mesh;
    f.verts.reserve(sc.meshes.size()); f.uvs.reserve(sc.meshes.size()); f.faces.reserve(sc.meshes.size()); f.nodes.reserve(sc.meshes.size());
    for (size_t si = 0; si < sc.meshes.size(); si++) {
        Shape* s = sc.meshes[si];
        dt_shape d; memset(&d, 0, sizeof d);
        fill_shape_common(sc, s, d);
        if (s->isInstance) {
            InstancedMesh* im = (InstancedMesh*)s;
            d.kind = DT_SHAPE_INSTANCE;
            d.material = im->material_id;                              // InstancedMesh shadows Shape::material_id (instancedMesh.hpp:22)
            std::map<const Mesh*, int>::iterator it = shape_of_mesh.find(im->baseMesh);
            if (it == shape_of_mesh.end()) { fprintf(stderr, "dt_flatten_scene: instance of an unknown base mesh\n"); return false; }
            d.base_shape = it->second;
            put3(d.bbox_min, im->bbox.minCorner); put3(d.bbox_max, im->bbox.maxCorner);
        } else {
            Mesh* m = (Mesh*)s;
            d.kind = DT_SHAPE_MESH;
            d.material = m->GetMaterial();
            dt_mesh dm; memset(&dm, 0, sizeof dm);
            f.verts.emplace_back(); f.uvs.emplace_back(); f.faces.emplace_back(); f.nodes.emplace_back();
            std::vector<float>& V = f.verts.back();
            V.reserve(m->vertices.size() * 3);
            for (const Vec3f& v : m->vertices) { V.push_back(v.x); V.push_back(v.y); V.push_back(v.z); }
            std::vector<float>& U = f.uvs.back();
            for (const Vec2f& v : m->uv) { U.push_back(v.x); U.push_back(v.y); }
            std::vector<dt_face>& F = f.faces.back();
            F.reserve(m->faces.size());
            for (const Face& fc : m->faces) {
                dt_face df; df.v0_id = fc.v0_id; df.v1_id = fc.v1_id; df.v2_id = fc.v2_id;
                df.n[0] = fc.n.x; df.n[1] = fc.n.y; df.n[2] = fc.n.z; df.area = fc.area;
                F.push_back(df);
            }
            // Mesh::bvh: raw child pointers into the same vector -> indices (bvh.hpp:16-20).  Only the nodes RecursiveBVHBuild
            // handed out are meaningful; the tail of the 2n-1 allocation (mesh.cpp:29) is never referenced.
            std::vector<dt_bvh2_node>& N = f.nodes.back();
            const size_t n_used = m->nextFreeNodeIdx > 0 ? (size_t)m->nextFreeNodeIdx : (m->bvh.empty() ? 0 : 1);
            N.reserve(n_used);
            for (size_t k = 0; k < n_used && k < m->bvh.size(); k++) {
                const BVH& b = m->bvh[k];
                dt_bvh2_node dn;
                dn.bmin[0] = b.bbox.minCorner.x; dn.bmin[1] = b.bbox.minCorner.y; dn.bmin[2] = b.bbox.minCorner.z;
                dn.bmax[0] = b.bbox.maxCorner.x; dn.bmax[1] = b.bbox.maxCorner.y; dn.bmax[2] = b.bbox.maxCorner.z;
                dn.left = b.left ? (int32_t)(b.left - &m->bvh[0]) : -1;
                dn.right = b.right ? (int32_t)(b.right - &m->bvh[0]) : -1;
                dn.first_face = b.firstFace; dn.face_count = b.faceCount;
                N.push_back(dn);
            }
            dm.vertices = V.data(); dm.n_vertices = (int32_t)(V.size() / 3);
            dm.uvs = U.data(); dm.n_uvs = (int32_t)(U.size() / 2);
            dm.vertex_offset = m->vertexOffset; dm.texture_offset = m->textureOffset;
            dm.faces = F.data(); dm.n_faces = (int32_t)F.size();
            dm.bvh = N.data(); dm.n_bvh_nodes = (int32_t)N.size();
            put3(dm.bbox_min, m->bbox.minCorner); put3(dm.bbox_max, m->bbox.maxCorner);
            dm.surface_area = m->surfaceArea;
            d.mesh = (int32_t)f.meshes.size();
            shape_of_mesh[m] = (int)si;
            f.meshes.push_back(dm);
            if (MeshLight* ml = dynamic_cast<MeshLight*>(m)) {          // LightMesh: also a sampled light (meshLight.h)
                dt_mesh_light l; l.shape = (int32_t)si; l.id = ml->id; put3(l.radiance, ml->radiance);
                f.mesh_lights.push_back(l);
            }
        }
        f.shapes.push_back(d);
    }
    // scene.meshLights order is the sampling order of SampleDirectLighting (raytracer.cpp:777); it equals the push order above
    for (Sphere* s : sc.spheres) {
        dt_shape d; memset(&d, 0, sizeof d);
        fill_shape_common(sc, s, d);
        d.kind = DT_SHAPE_SPHERE;
        d.material = s->material_id;                                   // Sphere shadows Shape::material_id too (sphere.hpp:13)
        put3(d.center, s->vertex_data[s->center_vertex_id - 1]);
        d.radius = s->radius;
        f.shapes.push_back(d);
    }
    D.materials = f.materials.data(); D.n_materials = (int32_t)f.materials.size();
    D.brdfs = f.brdfs.data(); D.n_brdfs = (int32_t)f.brdfs.size();
    D.point_lights = f.point_lights.data(); D.n_point_lights = (int32_t)f.point_lights.size();
    D.area_lights = f.area_lights.data(); D.n_area_lights = (int32_t)f.area_lights.size();
    D.directional_lights = f.directional_lights.data(); D.n_directional_lights = (int32_t)f.directional_lights.size();
    D.spot_lights = f.spot_lights.data(); D.n_spot_lights = (int32_t)f.spot_lights.size();
    D.env_lights = f.env_lights.data(); D.n_env_lights = (int32_t)f.env_lights.size();
    D.mesh_lights = f.mesh_lights.data(); D.n_mesh_lights = (int32_t)f.mesh_lights.size();
    D.images = f.images.data(); D.n_images = (int32_t)f.images.size();
    D.textures = f.textures.data(); D.n_textures = (int32_t)f.textures.size();
    D.meshes = f.meshes.data(); D.n_meshes = (int32_t)f.meshes.size();
    D.shapes = f.shapes.data(); D.n_shapes = (int32_t)f.shapes.size();
    D.n_mesh_shapes = (int32_t)sc.meshes.size();
    return true;
}

void fill_camera(Camera& cam, dt_camera_desc& c) {
    memset(&c, 0, sizeof c);
    put3(c.position, cam.position); put3(c.gaze, cam.gaze); put3(c.up, cam.up); put3(c.right, cam.right);
    put3(c.q, cam.m_q);
    c.left = cam.m_left; c.right_ = cam.m_right; c.bottom = cam.m_bottom; c.top = cam.m_top;
    c.near_dist = cam.nearDist;
    c.width = cam.imageWidth; c.height = cam.imageHeight;
    c.samples_per_pixel = cam.samplesPerPixel;
    c.focus_distance = cam.focusDistance; c.aperture_size = cam.apertureSize;
    RendererParams& rp = cam.GetRendererParams();
    c.path_tracing = rp.pathTracingEnabled ? 1 : 0;
    c.importance_sampling = rp.sampleImportance ? 1 : 0;
    c.next_event_estimation = rp.nextEventEstimationEnabled ? 1 : 0;
    c.russian_roulette = rp.russianRouletteEnabled ? 1 : 0;
    c.has_tonemapper = cam.hasTonemapper ? 1 : 0;
    if (cam.hasTonemapper && cam.tonemapper) {
        c.tm_key = cam.tonemapper->keyValue; c.tm_burn = cam.tonemapper->burnPerct;
        c.tm_saturation = cam.tonemapper->saturation; c.tm_gamma = cam.tonemapper->gamma;
    }
}

}  // namespace

// After Scene::loadFromXml: flatten once, create the device scene(s).  DT_GPUS=n (n > 1) renders every frame on n GPUs of
// the box from this ONE process (dt_multi_*: scene replicated, strips of the image dealt round-robin, peer-memory gather).
void* dt_dropin_create(Scene& scene) {
    DtDropin* d = new DtDropin();
    if (!flatten(scene, d->flat)) { delete d; return nullptr; }
    const char* e = getenv("DT_GPUS");
    const int n_gpus = e ? atoi(e) : 1;
    int rc;
    if (n_gpus > 1) rc = dt_multi_create(&d->flat.desc, n_gpus, &d->multi);
    else { rc = dt_gpu_init(0); if (rc >= 0) rc = dt_scene_create(&d->flat.desc, &d->gpu); }
    if (rc < 0) { fprintf(stderr, "dorktracer: %s\n", dt_last_error()); delete d; return nullptr; }
    return d;
}

// One camera = one frame: replaces main.cpp:164-185 (thread spawn / join over renderThreadMain) and main.cpp:190 (tonemap).
// image: W*H*3 bytes, hdrImage: W*H*3 floats or NULL -- the buffers main.cpp allocates (main.cpp:146-152).
int dt_dropin_render(void* handle, Camera& cam, unsigned char* image, float* hdrImage) {
    DtDropin* d = (DtDropin*)handle;
    dt_camera_desc c;
    fill_camera(cam, c);
    dt_render_params p; memset(&p, 0, sizeof p);
    p.seed = 1234; p.tile_world = 1;
    if (const char* e = getenv("DT_RENDER_FLAGS")) p.flags = atoi(e);        // DT_FLAG_* (dorktracer.h), e.g. 4096 = follow zero-weight paths as the CPU loop does
    dt_stats st;
    const int rc = d->multi ? dt_multi_render(d->multi, &c, &p, image, hdrImage, &st) : dt_render(d->gpu, &c, &p, image, hdrImage, &st);
    if (rc != DT_OK) { fprintf(stderr, "dorktracer: %s\n", dt_last_error()); return rc; }
    printf("DT_RAYS closest=%llu shadow=%llu\n", (unsigned long long)st.rays_closest, (unsigned long long)st.rays_shadow);
    printf("GPU frame: %.3f ms, %u kernel launches, %u waves\n", st.ms_total, st.kernel_launches, st.waves);
    return 0;
}

void dt_dropin_destroy(void* handle) {
    DtDropin* d = (DtDropin*)handle;
    if (!d) return;
    if (d->multi) dt_multi_destroy(d->multi);
    if (d->gpu) dt_scene_destroy(d->gpu);
    delete d;
}
