"""Fixture helpers: load the committed golden .npz files and materialise their scene XML."""
import os
import tempfile

import numpy as np

from dtb200.scene import HostScene

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PINS = ["simple", "two_spheres", "spheres", "spheres_mirror", "cornellbox_recursive_conductors", "scienceTree"]
DIELECTRIC = ["scienceTree_diamond", "cornellbox_recursive_alt2"]

_tmp = tempfile.mkdtemp(prefix="dt_golden_")


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def golden_scene(name):
    g = load_golden(name)
    path = os.path.join(_tmp, name + ".xml")
    if not os.path.exists(path):
        with open(path, "wb") as f:
            f.write(g["xml"].tobytes())
    return HostScene(path), g


BRDF_SCENE_XML = """<Scene><MaxRecursionDepth>2</MaxRecursionDepth><BackgroundColor>10 10 20</BackgroundColor><ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
<Cameras><Camera id="1"><Position>0 3 14</Position><Gaze>0 -0.15 -1</Gaze><Up>0 1 0</Up><NearPlane>-1 1 -0.5 0.5</NearPlane>
<NearDistance>1.6</NearDistance><ImageResolution>%d %d</ImageResolution><ImageName>brdf.png</ImageName></Camera></Cameras>
<Lights><AmbientLight>15 15 15</AmbientLight><PointLight id="1"><Position>4 9 9</Position><Intensity>22000 21000 20000</Intensity></PointLight>
<PointLight id="2"><Position>-7 4 6</Position><Intensity>9000 10000 12000</Intensity></PointLight></Lights>
<BRDFs><OriginalPhong id="1"><Exponent>25</Exponent></OriginalPhong><OriginalBlinnPhong id="2"><Exponent>40</Exponent></OriginalBlinnPhong>
<ModifiedPhong id="3"><Exponent>25</Exponent></ModifiedPhong><ModifiedPhong id="4" normalized="true"><Exponent>25</Exponent></ModifiedPhong>
<ModifiedBlinnPhong id="5"><Exponent>40</Exponent></ModifiedBlinnPhong><ModifiedBlinnPhong id="6" normalized="true"><Exponent>40</Exponent></ModifiedBlinnPhong>
<TorranceSparrow id="7"><Exponent>50</Exponent></TorranceSparrow><TorranceSparrow id="8" kdfresnel="true"><Exponent>50</Exponent></TorranceSparrow></BRDFs>
<Materials>
<Material id="1"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>0.5 0.5 0.5</DiffuseReflectance><SpecularReflectance>0.1 0.1 0.1</SpecularReflectance><PhongExponent>5</PhongExponent></Material>
%s</Materials>
<VertexData>-12 0 -8
12 0 -8
12 0 8
-12 0 8
%s</VertexData>
<Objects><Mesh id="1"><Material>1</Material><Faces>1 3 2
1 4 3</Faces></Mesh>
%s</Objects></Scene>"""


def brdf_scene(path, width=480, height=240):
    """Eight spheres, one per BRDF variant (brdf*.cpp: Phong, Blinn-Phong, modified Phong / Blinn-Phong with and
    without normalisation, Torrance-Sparrow with and without kdfresnel), two point lights, no sampling."""
    mats, verts, objs = "", "", ""
    for k in range(8):
        mats += ('<Material id="%d" BRDF="%d"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>%g %g %g</DiffuseReflectance>'
                 '<SpecularReflectance>0.5 0.5 0.5</SpecularReflectance><PhongExponent>30</PhongExponent><RefractionIndex>1.7</RefractionIndex></Material>\n'
                 % (k + 2, k + 1, 0.25 + 0.05 * k, 0.45 - 0.03 * k, 0.2 + 0.04 * (k % 3)))
        x, z, y = -7.5 + 2.15 * k, (-1.5 if k % 2 else 1.5), 1.0
        verts += "%g %g %g\n" % (x, y, z)
        objs += '<Sphere id="%d"><Material>%d</Material><Center>%d</Center><Radius>1</Radius></Sphere>\n' % (k + 1, k + 2, k + 5)
    with open(path, "w") as f:
        f.write(BRDF_SCENE_XML % (width, height, mats, verts, objs))
    return path


def blur_dof_scene(path, width=240, height=160, spp=100):
    """Motion-blurred sphere and mesh (Shape::motionBlurVector), a rough mirror (Material::roughness), and a thin-lens
    camera (ApertureSize / FocusDistance): the sampling paths of GenerateRay / Reflect that the deterministic scenes do
    not reach.  Point lights only; the randomness is the per-sample time / lens / roughness."""
    xml = """<Scene><MaxRecursionDepth>2</MaxRecursionDepth><BackgroundColor>20 30 50</BackgroundColor><ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
<Cameras><Camera id="1"><Position>0 2.5 12</Position><Gaze>0 -0.12 -1</Gaze><Up>0 1 0</Up><NearPlane>-1 1 -0.6667 0.6667</NearPlane>
<NearDistance>2</NearDistance><ImageResolution>%d %d</ImageResolution><NumSamples>%d</NumSamples><FocusDistance>11</FocusDistance><ApertureSize>0.35</ApertureSize>
<ImageName>blur.png</ImageName></Camera></Cameras>
<Lights><AmbientLight>25 25 25</AmbientLight><PointLight id="1"><Position>5 9 8</Position><Intensity>20000 20000 19000</Intensity></PointLight></Lights>
<Materials><Material id="1"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>0.55 0.5 0.45</DiffuseReflectance><SpecularReflectance>0.2 0.2 0.2</SpecularReflectance><PhongExponent>20</PhongExponent></Material>
<Material id="2"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>0.7 0.2 0.15</DiffuseReflectance><SpecularReflectance>0.4 0.4 0.4</SpecularReflectance><PhongExponent>40</PhongExponent></Material>
<Material id="3" type="mirror"><AmbientReflectance>0 0 0</AmbientReflectance><DiffuseReflectance>0.05 0.05 0.05</DiffuseReflectance><SpecularReflectance>0 0 0</SpecularReflectance><MirrorReflectance>0.85 0.85 0.9</MirrorReflectance><Roughness>0.15</Roughness></Material>
<Material id="4"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>0.15 0.5 0.25</DiffuseReflectance><SpecularReflectance>0.1 0.1 0.1</SpecularReflectance><PhongExponent>8</PhongExponent></Material></Materials>
<VertexData>-10 0 -10
10 0 -10
10 0 10
-10 0 10
-3 1.2 0
0.3 1.2 1.5
3.2 1.4 -1
1 0.01 3
3 0.01 3
2 2.2 3</VertexData>
<Objects><Mesh id="1"><Material>1</Material><Faces>1 3 2
1 4 3</Faces></Mesh>
<Mesh id="2"><Material>4</Material><MotionBlur>0.8 0 0</MotionBlur><Faces>8 9 10</Faces></Mesh>
<Sphere id="1"><Material>2</Material><Center>5</Center><Radius>1.2</Radius><MotionBlur>0 0.9 0.4</MotionBlur></Sphere>
<Sphere id="2"><Material>3</Material><Center>6</Center><Radius>1.2</Radius></Sphere>
<Sphere id="3"><Material>2</Material><Center>7</Center><Radius>1.4</Radius></Sphere></Objects></Scene>""" % (width, height, spp)
    with open(path, "w") as f:
        f.write(xml)
    return path


def glass_closeup_scene(path, width=160, height=160, depth=4):
    """A dielectric sphere filling the whole frame: every camera ray spawns a reflected and a refracted child
    (raytracer.cpp:316-411), so wave k+1 holds up to twice the rays of wave k -- the fan-out the wavefront queues are sized for."""
    xml = """<Scene><MaxRecursionDepth>%d</MaxRecursionDepth><BackgroundColor>30 40 70</BackgroundColor><ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
<Cameras><Camera id="1"><Position>0 0 3</Position><Gaze>0 0 -1</Gaze><Up>0 1 0</Up><NearPlane>-0.3 0.3 -0.3 0.3</NearPlane>
<NearDistance>1</NearDistance><ImageResolution>%d %d</ImageResolution><ImageName>glass.png</ImageName></Camera></Cameras>
<Lights><AmbientLight>20 20 20</AmbientLight><PointLight id="1"><Position>4 6 6</Position><Intensity>9000 9000 9000</Intensity></PointLight>
<PointLight id="2"><Position>-5 3 4</Position><Intensity>4000 5000 6000</Intensity></PointLight></Lights>
<Materials><Material id="1" type="dielectric"><AmbientReflectance>0 0 0</AmbientReflectance><DiffuseReflectance>0 0 0</DiffuseReflectance><SpecularReflectance>0 0 0</SpecularReflectance>
<AbsorptionCoefficient>0.05 0.01 0.01</AbsorptionCoefficient><RefractionIndex>1.5</RefractionIndex></Material>
<Material id="2"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>0.6 0.5 0.3</DiffuseReflectance><SpecularReflectance>0.2 0.2 0.2</SpecularReflectance><PhongExponent>10</PhongExponent></Material></Materials>
<VertexData>0 0 0
-8 -2 -8
8 -2 -8
8 -2 8
-8 -2 8</VertexData>
<Objects><Mesh id="1"><Material>2</Material><Faces>2 4 3
2 5 4</Faces></Mesh>
<Sphere id="1"><Material>1</Material><Center>1</Center><Radius>1.2</Radius></Sphere></Objects></Scene>""" % (depth, width, height)
    with open(path, "w") as f:
        f.write(xml)
    return path


def quirks_scene(out_dir, width=200, height=136, tonemap=False):
    """Deterministic scene over reference behaviours no other scene renders: a `replace_background` image texture
    behind everything (raytracer.cpp:49-62), a `replace_ks` texture -- whose lookup reads the shape's DIFFUSE texture slot
    (raytracer.cpp:516-531; a shape with a replace_ks texture and NO diffuse one makes the reference dereference nullptr, so the
    sphere carries both) -- next to `replace_kd` / `blend_kd` ones, a `degamma="true"` material (parser.cpp:1154-1210), and a light strong
    enough to push radiance past 2^31, where the reference's `(int)` conversion yields INT_MIN and the clamp stores 0, not 255
    (helperMath.cpp:140-152).  tonemap=True attaches a photographic tonemapper instead of the LDR clamp."""
    import numpy as np
    from dtb200 import scenegen
    os.makedirs(os.path.join(out_dir, "inputs"), exist_ok=True)
    yy, xx = np.mgrid[0:96, 0:128]
    bg = np.stack([(40 + xx).clip(0, 255), (30 + 2 * yy).clip(0, 255), (90 + xx // 2 + yy // 2).clip(0, 255)], axis=-1).astype(np.uint8)
    scenegen.write_png(os.path.join(out_dir, "inputs", "bg.png"), bg)
    chk = (((xx // 16) + (yy // 16)) % 2).astype(np.uint8)
    tex = np.stack([80 + 150 * chk, 200 - 120 * chk, 60 + 40 * chk], axis=-1).astype(np.uint8)
    scenegen.write_png(os.path.join(out_dir, "inputs", "chk.png"), tex)
    tm = ("<Tonemap><TMO>Photographic</TMO><TMOOptions>0.18 1</TMOOptions><Saturation>1.0</Saturation><Gamma>2.2</Gamma></Tonemap>" if tonemap else "")
    xml = """<Scene><MaxRecursionDepth>2</MaxRecursionDepth><BackgroundColor>0 0 0</BackgroundColor><ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
<Cameras><Camera id="1"><Position>0 2.5 11</Position><Gaze>0 -0.15 -1</Gaze><Up>0 1 0</Up><NearPlane>-1 1 -0.68 0.68</NearPlane>
<NearDistance>1.8</NearDistance><ImageResolution>%d %d</ImageResolution>%s<ImageName>quirks.png</ImageName></Camera></Cameras>
<Lights><AmbientLight>20 20 20</AmbientLight><PointLight id="1"><Position>3 8 7</Position><Intensity>30000 29000 27000</Intensity></PointLight>
<PointLight id="2"><Position>-3.2 0.65 2.2</Position><Intensity>4e13 4e13 4e13</Intensity></PointLight></Lights>
<Materials>
<Material id="1"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>0.5 0.5 0.45</DiffuseReflectance><SpecularReflectance>0.3 0.3 0.3</SpecularReflectance><PhongExponent>12</PhongExponent></Material>
<Material id="2" degamma="true"><AmbientReflectance>0.8 0.8 0.8</AmbientReflectance><DiffuseReflectance>0.7 0.35 0.2</DiffuseReflectance><SpecularReflectance>0.5 0.5 0.5</SpecularReflectance><PhongExponent>30</PhongExponent></Material>
<Material id="3"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>0.4 0.4 0.4</DiffuseReflectance><SpecularReflectance>0.6 0.6 0.6</SpecularReflectance><PhongExponent>40</PhongExponent></Material>
<Material id="4" type="mirror"><AmbientReflectance>0 0 0</AmbientReflectance><DiffuseReflectance>0.1 0.1 0.1</DiffuseReflectance><SpecularReflectance>0 0 0</SpecularReflectance><MirrorReflectance>0.8 0.8 0.8</MirrorReflectance></Material>
</Materials>
<Textures><Images><Image id="1">bg.png</Image><Image id="2">chk.png</Image></Images>
<TextureMap id="1" type="image"><ImageId>1</ImageId><DecalMode>replace_background</DecalMode><Interpolation>bilinear</Interpolation></TextureMap>
<TextureMap id="2" type="image"><ImageId>2</ImageId><DecalMode>replace_ks</DecalMode><Interpolation>nearest</Interpolation></TextureMap>
<TextureMap id="3" type="image"><ImageId>2</ImageId><DecalMode>replace_kd</DecalMode><Interpolation>bilinear</Interpolation></TextureMap>
<TextureMap id="4" type="image"><ImageId>2</ImageId><DecalMode>blend_kd</DecalMode><Interpolation>nearest</Interpolation></TextureMap>
</Textures>
<VertexData>-9 0 -6
9 0 -6
9 0 6
-9 0 6
-3.2 1.1 0
0 1.1 1
3.2 1.1 0
0 3.4 -2.5</VertexData>
<TexCoordData>0 0
3 0
3 2
0 2
0 0
0 0
0 0
0 0</TexCoordData>
<Objects><Mesh id="1"><Material>1</Material><Textures>4</Textures><Faces>1 3 2
1 4 3</Faces></Mesh>
<Sphere id="1"><Material>2</Material><Center>5</Center><Radius>1.1</Radius></Sphere>
<Sphere id="2"><Material>3</Material><Textures>2 3</Textures><Center>6</Center><Radius>1.1</Radius></Sphere>
<Sphere id="3"><Material>3</Material><Textures>3</Textures><Center>7</Center><Radius>1.1</Radius></Sphere>
<Sphere id="4"><Material>4</Material><Center>8</Center><Radius>1.3</Radius></Sphere></Objects></Scene>""" % (width, height, tm)
    path = os.path.join(out_dir, "quirks.xml")
    with open(path, "w") as f:
        f.write(xml)
    return path


def env_whitted_scene(out_dir, width=160, height=112, spp=4, blur_instance=False):
    """Whitted (no path tracing) with a spherical environment light: mirror and dielectric children that MISS read the map
    (raytracer.cpp:351-356,404-409,461-469; the refracted child along the REFLECTED direction, :408), a conductor child that
    misses stays black (:247), camera rays that miss read it too (:49-62); the light is sampled by rejection at every lit hit
    (sphericalEnvironmentLight.h:37-65), so the scene draws random numbers although it has no `Renderer` element.  Plus a
    MeshInstance with composed transforms and a motion-blurred mesh.
    blur_instance: the instance moves too (instancedMesh.cpp:21-38).  NOT exactly comparable with the reference: when the ray
    misses the instance's box the reference leaves ray.origin SHIFTED by motionBlurVector * time for every shape scanned after
    it (instancedMesh.cpp:22-27 restores the origin only inside the `if`), which displaces all later shapes of the scan order;
    neither the oracle nor the GPU path reproduces that leak (documented deviations)."""
    import numpy as np
    from dtb200 import scenegen
    os.makedirs(os.path.join(out_dir, "inputs"), exist_ok=True)
    yy, xx = np.mgrid[0:32, 0:64].astype(np.float32)                  # a smooth sky (no hot spot: the estimator stays low-variance)
    sky = 8.0 + 14.0 * (1 - yy / 32) + 4.0 * np.sin(xx / 64 * 2 * np.pi)
    scenegen.write_exr(os.path.join(out_dir, "inputs", "env.exr"), np.stack([0.7 * sky, 0.85 * sky, 1.0 * sky], axis=-1).astype(np.float32))
    verts, faces = scenegen.blob_mesh(24, 13, 0.9, (0.0, 0.0, 0.0))
    scenegen.write_ply(os.path.join(out_dir, "base.ply"), verts, faces)
    xml = """<Scene><MaxRecursionDepth>3</MaxRecursionDepth><BackgroundColor>0 0 0</BackgroundColor><ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
<Cameras><Camera id="1"><Position>0 2.2 10</Position><Gaze>0 -0.1 -1</Gaze><Up>0 1 0</Up><NearPlane>-1 1 -0.7 0.7</NearPlane>
<NearDistance>1.7</NearDistance><ImageResolution>%d %d</ImageResolution><NumSamples>%d</NumSamples><ImageName>envw.png</ImageName></Camera></Cameras>
<Lights><AmbientLight>10 10 10</AmbientLight><PointLight id="1"><Position>4 7 6</Position><Intensity>9000 9000 8500</Intensity></PointLight>
<SphericalDirectionalLight id="2"><ImageId>1</ImageId></SphericalDirectionalLight></Lights>
<Materials>
<Material id="1"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>0.5 0.45 0.4</DiffuseReflectance><SpecularReflectance>0.2 0.2 0.2</SpecularReflectance><PhongExponent>10</PhongExponent></Material>
<Material id="2" type="mirror"><AmbientReflectance>0 0 0</AmbientReflectance><DiffuseReflectance>0.05 0.05 0.05</DiffuseReflectance><SpecularReflectance>0 0 0</SpecularReflectance><MirrorReflectance>0.85 0.85 0.85</MirrorReflectance></Material>
<Material id="3" type="dielectric"><AmbientReflectance>0 0 0</AmbientReflectance><DiffuseReflectance>0 0 0</DiffuseReflectance><SpecularReflectance>0 0 0</SpecularReflectance><AbsorptionCoefficient>0.02 0.01 0.01</AbsorptionCoefficient><RefractionIndex>1.5</RefractionIndex></Material>
<Material id="4" type="conductor"><AmbientReflectance>0 0 0</AmbientReflectance><DiffuseReflectance>0.1 0.1 0.05</DiffuseReflectance><SpecularReflectance>0 0 0</SpecularReflectance><MirrorReflectance>0.9 0.8 0.5</MirrorReflectance><RefractionIndex>0.4</RefractionIndex><AbsorptionIndex>2.8</AbsorptionIndex></Material>
<Material id="5"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>0.25 0.5 0.3</DiffuseReflectance><SpecularReflectance>0.3 0.3 0.3</SpecularReflectance><PhongExponent>25</PhongExponent></Material>
</Materials>
<Textures><Images><Image id="1">env.exr</Image></Images></Textures>
<VertexData>-7 0 -5
7 0 -5
7 0 5
-7 0 5
-3.3 1.1 0
-1.1 1.1 0.8
1.1 1.1 0.8
-1 0.01 3
1 0.01 3
0 1.4 3</VertexData>
<Transformations><Translation id="1">3.3 1.0 0</Translation><Translation id="2">0 3.0 -2</Translation><Scaling id="1">1 1.2 1</Scaling><Rotation id="1">35 0 1 0</Rotation></Transformations>
<Objects><Mesh id="1"><Material>1</Material><Faces>1 3 2
1 4 3</Faces></Mesh>
<Mesh id="2"><Material>5</Material><Transformations>t2</Transformations><Faces plyFile="base.ply" /></Mesh>
<Mesh id="3"><Material>5</Material><MotionBlur>0.6 0 0</MotionBlur><Faces>8 9 10</Faces></Mesh>
<MeshInstance id="4" baseMeshId="2" resetTransform="true"><Material>5</Material><Transformations>s1 r1 t1</Transformations>%s</MeshInstance>
<Sphere id="1"><Material>2</Material><Center>5</Center><Radius>1.1</Radius></Sphere>
<Sphere id="2"><Material>3</Material><Center>6</Center><Radius>1.1</Radius></Sphere>
<Sphere id="3"><Material>4</Material><Center>7</Center><Radius>1.1</Radius></Sphere></Objects></Scene>""" % (width, height, spp, "<MotionBlur>0 0.5 0.3</MotionBlur>" if blur_instance else "")
    path = os.path.join(out_dir, "envw.xml")
    with open(path, "w") as f:
        f.write(xml)
    return path


def random_scene(out_dir, seed, width=112, height=80, textures=False, extras=False, mc=False, spp=None):
    """A seeded random DETERMINISTIC scene (no sampling anywhere, so the reference, the oracle and the GPU path must agree bit
    for bit on hits and ray counts): a floor, 2-5 spheres and 1-3 small meshes (boxes, tetrahedra, triangle soups, single
    <Triangle>s) with random materials -- plain with any of the eight BRDF variants, mirror, conductor, dielectric --, random
    composed transformations on some of them, a MeshInstance of the first mesh (with and without resetTransform), 1-2 point
    lights and, by chance, a directional and a spot light; recursion depth 1-4 and a random camera.
    textures=True: random texture coordinates on every vertex and, on about half of the shapes, one or two of seven texture maps
    (image replace_kd / blend_kd / replace_all, image normal map, image bump map, Perlin replace_kd, Perlin bump map).
    extras=True (a different random stream): lookAt cameras (GazePoint / FovY), a photographic tonemapper on a third of the frames,
    `degamma` materials, recursion depth up to 6, a binary-PLY blob mesh with a MeshInstance of it and, with textures, a
    `replace_background` map and `replace_ks` maps (always next to a diffuse map: without one the reference dereferences nullptr).
    mc=True (another stream; NOT deterministic): several samples per pixel, by chance a thin lens, the path-tracing renderer with a
    random subset of ImportanceSampling / NextEventEstimation / RussianRoulette, an area light, a spherical environment light, a
    LightMesh, motion blur on a sphere and rough mirrors; spp overrides the sample count."""
    rng = np.random.RandomState(1000 + seed + (50000 if extras else 0) + (90000 if mc else 0))
    os.makedirs(os.path.join(out_dir, "inputs"), exist_ok=True)
    u = lambda a, b: float(rng.uniform(a, b))
    f3 = lambda v: "%.6g %.6g %.6g" % (v[0], v[1], v[2])
    cam_pos = (u(-2.5, 2.5), u(1.5, 4.0), u(6.5, 9))
    gaze = (-cam_pos[0] * 0.08 + u(-0.05, 0.05), -0.12 + u(-0.08, 0.05), -1.0)
    xml = ("<Scene><MaxRecursionDepth>%d</MaxRecursionDepth><BackgroundColor>%d %d %d</BackgroundColor><ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
           % (rng.randint(1, 7 if extras else 5), rng.randint(20, 120), rng.randint(20, 120), rng.randint(30, 160)))
    tm = ""
    if extras and rng.rand() < 0.34:
        tm = ("<Tonemap><TMO>Photographic</TMO><TMOOptions>%.4g %.4g</TMOOptions><Saturation>%.4g</Saturation><Gamma>%.4g</Gamma></Tonemap>"
              % (u(0.1, 0.3), u(0.5, 3), u(0.6, 1.2), u(1.8, 2.4)))
    env_light = False
    if mc:
        n_spp = [1, 4, 4, 9][rng.randint(4)]
        if spp is not None:
            n_spp = spp                              # (drawn first: the rest of the scene does not depend on the override)
        tm += "<NumSamples>%d</NumSamples>" % n_spp
        if rng.rand() < 0.3:
            tm += "<FocusDistance>%.4g</FocusDistance><ApertureSize>%.4g</ApertureSize>" % (u(6, 10), u(0.05, 0.3))
        if rng.rand() < 0.75:
            params = [w for w in ("ImportanceSampling", "NextEventEstimation", "RussianRoulette") if rng.rand() < 0.6]
            tm += "<Renderer>PathTracing</Renderer><RendererParams>%s</RendererParams>" % " ".join(params)
        env_light = rng.rand() < 0.4
    if extras and rng.rand() < 0.5:
        xml += ("<Cameras><Camera id=\"1\" type=\"lookAt\"><Position>%s</Position><GazePoint>%s</GazePoint><Up>%s</Up><FovY>%.4g</FovY>"
                "<NearDistance>%.4g</NearDistance><ImageResolution>%d %d</ImageResolution>%s<ImageName>rnd%d.png</ImageName></Camera></Cameras>\n"
                % (f3(cam_pos), f3((u(-1, 1), u(0.5, 1.5), u(-1, 1))), f3((u(-0.2, 0.2), 1.0, u(-0.1, 0.1))), u(35, 60), u(1.0, 2.0), width, height, tm, seed))
    else:
        xml += ("<Cameras><Camera id=\"1\"><Position>%s</Position><Gaze>%s</Gaze><Up>0 1 0</Up><NearPlane>-1 1 -0.714 0.714</NearPlane>"
                "<NearDistance>%.4g</NearDistance><ImageResolution>%d %d</ImageResolution>%s<ImageName>rnd%d.png</ImageName></Camera></Cameras>\n"
                % (f3(cam_pos), f3(gaze), u(1.6, 2.4), width, height, tm, seed))
    xml += "<Lights><AmbientLight>%d %d %d</AmbientLight>\n" % (rng.randint(5, 40), rng.randint(5, 40), rng.randint(5, 40))
    for k in range(rng.randint(1, 3)):
        xml += ("<PointLight id=\"%d\"><Position>%s</Position><Intensity>%s</Intensity></PointLight>\n"
                % (k + 1, f3((u(-7, 7), u(4, 10), u(-2, 9))), f3((u(5e3, 3e4), u(5e3, 3e4), u(5e3, 3e4)))))
    if rng.rand() < 0.5:
        xml += ("<DirectionalLight id=\"7\"><Direction>%s</Direction><Radiance>%s</Radiance></DirectionalLight>\n"
                % (f3((u(-0.6, 0.6), -1.0, u(-0.6, 0.3))), f3((u(20, 120), u(20, 120), u(20, 120)))))
    if rng.rand() < 0.5:
        sp = (u(-4, 4), u(5, 9), u(0, 6))
        xml += ("<SpotLight id=\"8\"><Position>%s</Position><Direction>%s</Direction><Intensity>%s</Intensity>"
                "<CoverageAngle>%.4g</CoverageAngle><FalloffAngle>%.4g</FalloffAngle></SpotLight>\n"
                % (f3(sp), f3((-sp[0] * 0.15 + u(-0.1, 0.1), -1.0, -sp[2] * 0.15)), f3((u(2e4, 9e4), u(2e4, 9e4), u(2e4, 9e4))), u(35, 70), u(10, 30)))
    if mc and rng.rand() < 0.6:
        xml += ("<AreaLight id=\"5\"><Position>%s</Position><Normal>%s</Normal><Radiance>%s</Radiance><Size>%.4g</Size></AreaLight>\n"
                % (f3((u(-3, 3), u(5, 8), u(-2, 3))), f3((u(-0.3, 0.3), -1.0, u(-0.3, 0.3))), f3((u(8, 30), u(8, 30), u(8, 30))), u(1, 4)))
    if env_light:
        xml += "<SphericalDirectionalLight id=\"6\"><ImageId>3</ImageId></SphericalDirectionalLight>\n"
    xml += "</Lights>\n"
    xml += ("<BRDFs><OriginalPhong id=\"1\"><Exponent>%.4g</Exponent></OriginalPhong><OriginalBlinnPhong id=\"2\"><Exponent>%.4g</Exponent></OriginalBlinnPhong>"
            "<ModifiedPhong id=\"3\"><Exponent>%.4g</Exponent></ModifiedPhong><ModifiedPhong id=\"4\" normalized=\"true\"><Exponent>%.4g</Exponent></ModifiedPhong>"
            "<ModifiedBlinnPhong id=\"5\"><Exponent>%.4g</Exponent></ModifiedBlinnPhong><ModifiedBlinnPhong id=\"6\" normalized=\"true\"><Exponent>%.4g</Exponent></ModifiedBlinnPhong>"
            "<TorranceSparrow id=\"7\"><Exponent>%.4g</Exponent></TorranceSparrow><TorranceSparrow id=\"8\" kdfresnel=\"true\"><Exponent>%.4g</Exponent></TorranceSparrow></BRDFs>\n"
            % tuple(u(5, 60) for _ in range(8)))
    n_mat = 7
    xml += "<Materials>\n"
    for m in range(1, n_mat + 1):
        kind = "plain" if m <= 2 else ["plain", "plain", "mirror", "conductor", "dielectric"][rng.randint(5)]
        kd = f3((u(0.1, 0.8), u(0.1, 0.8), u(0.1, 0.8)))
        if kind == "plain":
            brdf = rng.randint(0, 9)
            dg = " degamma=\"true\"" if extras and rng.rand() < 0.3 else ""
            xml += ("<Material id=\"%d\"%s%s><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>%s</DiffuseReflectance>"
                    "<SpecularReflectance>%s</SpecularReflectance><PhongExponent>%.4g</PhongExponent><RefractionIndex>%.4g</RefractionIndex></Material>\n"
                    % (m, " BRDF=\"%d\"" % brdf if brdf else "", dg, kd, f3((u(0, 0.6),) * 3), u(3, 60), u(1.2, 2.2)))
        elif kind == "mirror":
            xml += ("<Material id=\"%d\" type=\"mirror\"><AmbientReflectance>0 0 0</AmbientReflectance><DiffuseReflectance>%s</DiffuseReflectance>"
                    "<SpecularReflectance>0.2 0.2 0.2</SpecularReflectance><PhongExponent>20</PhongExponent><MirrorReflectance>%s</MirrorReflectance>%s</Material>\n"
                    % (m, f3((u(0, 0.2),) * 3), f3((u(0.4, 0.95), u(0.4, 0.95), u(0.4, 0.95))), "<Roughness>%.4g</Roughness>" % u(0.02, 0.3) if mc and rng.rand() < 0.5 else ""))
        elif kind == "conductor":
            xml += ("<Material id=\"%d\" type=\"conductor\"><AmbientReflectance>0 0 0</AmbientReflectance><DiffuseReflectance>%s</DiffuseReflectance>"
                    "<SpecularReflectance>0 0 0</SpecularReflectance><MirrorReflectance>%s</MirrorReflectance><RefractionIndex>%.4g</RefractionIndex>"
                    "<AbsorptionIndex>%.4g</AbsorptionIndex></Material>\n"
                    % (m, f3((u(0, 0.2),) * 3), f3((u(0.5, 0.95), u(0.5, 0.95), u(0.4, 0.9))), u(0.2, 1.5), u(2.0, 4.0)))
        else:
            xml += ("<Material id=\"%d\" type=\"dielectric\"><AmbientReflectance>0 0 0</AmbientReflectance><DiffuseReflectance>0 0 0</DiffuseReflectance>"
                    "<SpecularReflectance>0 0 0</SpecularReflectance><AbsorptionCoefficient>%s</AbsorptionCoefficient><RefractionIndex>%.4g</RefractionIndex></Material>\n"
                    % (m, f3((u(0, 0.08), u(0, 0.08), u(0, 0.08))), u(1.2, 2.0)))
    if mc:
        xml += ("<Material id=\"%d\"><AmbientReflectance>0 0 0</AmbientReflectance><DiffuseReflectance>0 0 0</DiffuseReflectance>"
                "<SpecularReflectance>0 0 0</SpecularReflectance></Material>\n" % (n_mat + 1))                # becomes emissive (LightMesh)
    xml += "</Materials>\n"
    verts = [(-11, 0, -9), (11, 0, -9), (11, 0, 9), (-11, 0, 9)]
    objs = "<Mesh id=\"1\"><Material>%d</Material>%s<Faces>1 3 2\n1 4 3</Faces></Mesh>\n" % (rng.randint(1, 3), "PLACEHOLDER_FLOOR_TEX")
    xml_tr = ("<Transformations>" + "".join("<Translation id=\"%d\">%s</Translation>" % (k + 1, f3((u(-2.5, 2.5), u(0, 1.5), u(-2, 2)))) for k in range(3))
              + "".join("<Scaling id=\"%d\">%s</Scaling>" % (k + 1, f3((u(0.6, 1.5), u(0.6, 1.5), u(0.6, 1.5)))) for k in range(3))
              + "".join("<Rotation id=\"%d\">%.4g %s</Rotation>" % (k + 1, u(-80, 80), ["1 0 0", "0 1 0", "0 0 1"][rng.randint(3)]) for k in range(3)) + "</Transformations>\n")   # axis-aligned axes only (parser.cpp:672-683)

    tex_xml = ""
    if textures:
        from dtb200 import scenegen
        yy, xx = np.mgrid[0:48, 0:64]
        a = np.stack([127 + 100 * np.sin(xx / (3.0 + seed % 5)), 127 + 100 * np.cos(yy / 6.0), 90 + 2 * xx], -1).clip(0, 255).astype(np.uint8)
        scenegen.write_png(os.path.join(out_dir, "inputs", "rnd_a.png"), a)
        nm = np.stack([127 + 45 * np.sin(xx / 3.0), 127 + 45 * np.cos(yy / 4.0), 225 + 0 * xx], -1).clip(0, 255).astype(np.uint8)
        scenegen.write_png(os.path.join(out_dir, "inputs", "rnd_n.png"), nm)
        tex_xml = ("<Textures><Images><Image id=\"1\">rnd_a.png</Image><Image id=\"2\">rnd_n.png</Image></Images>\n"
                   "<TextureMap id=\"1\" type=\"image\"><ImageId>1</ImageId><DecalMode>replace_kd</DecalMode><Interpolation>bilinear</Interpolation></TextureMap>\n"
                   "<TextureMap id=\"2\" type=\"image\"><ImageId>1</ImageId><DecalMode>blend_kd</DecalMode><Interpolation>nearest</Interpolation></TextureMap>\n"
                   "<TextureMap id=\"3\" type=\"image\"><ImageId>2</ImageId><DecalMode>replace_normal</DecalMode><Interpolation>bilinear</Interpolation></TextureMap>\n"
                   "<TextureMap id=\"4\" type=\"image\"><ImageId>1</ImageId><DecalMode>bump_normal</DecalMode><BumpFactor>%.4g</BumpFactor></TextureMap>\n"
                   "<TextureMap id=\"5\" type=\"perlin\"><DecalMode>replace_kd</DecalMode><NoiseConversion>absval</NoiseConversion><NoiseScale>%.4g</NoiseScale></TextureMap>\n"
                   "<TextureMap id=\"6\" type=\"perlin\"><DecalMode>bump_normal</DecalMode><NoiseConversion>linear</NoiseConversion><NoiseScale>%.4g</NoiseScale><BumpFactor>%.4g</BumpFactor></TextureMap>\n"
                   "<TextureMap id=\"7\" type=\"image\"><ImageId>1</ImageId><DecalMode>replace_all</DecalMode><Interpolation>bilinear</Interpolation></TextureMap>\n"
                   % (u(0.5, 3), u(0.5, 4), u(1, 4), u(0.2, 1.0)))
        if extras:
            if rng.rand() < 0.5:
                tex_xml += "<TextureMap id=\"8\" type=\"image\"><ImageId>1</ImageId><DecalMode>replace_background</DecalMode><Interpolation>bilinear</Interpolation></TextureMap>\n"
            tex_xml += "<TextureMap id=\"9\" type=\"image\"><ImageId>2</ImageId><DecalMode>replace_ks</DecalMode><Interpolation>nearest</Interpolation></TextureMap>\n"
        tex_xml += "</Textures>\n"

    if env_light:
        from dtb200 import scenegen
        yy, xx = np.mgrid[0:16, 0:32].astype(np.float32)
        sky = 6.0 + 10.0 * (1 - yy / 16) + 3.0 * np.sin(xx / 32 * 2 * np.pi + seed)
        scenegen.write_exr(os.path.join(out_dir, "inputs", "rnd_env.exr"), np.stack([0.8 * sky, 0.9 * sky, 1.0 * sky], axis=-1).astype(np.float32))
        if tex_xml:
            tex_xml = tex_xml.replace("</Images>", "<Image id=\"3\">rnd_env.exr</Image></Images>")
        else:
            tex_xml = "<Textures><Images><Image id=\"1\">rnd_env.exr</Image><Image id=\"2\">rnd_env.exr</Image><Image id=\"3\">rnd_env.exr</Image></Images></Textures>\n"

    def shape_textures(sphere=False):
        if not textures or rng.rand() < 0.5:
            return ""
        ids = []
        if rng.rand() < 0.75:
            ids.append([1, 2, 5, 7][rng.randint(4)])
            if extras and ids[0] in (1, 2, 5) and rng.rand() < 0.3:
                ids.insert(0, 9)                                      # replace_ks reads the DIFFUSE slot (raytracer.cpp:516-531)
        if rng.rand() < 0.5 or not ids:
            ids.append([3, 4, 6][rng.randint(3)])
        if mc and sphere:
            # the image bump map of a sphere reads the texel row below the last one near the south pole (sphere.cpp:127-140 ->
            # LDRImage.h:16-26, no bounds check): heap garbage in the reference, clamped here.  Camera rays of these scenes do not
            # see that row, but GI rays do.
            ids = [6 if k == 4 else k for k in ids]
        return "<Textures>%s</Textures>" % " ".join(str(k) for k in ids)

    def transforms():
        if rng.rand() < 0.45:
            return ""
        toks = [["t", "s", "r"][rng.randint(3)] + str(rng.randint(1, 4)) for _ in range(rng.randint(1, 4))]
        return "<Transformations>%s</Transformations>" % " ".join(toks)

    mesh_id = 2
    first_mesh = None
    for _ in range(rng.randint(1, 4)):
        c = np.array([u(-5, 5), u(0.8, 2.4), u(-3.5, 3.5)])
        base = len(verts)
        kind = rng.randint(4)
        faces = []
        if kind == 0:                                             # box
            h = np.array([u(0.4, 1.1), u(0.4, 1.1), u(0.4, 1.1)])
            for dz in (-1, 1):
                for dy in (-1, 1):
                    for dx in (-1, 1):
                        verts.append(tuple(c + h * (dx, dy, dz)))
            q = [(0, 2, 3, 1), (4, 5, 7, 6), (0, 1, 5, 4), (2, 6, 7, 3), (0, 4, 6, 2), (1, 3, 7, 5)]
            for a, b, cc, d in q:
                faces += [(a, b, cc), (a, cc, d)]
        elif kind == 1:                                           # tetrahedron
            p = [c + (u(-1, 1), u(-0.7, 1.2), u(-1, 1)) for _ in range(4)]
            verts += [tuple(x) for x in p]
            faces = [(0, 1, 2), (0, 3, 1), (1, 3, 2), (0, 2, 3)]
        elif kind == 2:                                           # triangle soup
            n = rng.randint(3, 9)
            for k in range(n):
                o = c + (u(-1.2, 1.2), u(-0.6, 1.0), u(-1.2, 1.2))
                verts += [tuple(o), tuple(o + (u(0.3, 1.2), u(-0.4, 0.4), u(-0.5, 0.5))), tuple(o + (u(-0.4, 0.4), u(0.3, 1.2), u(-0.5, 0.5)))]
                faces.append((3 * k, 3 * k + 1, 3 * k + 2))
        else:                                                     # a single <Triangle>
            verts += [tuple(c + (-1, -0.5, 0)), tuple(c + (1, -0.5, u(-0.5, 0.5))), tuple(c + (u(-0.3, 0.3), 1.0, 0))]
            objs += ("<Triangle id=\"%d\"><Material>%d</Material>%s%s<Indices>%d %d %d</Indices></Triangle>\n"
                     % (mesh_id, rng.randint(1, n_mat + 1), shape_textures(), transforms(), base + 1, base + 2, base + 3))
            mesh_id += 1
            continue
        objs += ("<Mesh id=\"%d\"><Material>%d</Material>%s%s<Faces>%s</Faces></Mesh>\n"
                 % (mesh_id, rng.randint(1, n_mat + 1), shape_textures(), transforms(), "\n".join("%d %d %d" % (base + a + 1, base + b + 1, base + cc + 1) for a, b, cc in faces)))
        if first_mesh is None:
            first_mesh = mesh_id
        mesh_id += 1
    if first_mesh is not None and rng.rand() < 0.7:
        for reset in (["true", "false"] if rng.rand() < 0.5 else ["true"]):
            objs += ("<MeshInstance id=\"%d\" baseMeshId=\"%d\" resetTransform=\"%s\"><Material>%d</Material><Transformations>s%d r%d t%d</Transformations></MeshInstance>\n"
                     % (mesh_id, first_mesh, reset, rng.randint(1, n_mat + 1), rng.randint(1, 4), rng.randint(1, 4), rng.randint(1, 4)))
            mesh_id += 1
    if extras and rng.rand() < 0.6:
        from dtb200 import scenegen
        bv, bf = scenegen.blob_mesh(int(rng.randint(8, 20)), int(rng.randint(5, 11)), u(0.6, 1.0), (u(-3, 3), u(1.0, 2.0), u(-2, 2)))
        scenegen.write_ply(os.path.join(out_dir, "rnd_blob%d.ply" % seed), bv, bf)
        objs += ("<Mesh id=\"%d\"><Material>%d</Material>%s<Faces plyFile=\"rnd_blob%d.ply\" /></Mesh>\n"
                 % (mesh_id, rng.randint(1, n_mat + 1), transforms(), seed))
        objs += ("<MeshInstance id=\"%d\" baseMeshId=\"%d\" resetTransform=\"%s\"><Material>%d</Material><Transformations>t%d r%d</Transformations></MeshInstance>\n"
                 % (mesh_id + 1, mesh_id, ["true", "false"][rng.randint(2)], rng.randint(1, n_mat + 1), rng.randint(1, 4), rng.randint(1, 4)))
        mesh_id += 2
    if mc and rng.rand() < 0.5:
        c = np.array([u(-3, 3), u(4.5, 7), u(-2, 2)])
        base = len(verts)
        h = u(0.6, 1.5)
        verts += [tuple(c + (-h, 0, -h)), tuple(c + (h, 0, -h)), tuple(c + (h, u(-0.3, 0.3), h)), tuple(c + (-h, 0, h))]
        objs += ("<LightMesh id=\"%d\"><Material>%d</Material><Faces>%d %d %d\n%d %d %d</Faces><Radiance>%s</Radiance></LightMesh>\n"
                 % (mesh_id, n_mat + 1, base + 1, base + 2, base + 3, base + 3, base + 4, base + 1, f3((u(5, 20), u(5, 20), u(5, 20)))))
        mesh_id += 1
    for k in range(rng.randint(2, 6)):
        r = u(0.5, 1.2)
        verts.append((u(-5.5, 5.5), r + u(0.0, 1.5), u(-3.5, 4.0)))
        blur = "<MotionBlur>%s</MotionBlur>" % f3((u(-0.6, 0.6), u(0, 0.5), u(-0.4, 0.4))) if mc and rng.rand() < 0.25 else ""
        objs += ("<Sphere id=\"%d\"><Material>%d</Material>%s%s<Center>%d</Center><Radius>%.4g</Radius>%s</Sphere>\n"
                 % (k + 1, rng.randint(1, n_mat + 1), shape_textures(sphere=True), transforms() if rng.rand() < 0.4 else "", len(verts), r, blur))
    objs = objs.replace("PLACEHOLDER_FLOOR_TEX", shape_textures())
    uv_xml = ""
    if textures:
        uv_xml = "<TexCoordData>" + "\n".join("%.5g %.5g" % (u(0, 2), u(0, 2)) for _ in verts) + "</TexCoordData>\n"
    xml += tex_xml + "<VertexData>" + "\n".join(f3(v) for v in verts) + "</VertexData>\n" + uv_xml + xml_tr + "<Objects>\n" + objs + "</Objects></Scene>\n"
    path = os.path.join(out_dir, "rnd%d.xml" % seed)
    with open(path, "w") as f:
        f.write(xml)
    return path
