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
