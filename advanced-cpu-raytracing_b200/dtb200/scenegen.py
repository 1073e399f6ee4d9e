"""Seeded procedural scene generators for BASELINE.json configs 2-5: the named bunny / dragon
assets are stripped from the reference mount (.MISSING_LARGE_BLOBS), so each config is generated as
binary-LE PLY + XML in the reference's own schema — loadable by the reference binary
and by the host mirror alike.  All sizes are parameters so tests can use small instances of the same shape.
"""
import os
import struct

import numpy as np


# ------------------------------------------------------------------ file writers
def write_ply(path, verts, faces):
    verts = np.ascontiguousarray(verts, dtype="<f4")
    faces = np.ascontiguousarray(faces, dtype="<i4")
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                 "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(verts), len(faces))).encode())
        f.write(verts.tobytes())
        rec = np.empty(len(faces), dtype=[("n", "u1"), ("i", "<i4", (3,))])
        rec["n"] = 3
        rec["i"] = faces
        f.write(rec.tobytes())


def write_exr(path, rgb):
    """Uncompressed scanline OpenEXR, float32 channels B,G,R (readable by tinyexr's LoadEXR and by dth_io.cpp)."""
    rgb = np.ascontiguousarray(rgb, dtype="<f4")
    h, w = rgb.shape[:2]

    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data
    chl = b""
    for c in ("B", "G", "R"):
        chl += c.encode() + b"\0" + struct.pack("<iBBBBii", 2, 0, 0, 0, 0, 1, 1)
    chl += b"\0"
    hdr = struct.pack("<II", 20000630, 2)
    hdr += attr("channels", "chlist", chl)
    hdr += attr("compression", "compression", b"\0")
    hdr += attr("dataWindow", "box2i", struct.pack("<iiii", 0, 0, w - 1, h - 1))
    hdr += attr("displayWindow", "box2i", struct.pack("<iiii", 0, 0, w - 1, h - 1))
    hdr += attr("lineOrder", "lineOrder", b"\0")
    hdr += attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    hdr += attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
    hdr += attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    hdr += b"\0"
    line_bytes = w * 4 * 3
    off0 = len(hdr) + 8 * h
    with open(path, "wb") as f:
        f.write(hdr)
        for y in range(h):
            f.write(struct.pack("<Q", off0 + y * (8 + line_bytes)))
        for y in range(h):
            f.write(struct.pack("<ii", y, line_bytes))
            f.write(rgb[y, :, 2].tobytes()); f.write(rgb[y, :, 1].tobytes()); f.write(rgb[y, :, 0].tobytes())


def write_png(path, rgb):
    from PIL import Image
    Image.fromarray(np.ascontiguousarray(rgb, dtype=np.uint8)).save(path)


# ------------------------------------------------------------------ geometry
def blob_mesh(nlon, nlat, radius=60.0, center=(0.0, 0.0, 0.0)):
    """Displaced lat-long sphere of SURVEY.md 8d config 2: r = R(1 + 0.15 sin8u cos6v + 0.05 sin(31u+17v)).
    2*nlon*(nlat-1) triangles (nlon=1000, nlat=499 -> 996 000)."""
    u = (np.arange(nlon, dtype=np.float64) / nlon) * 2 * np.pi
    v = ((np.arange(nlat, dtype=np.float64) + 0.5) / nlat) * np.pi
    U, Vv = np.meshgrid(u, v)                       # [nlat, nlon]
    r = radius * (1 + 0.15 * np.sin(8 * U) * np.cos(6 * Vv) + 0.05 * np.sin(31 * U + 17 * Vv))
    x = r * np.sin(Vv) * np.cos(U) + center[0]
    y = r * np.cos(Vv) + center[1]
    z = r * np.sin(Vv) * np.sin(U) + center[2]
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    i = np.arange(nlat - 1)[:, None]
    j = np.arange(nlon)[None, :]
    a = i * nlon + j
    b = i * nlon + (j + 1) % nlon
    c = (i + 1) * nlon + j
    d = (i + 1) * nlon + (j + 1) % nlon
    # outward-facing winding
    f1 = np.stack([a, b, c], axis=-1).reshape(-1, 3)
    f2 = np.stack([b, d, c], axis=-1).reshape(-1, 3)
    faces = np.concatenate([f1, f2], axis=0).astype(np.int32)
    return verts, faces


def _xml_camera(cam_id, pos, gaze_point, up, fovy, near, w, h, name, spp=1, extra=""):
    s = ('    <Camera id="%d" type="lookAt">\n        <Position>%g %g %g</Position>\n        <GazePoint>%g %g %g</GazePoint>\n'
         '        <Up>%g %g %g</Up>\n        <FovY>%g</FovY>\n        <NearDistance>%g</NearDistance>\n'
         '        <ImageResolution>%d %d</ImageResolution>\n' % ((cam_id,) + tuple(pos) + tuple(gaze_point) + tuple(up) + (fovy, near, w, h)))
    if spp > 1:
        s += "        <NumSamples>%d</NumSamples>\n" % spp
    s += extra
    s += "        <ImageName>%s</ImageName>\n    </Camera>\n" % name
    return s


# ------------------------------------------------------------------ config 2
def gen_config2(out_dir, nlon=1000, nlat=499, width=1920, height=1080, depth=6, name="config2"):
    """Procedural blob (mirror) + ground quad + one dielectric sphere, 2 point lights, recursion depth 6."""
    os.makedirs(out_dir, exist_ok=True)
    verts, faces = blob_mesh(nlon, nlat, 60.0, (0.0, 75.0, 0.0))
    write_ply(os.path.join(out_dir, "blob.ply"), verts, faces)
    xml = "<Scene>\n    <MaxRecursionDepth>%d</MaxRecursionDepth>\n    <BackgroundColor>20 30 60</BackgroundColor>\n    <ShadowRayEpsilon>1e-2</ShadowRayEpsilon>\n" % depth
    xml += "    <Cameras>\n" + _xml_camera(1, (0, 110, 260), (0, 70, 0), (0, 1, 0), 40, 1, width, height, name + ".png") + "    </Cameras>\n"
    xml += ("    <Lights>\n        <AmbientLight>25 25 25</AmbientLight>\n"
            "        <PointLight id=\"1\"><Position>200 300 250</Position><Intensity>9000000 9000000 9000000</Intensity></PointLight>\n"
            "        <PointLight id=\"2\"><Position>-250 200 100</Position><Intensity>4000000 3500000 3000000</Intensity></PointLight>\n    </Lights>\n")
    xml += ("    <Materials>\n"
            "        <Material id=\"1\" type=\"mirror\">\n            <AmbientReflectance>0.2 0.2 0.25</AmbientReflectance>\n            <DiffuseReflectance>0.25 0.3 0.45</DiffuseReflectance>\n"
            "            <SpecularReflectance>0.6 0.6 0.6</SpecularReflectance>\n            <PhongExponent>40</PhongExponent>\n            <MirrorReflectance>0.55 0.55 0.55</MirrorReflectance>\n        </Material>\n"
            "        <Material id=\"2\">\n            <AmbientReflectance>0.3 0.3 0.3</AmbientReflectance>\n            <DiffuseReflectance>0.55 0.5 0.4</DiffuseReflectance>\n"
            "            <SpecularReflectance>0.1 0.1 0.1</SpecularReflectance>\n            <PhongExponent>5</PhongExponent>\n        </Material>\n"
            "        <Material id=\"3\" type=\"dielectric\">\n            <AmbientReflectance>0 0 0</AmbientReflectance>\n            <DiffuseReflectance>0 0 0</DiffuseReflectance>\n"
            "            <SpecularReflectance>0 0 0</SpecularReflectance>\n            <AbsorptionCoefficient>0.01 0.002 0.002</AbsorptionCoefficient>\n            <RefractionIndex>1.5</RefractionIndex>\n        </Material>\n"
            "    </Materials>\n")
    xml += ("    <VertexData>\n        -600 0 -600\n        600 0 -600\n        600 0 600\n        -600 0 600\n        95 35 95\n    </VertexData>\n")
    xml += ("    <Objects>\n        <Mesh id=\"1\">\n            <Material>1</Material>\n            <Faces plyFile=\"blob.ply\" />\n        </Mesh>\n"
            "        <Mesh id=\"2\">\n            <Material>2</Material>\n            <Faces>\n                1 3 2\n                1 4 3\n            </Faces>\n        </Mesh>\n"
            "        <Sphere id=\"1\">\n            <Material>3</Material>\n            <Center>5</Center>\n            <Radius>35</Radius>\n        </Sphere>\n    </Objects>\n</Scene>\n")
    path = os.path.join(out_dir, name + ".xml")
    with open(path, "w") as f:
        f.write(xml)
    return path


# ------------------------------------------------------------------ config 3
def gen_config3(out_dir, grid=64, base_nlon=128, base_nlat=65, width=1920, height=1080, spp=16, seed=1234, name="config3"):
    """grid*grid MeshInstances of one ~16k-triangle base mesh (transform ids are single-digit in the reference
    parser, so positions are composed from repeated translations t1..t4), an image-textured ground with
    TexCoordData, a Perlin texture on the instances, point lights only (deterministic)."""
    os.makedirs(os.path.join(out_dir, "inputs"), exist_ok=True)
    rng = np.random.RandomState(seed)
    verts, faces = blob_mesh(base_nlon, base_nlat, 1.0, (0.0, 0.0, 0.0))
    write_ply(os.path.join(out_dir, "base.ply"), verts, faces)
    # ground texture: soft checker with gradients
    yy, xx = np.mgrid[0:256, 0:256]
    tex = np.zeros((256, 256, 3), np.uint8)
    chk = ((xx // 32 + yy // 32) % 2).astype(np.float32)
    tex[..., 0] = (90 + 120 * chk + 0.1 * xx).clip(0, 255)
    tex[..., 1] = (80 + 100 * (1 - chk) + 0.2 * yy).clip(0, 255)
    tex[..., 2] = (70 + 60 * chk).clip(0, 255)
    write_png(os.path.join(out_dir, "inputs", "ground.png"), tex)
    span = 3.2
    half = grid * span / 2
    xml = "<Scene>\n    <MaxRecursionDepth>0</MaxRecursionDepth>\n    <BackgroundColor>10 12 20</BackgroundColor>\n    <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n"
    xml += "    <Cameras>\n" + _xml_camera(1, (0, half * 0.55, half * 1.25), (0, 0, 0), (0, 1, 0), 45, 1, width, height, name + ".png", spp) + "    </Cameras>\n"
    xml += ("    <Lights>\n        <AmbientLight>30 30 30</AmbientLight>\n"
            "        <PointLight id=\"1\"><Position>%g %g %g</Position><Intensity>%g %g %g</Intensity></PointLight>\n"
            "        <PointLight id=\"2\"><Position>%g %g %g</Position><Intensity>%g %g %g</Intensity></PointLight>\n    </Lights>\n"
            % (half, half * 1.5, half, 900 * half * half, 900 * half * half, 850 * half * half,
               -half * 0.8, half, -half * 0.3, 400 * half * half, 420 * half * half, 500 * half * half))
    xml += ("    <Materials>\n"
            "        <Material id=\"1\">\n            <AmbientReflectance>0.3 0.3 0.3</AmbientReflectance>\n            <DiffuseReflectance>0.6 0.6 0.6</DiffuseReflectance>\n"
            "            <SpecularReflectance>0.3 0.3 0.3</SpecularReflectance>\n            <PhongExponent>20</PhongExponent>\n        </Material>\n"
            "        <Material id=\"2\">\n            <AmbientReflectance>0.3 0.3 0.3</AmbientReflectance>\n            <DiffuseReflectance>0.5 0.5 0.5</DiffuseReflectance>\n"
            "            <SpecularReflectance>0.05 0.05 0.05</SpecularReflectance>\n            <PhongExponent>3</PhongExponent>\n        </Material>\n"
            "        <Material id=\"3\">\n            <AmbientReflectance>0.3 0.2 0.2</AmbientReflectance>\n            <DiffuseReflectance>0.7 0.3 0.2</DiffuseReflectance>\n"
            "            <SpecularReflectance>0.4 0.4 0.4</SpecularReflectance>\n            <PhongExponent>50</PhongExponent>\n        </Material>\n"
            "    </Materials>\n")
    xml += ("    <Textures>\n        <Images>\n            <Image id=\"1\">ground.png</Image>\n        </Images>\n"
            "        <TextureMap id=\"1\" type=\"image\">\n            <ImageId>1</ImageId>\n            <DecalMode>replace_kd</DecalMode>\n            <Interpolation>bilinear</Interpolation>\n        </TextureMap>\n"
            "        <TextureMap id=\"2\" type=\"perlin\">\n            <DecalMode>replace_kd</DecalMode>\n            <NoiseConversion>absval</NoiseConversion>\n            <NoiseScale>2</NoiseScale>\n        </TextureMap>\n"
            "    </Textures>\n")
    g = half * 1.3
    xml += "    <VertexData>\n        %g 0 %g\n        %g 0 %g\n        %g 0 %g\n        %g 0 %g\n    </VertexData>\n" % (-g, -g, g, -g, g, g, -g, g)
    xml += "    <TexCoordData>\n        0 0\n        4 0\n        4 4\n        0 4\n    </TexCoordData>\n"
    # transformations: t1/t2 = x steps (1, 8 cells), t3/t4 = z steps, t5 = origin shift, s6..s8 scalings, r9/r... rotations
    xml += ("    <Transformations>\n"
            "        <Translation id=\"1\">%g 0 0</Translation>\n        <Translation id=\"2\">%g 0 0</Translation>\n"
            "        <Translation id=\"3\">0 0 %g</Translation>\n        <Translation id=\"4\">0 0 %g</Translation>\n"
            "        <Translation id=\"5\">%g 1.2 %g</Translation>\n"
            "        <Scaling id=\"1\">1 1 1</Scaling>\n        <Scaling id=\"2\">0.8 1.3 0.8</Scaling>\n        <Scaling id=\"3\">1.2 0.7 1.2</Scaling>\n"
            "        <Rotation id=\"1\">30 0 1 0</Rotation>\n        <Rotation id=\"2\">75 0 1 0</Rotation>\n        <Rotation id=\"3\">20 1 0 0</Rotation>\n"
            "    </Transformations>\n" % (span, span * 8, span, span * 8, -half + span / 2, -half + span / 2))
    xml += ("    <Objects>\n        <Mesh id=\"1\">\n            <Material>1</Material>\n            <Textures>2</Textures>\n            <Transformations>t5</Transformations>\n            <Faces plyFile=\"base.ply\" />\n        </Mesh>\n"
            "        <Mesh id=\"2\">\n            <Material>2</Material>\n            <Textures>1</Textures>\n            <Faces>\n                1 3 2\n                1 4 3\n            </Faces>\n        </Mesh>\n")
    inst_id = 10
    for iz in range(grid):
        for ix in range(grid):
            if ix == 0 and iz == 0:
                continue              # the base mesh itself sits in cell (0,0)
            toks = ["s%d" % (1 + rng.randint(3)), "r%d" % (1 + rng.randint(3))]
            toks += ["t1"] * (ix % 8) + ["t2"] * (ix // 8) + ["t3"] * (iz % 8) + ["t4"] * (iz // 8) + ["t5"]
            mat = 1 if rng.rand() < 0.7 else 3
            xml += ("        <MeshInstance id=\"%d\" baseMeshId=\"1\" resetTransform=\"true\">\n            <Material>%d</Material>\n            <Textures>2</Textures>\n"
                    "            <Transformations>%s</Transformations>\n        </MeshInstance>\n" % (inst_id, mat, " ".join(toks)))
            inst_id += 1
    xml += "    </Objects>\n</Scene>\n"
    path = os.path.join(out_dir, name + ".xml")
    with open(path, "w") as f:
        f.write(xml)
    return path


# ------------------------------------------------------------------ config 4 / 5
def _env_map(w=256, h=128):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    sky = 0.25 + 0.5 * (1 - yy / h)
    env = np.stack([0.6 * sky, 0.75 * sky, 1.0 * sky], axis=-1)
    hot = np.exp(-(((xx - 0.7 * w) / (0.04 * w)) ** 2 + ((yy - 0.25 * h) / (0.06 * h)) ** 2))
    env += hot[..., None] * np.array([30.0, 26.0, 20.0], np.float32)
    return env.astype(np.float32)


def gen_config4(out_dir, width=1920, height=1080, spp=256, depth=4, blob=(0, 0), name="config4", nee=True, rr=True, importance=True, n_cameras=1, sphere_lift=0.0,
                lights=("area", "mesh", "env"), area_light_y=9.9, mirror_brdf=False):
    """Path-traced box + spheres: area light + light mesh + spherical HDR environment light, Torrance-Sparrow
    (kdfresnel) and modified Blinn-Phong BRDFs, photographic tonemapper.  blob=(nlon,nlat) adds a displaced
    sphere mesh of that tessellation (config 5 uses 3162 x 1581 ~ 10 M triangles).
    lights: which of the three sampled light types the scene holds (one-light variants pin each estimator on its own).
    area_light_y: height of the area light.  At 9.9 it floats 0.1 under the ceiling, whose points see it at d -> 0.1: the
    L A cos / d^2 estimator (raytracer.cpp:722-739) then has no finite variance and images do not converge in PSNR however
    many samples are taken (measured: 17.5 dB at 256 spp, 19.8 dB at 65536 spp); the area-light-only fixture lowers it.
    sphere_lift: the three spheres rest ON the floor when 0 (round 1's scene).  The zero-angle wedge at a contact point traps
    pure-GI paths -- the reference's Russian roulette never ends a pure GI chain (raytracer.cpp:137-147) -- for thousands of
    bounces, a latency-bound tail that is < 0.4 % of a 1024-spp frame but 16 % of a 16-spp one; config 5 lifts them."""
    os.makedirs(os.path.join(out_dir, "inputs"), exist_ok=True)
    write_exr(os.path.join(out_dir, "inputs", "env.exr"), _env_map())
    params = " ".join(p for p, on in (("NextEventEstimation", nee), ("ImportanceSampling", importance), ("RussianRoulette", rr)) if on)
    extra = ("        <Renderer>PathTracing</Renderer>\n        <RendererParams>%s</RendererParams>\n"
             "        <Tonemap>\n            <TMO>Photographic</TMO>\n            <TMOOptions>0.18 1</TMOOptions>\n            <Saturation>1.0</Saturation>\n            <Gamma>2.2</Gamma>\n        </Tonemap>\n" % params)
    xml = "<Scene>\n    <MaxRecursionDepth>%d</MaxRecursionDepth>\n    <BackgroundColor>0 0 0</BackgroundColor>\n    <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>\n" % depth
    # n_cameras identical cameras: the reference renders every camera of a scene in one process (main.cpp:142), which gives
    # the CPU arm of bench.py several steps per scene load
    xml += "    <Cameras>\n" + "".join(_xml_camera(k + 1, (0, 0, 24), (0, -1, 0), (0, 1, 0), 40, 1, width, height, name + ".exr", spp, extra) for k in range(n_cameras)) + "    </Cameras>\n"
    xml += "    <Lights>\n"
    if "area" in lights:
        xml += "        <AreaLight id=\"1\">\n            <Position>0 %g 0</Position>\n            <Normal>0 -1 0</Normal>\n            <Radiance>18 17 15</Radiance>\n            <Size>4</Size>\n        </AreaLight>\n" % area_light_y
    if "env" in lights:
        xml += "        <SphericalDirectionalLight id=\"2\">\n            <ImageId>1</ImageId>\n        </SphericalDirectionalLight>\n"
    xml += "    </Lights>\n"
    xml += ("    <BRDFs>\n        <TorranceSparrow id=\"1\" kdfresnel=\"true\">\n            <Exponent>40</Exponent>\n        </TorranceSparrow>\n"
            "        <ModifiedBlinnPhong id=\"2\" normalized=\"true\">\n            <Exponent>30</Exponent>\n        </ModifiedBlinnPhong>\n    </BRDFs>\n")
    def mat(i, kd, ks=(0, 0, 0), attrs="", more=""):
        return ("        <Material id=\"%d\"%s>\n            <AmbientReflectance>0 0 0</AmbientReflectance>\n            <DiffuseReflectance>%g %g %g</DiffuseReflectance>\n"
                "            <SpecularReflectance>%g %g %g</SpecularReflectance>\n            <PhongExponent>10</PhongExponent>\n%s        </Material>\n" % ((i, attrs) + tuple(kd) + tuple(ks) + (more,)))
    xml += "    <Materials>\n"
    # The reference's GI estimator multiplies every bounce by f*cos*2pi whatever the sampling pdf (raytracer.cpp:187-188)
    # and Russian roulette never shortens pure GI chains (throughput stays 1), so un-normalised Phong materials with
    # kd > ~0.24 make its radiance diverge.  The diffuse surfaces therefore use the normalised (kd/pi) BRDF.
    xml += mat(1, (0.5, 0.5, 0.5), (0, 0, 0), ' BRDF="2"')                # white walls
    xml += mat(2, (0.5, 0.1, 0.08), (0, 0, 0), ' BRDF="2"')               # red
    xml += mat(3, (0.1, 0.45, 0.15), (0, 0, 0), ' BRDF="2"')              # green
    xml += mat(4, (0.4, 0.35, 0.2), (0.5, 0.5, 0.5), ' BRDF="1"', "            <RefractionIndex>1.8</RefractionIndex>\n")    # Torrance-Sparrow
    xml += mat(5, (0.2, 0.3, 0.6), (0.4, 0.4, 0.4), ' BRDF="2"')         # modified Blinn-Phong
    xml += mat(6, (0, 0, 0))                                               # emissive (set by LightMesh)
    # mirror_brdf: the mirror also carries a BRDF, the one case where the GPU path's Russian-roulette throughput differs from the
    # reference's (DESIGN.md "Documented deviations": Shade() scales ray.throughput for UNSHADOWED lights only, raytracer.cpp:192-206)
    if mirror_brdf: xml += mat(7, (0.3, 0.3, 0.3), (0.3, 0.3, 0.3), ' type="mirror" BRDF="2"', "            <MirrorReflectance>0.9 0.9 0.9</MirrorReflectance>\n")
    else: xml += mat(7, (0, 0, 0), (0, 0, 0), ' type="mirror"', "            <MirrorReflectance>0.9 0.9 0.9</MirrorReflectance>\n")
    xml += "    </Materials>\n"
    xml += "    <Textures>\n        <Images>\n            <Image id=\"1\">env.exr</Image>\n        </Images>\n    </Textures>\n"
    v = [(-10, -10, 10), (10, -10, 10), (10, 10, 10), (-10, 10, 10), (-10, -10, -10), (10, -10, -10), (10, 10, -10), (-10, 10, -10),
         (-5, -6.5 + sphere_lift, -2), (5, -7 + sphere_lift, 2), (0, -7.5 + sphere_lift, 5),      # sphere centres 9..11
         (-9.99, 2, -3), (-9.99, 2, 3), (-9.99, 6, 3), (-9.99, 6, -3)]      # light mesh quad 12..15 on the left wall
    xml += "    <VertexData>\n" + "".join("        %g %g %g\n" % p for p in v) + "    </VertexData>\n"
    xml += "    <Objects>\n"
    def mesh(i, m, tris):
        return ("        <Mesh id=\"%d\">\n            <Material>%d</Material>\n            <Faces>\n" % (i, m) +
                "".join("                %d %d %d\n" % t for t in tris) + "            </Faces>\n        </Mesh>\n")
    xml += mesh(1, 1, [(1, 2, 6), (6, 5, 1)])      # floor
    xml += mesh(2, 1, [(5, 6, 7), (7, 8, 5)])      # back
    xml += mesh(3, 2, [(8, 4, 1), (8, 1, 5)])      # left (red)
    xml += mesh(4, 3, [(2, 3, 7), (2, 7, 6)])      # right (green)
    # the ceiling is left open towards the environment light except a strip
    xml += mesh(5, 1, [(3, 4, 8), (8, 7, 3)])      # ceiling
    if blob[0] > 0:
        verts, faces = blob_mesh(blob[0], blob[1], 3.2, (3.5, -3.0, -3.0))
        write_ply(os.path.join(out_dir, "blob.ply"), verts, faces)
        xml += "        <Mesh id=\"6\">\n            <Material>5</Material>\n            <Faces plyFile=\"blob.ply\" />\n        </Mesh>\n"
    if "mesh" in lights:
        xml += ("        <LightMesh id=\"7\">\n            <Material>6</Material>\n            <Faces>\n                12 13 14\n                14 15 12\n            </Faces>\n            <Radiance>6 7 9</Radiance>\n        </LightMesh>\n")
    xml += "        <Sphere id=\"1\">\n            <Material>4</Material>\n            <Center>9</Center>\n            <Radius>3.5</Radius>\n        </Sphere>\n"
    xml += "        <Sphere id=\"2\">\n            <Material>7</Material>\n            <Center>10</Center>\n            <Radius>3</Radius>\n        </Sphere>\n"
    xml += "        <Sphere id=\"3\">\n            <Material>5</Material>\n            <Center>11</Center>\n            <Radius>2.5</Radius>\n        </Sphere>\n"
    xml += "    </Objects>\n</Scene>\n"
    path = os.path.join(out_dir, name + ".xml")
    with open(path, "w") as f:
        f.write(xml)
    return path


def gen_config5(out_dir, nlon=3162, nlat=1581, width=3840, height=2160, spp=1024, name="config5", sphere_lift=0.25, **kw):
    """Config 5: the config-4 scene with a ~10 M-triangle blob (2*nlon*(nlat-1) = 9 991 920 for the defaults); the spheres
    float 0.25 above the floor (see gen_config4; sphere_lift=0 gives round 1's scene)."""
    return gen_config4(out_dir, width=width, height=height, spp=spp, blob=(nlon, nlat), name=name, sphere_lift=sphere_lift, **kw)
