"""PLY input variants (the loader): ascii / binary little- and big-endian, float and double positions, extra vertex
and face properties, every list count / index type, `vertex_index` naming, comments, quads (split 0-1-2 / 2-3-0, parser.cpp:1428-1438)
and a pentagon (skipped, :1440).  The host loader (host/dth_io.cpp) must read what the reference's happly reads: with the compiled
reference present the oracle render of the loaded scene is compared bit for bit with the reference's render of the same files;
everywhere, all variants of one mesh must load to the same geometry."""
import os
import struct

import numpy as np
import pytest

from dtb200 import scenegen
from dtb200.scene import HostScene
from oracle_util import have_ref, oracle_render, run_reference

def write_variant(path, verts, faces, variant, rng):
    """faces: list of index lists (3, 4 or 5 long)."""
    nv = len(verts)
    fmt = {"ascii": "ascii", "ble": "binary_little_endian", "bbe": "binary_big_endian"}[variant["fmt"]]
    e = ">" if variant["fmt"] == "bbe" else "<"
    pos_t = variant.get("pos", "float")
    extra = variant.get("extra", False)
    cnt_t, idx_t = variant.get("list", ("uchar", "int"))
    names = {"float": "f", "double": "d", "float32": "f", "float64": "d", "uchar": "B", "uint8": "B", "int": "i", "int32": "i", "uint": "I", "uint32": "I", "ushort": "H", "short": "h"}
    hdr = "ply\nformat %s 1.0\n" % fmt
    if variant.get("comments"):
        hdr += "comment made by a test\nobj_info something 1 2 3\n"
    hdr += "element vertex %d\nproperty %s x\nproperty %s y\nproperty %s z\n" % (nv, pos_t, pos_t, pos_t)
    if extra:
        hdr += "property float nx\nproperty float ny\nproperty float nz\nproperty float u\nproperty float v\nproperty uchar red\n"
    hdr += "element face %d\nproperty list %s %s %s\n" % (len(faces), cnt_t, idx_t, variant.get("face_name", "vertex_indices"))
    if variant.get("face_extra"):
        hdr += "property uchar flags\n"
    hdr += "end_header\n"
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if variant["fmt"] == "ascii":
            for v in verts:
                row = ["%.9g" % c for c in v]
                if extra: row += ["0", "1", "0", "%.4g" % rng.rand(), "%.4g" % rng.rand(), "200"]
                f.write((" ".join(row) + "\n").encode())
            for fc in faces:
                row = [str(len(fc))] + [str(k) for k in fc]
                if variant.get("face_extra"): row.append("7")
                f.write((" ".join(row) + "\n").encode())
        else:
            for v in verts:
                f.write(struct.pack(e + "3" + names[pos_t], *[float(c) for c in v]))
                if extra: f.write(struct.pack(e + "5fB", 0, 1, 0, rng.rand(), rng.rand(), 200))
            for fc in faces:
                f.write(struct.pack(e + names[cnt_t], len(fc)) + struct.pack(e + "%d%s" % (len(fc), names[idx_t]), *fc))
                if variant.get("face_extra"): f.write(struct.pack("B", 7))

VARIANTS = [dict(fmt="ascii"), dict(fmt="ascii", extra=True, comments=True), dict(fmt="ble", extra=True), dict(fmt="bbe"), dict(fmt="ble", pos="double"),
            dict(fmt="ble", list=("uchar", "uint")), dict(fmt="ble", list=("uint8", "int32"), pos="float32"), dict(fmt="ascii", face_extra=True),
            dict(fmt="ble", face_extra=True, extra=True), dict(fmt="ble", face_name="vertex_index"), dict(fmt="bbe", pos="double", extra=True, list=("ushort", "uint")),
            dict(fmt="ble", list=("uchar", "short"))]

def scene(out_dir, k, variant, quads):
    rng = np.random.RandomState(40)                     # (the same mesh for every variant)
    os.makedirs(out_dir, exist_ok=True)
    verts, tris = scenegen.blob_mesh(12, 7, 1.0, (0.0, 1.2, 0.0))
    faces = [list(map(int, t)) for t in tris]
    if quads:
        # a few quads (two triangles sharing an edge are not guaranteed planar: the reference splits 0-1-2 / 2-3-0 anyway) and one pentagon
        faces = faces[: len(faces) // 2] + [[int(a), int(b), int(c), int(rng.randint(len(verts)))] for a, b, c in tris[len(tris) // 2:]]
        faces.append([0, 1, 2, 3, 4])
    write_variant(os.path.join(out_dir, "m%d.ply" % k), verts, faces, variant, rng)
    xml = ("<Scene><MaxRecursionDepth>2</MaxRecursionDepth><BackgroundColor>20 30 60</BackgroundColor><ShadowRayEpsilon>1e-3</ShadowRayEpsilon>"
           "<Cameras><Camera id=\"1\"><Position>0 1.6 4.5</Position><Gaze>0 -0.1 -1</Gaze><Up>0 1 0</Up><NearPlane>-1 1 -0.7 0.7</NearPlane>"
           "<NearDistance>1.6</NearDistance><ImageResolution>96 64</ImageResolution><ImageName>p%d.png</ImageName></Camera></Cameras>"
           "<Lights><AmbientLight>25 25 25</AmbientLight><PointLight id=\"1\"><Position>3 5 4</Position><Intensity>9000 9000 8000</Intensity></PointLight></Lights>"
           "<Materials><Material id=\"1\"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>0.6 0.5 0.3</DiffuseReflectance><SpecularReflectance>0.3 0.3 0.3</SpecularReflectance><PhongExponent>20</PhongExponent></Material>"
           "<Material id=\"2\" type=\"mirror\"><AmbientReflectance>0 0 0</AmbientReflectance><DiffuseReflectance>0.1 0.1 0.1</DiffuseReflectance><SpecularReflectance>0 0 0</SpecularReflectance><MirrorReflectance>0.7 0.7 0.7</MirrorReflectance></Material></Materials>"
           "<VertexData>-6 0 -6\n6 0 -6\n6 0 6\n-6 0 6</VertexData>"
           "<Objects><Mesh id=\"1\"><Material>2</Material><Faces>1 3 2\n1 4 3</Faces></Mesh>"
           "<Mesh id=\"2\"><Material>1</Material><Faces plyFile=\"m%d.ply\" /></Mesh></Objects></Scene>" % (k, k))
    p = os.path.join(out_dir, "p%d.xml" % k)
    open(p, "w").write(xml)
    return p

needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (GPU box or fresh clone)")


@needs_ref
@pytest.mark.parametrize("quads", [False, True])
@pytest.mark.parametrize("k", range(len(VARIANTS)))
def test_ply_variant_oracle_bit_exact_vs_reference(tmp_path, k, quads):
    p = scene(str(tmp_path / "ply"), k, VARIANTS[k], quads)
    hs = HostScene(p)
    assert hs.n_triangles() == (218 if quads else 146)
    ref = run_reference(p)
    ldr, hdr, st = oracle_render(hs, hs.camera(0))
    assert np.array_equal(ldr, ref["png"])
    assert np.array_equal(hdr.view(np.uint32), ref["hdr"].view(np.uint32))
    assert (int(st.rays_closest), int(st.rays_shadow)) == (ref["closest"], ref["shadow"])


@pytest.mark.parametrize("quads", [False, True])
def test_ply_variants_load_to_the_same_geometry(tmp_path, quads):
    """No reference needed: float-position variants of one mesh give the same frame (double positions are rounded to float by the
    loader like Vec3f{vPos[i][0], ...} does, parser.cpp:1408, so they agree too)."""
    frames = []
    for k, v in enumerate(VARIANTS):
        hs = HostScene(scene(str(tmp_path / "ply"), k, v, quads))
        frames.append(oracle_render(hs, hs.camera(0))[1])
    for k, f in enumerate(frames[1:]):
        assert np.array_equal(f.view(np.uint32), frames[0].view(np.uint32)), VARIANTS[k + 1]
