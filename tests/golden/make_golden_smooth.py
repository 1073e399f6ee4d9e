#!/usr/bin/env python3
"""Fixtures for DT_FLAG_SMOOTH_SHADING: two shadingMode="smooth" scenes of archive/hw1_inputs/akif_uslu.

Run in the build container only (needs /root/reference and oracle/_ref):  python tests/golden/make_golden_smooth.py
Per scene one .npz:
    xml      the scene file bytes
    golden   the course-provided expected render archive/hw1_outputs/akif_uslu/<image>.png (SMOOTH shaded, by the course's renderer)
    ref_ldr  what the compiled reference writes for the same file (it ignores shadingMode: FLAT shaded)
    rays     [closest, shadow] ray counts of the reference run
"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_util import run_reference  # noqa: E402

REF = "/root/reference/archive"
SCENES = {"berserker_smooth": "berserker_smooth.png", "low_poly_smooth": "low_poly_scene_smooth.png"}      # tower_smooth: the reference dies of bad_alloc (every Mesh copies the whole VertexData, mesh.cpp:7-13)


def main():
    for name, png in SCENES.items():
        xml_path = os.path.join(REF, "hw1_inputs", "akif_uslu", name + ".xml")
        xml = np.frombuffer(open(xml_path, "rb").read(), dtype=np.uint8)
        ref = run_reference(xml_path, probe=True)
        out = {"xml": xml, "ref_ldr": ref["png"], "rays": np.array([ref["closest"], ref["shadow"]], dtype=np.int64),
               "golden": np.array(Image.open(os.path.join(REF, "hw1_outputs", "akif_uslu", png)).convert("RGB"))}
        path = os.path.join(HERE, "smooth_" + name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
