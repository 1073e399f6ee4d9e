#!/usr/bin/env python3
"""Generate the committed golden fixtures for the six "pins" scenes.

Run in the build container only (needs /root/reference and oracle/_ref built by oracle/build_ref.py):
    python tests/golden/make_golden.py

For every pins scene of archive/hw1_inputs it stores ONE compressed .npz holding
    xml        the scene file bytes (so the GPU box, which has no /root/reference, can load it)
    golden     the course-provided expected render archive/hw1_outputs/<scene>.png as uint8 HxWx3
    ref_ldr    the LDR image the compiled, unmodified-algorithm reference (oracle/_ref/raytracer_probe) writes
    hit_shape  primary-ray hit shape index per pixel (int16), -1 = miss          } dumped by the probe build
    hit_face   primary-ray canonical face index per pixel (int32), -1 = sphere   } (oracle/build_ref.py I2)
    hit_t      primary-ray hit distance per pixel (float32 bits)
    rays       [closest, shadow] ray counts of the reference run
Nothing else in the repository reads /root/reference at test time.
"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_util import run_reference  # noqa: E402

REF = "/root/reference/archive"
PINS = ["simple", "two_spheres", "spheres", "spheres_mirror", "cornellbox_recursive_conductors", "scienceTree"]
# dielectric scenes: the course PNGs are not reproducible by the reference itself, so only
# the compiled reference's own output is stored for them.
EXTRA = ["scienceTree_diamond", "cornellbox_recursive_alt2"]


def main():
    for name in PINS + EXTRA:
        xml_path = os.path.join(REF, "hw1_inputs", name + ".xml")
        with open(xml_path, "rb") as f:
            xml = np.frombuffer(f.read(), dtype=np.uint8)
        ref = run_reference(xml_path, probe=True)
        out = {
            "xml": xml,
            "ref_ldr": ref["png"],
            "hit_shape": ref["hit_shape"].astype(np.int16),
            "hit_face": ref["hit_face"].astype(np.int32),
            "hit_t": ref["hit_t"].astype(np.float32),
            "rays": np.array([ref["closest"], ref["shadow"]], dtype=np.int64),
        }
        if name in PINS:
            out["golden"] = np.array(Image.open(os.path.join(REF, "hw1_outputs", name + ".png")).convert("RGB"))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
