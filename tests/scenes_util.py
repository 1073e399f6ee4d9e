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
