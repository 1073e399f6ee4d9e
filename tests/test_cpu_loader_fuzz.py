"""Mutation fuzz of the host scene loader (host/dth_scene.cpp, dth_io.cpp): truncated files, dropped tags / elements / bytes, odd
numbers (nan, inf, empty, text), out-of-range material / vertex / texture / image / transformation ids, bad face indices, stray
brackets.  The reference's parser crashes on most of these; the host mirror must either load the scene or return an error message
-- never take the process down (it runs inside the caller's process, behind a C ABI).  The loads run in a child process so that a
crash is a test failure, not the end of the test session."""
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CHILD = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from dtb200.scene import HostScene
for p in sys.argv[1:]:
    print("BEGIN", p, flush=True)
    try:
        print("LOADED", p, HostScene(p).n_triangles(), flush=True)
    except Exception as e:
        print("ERROR", p, str(e)[:120].replace("\n", " "), flush=True)
''' % (os.path.join(os.path.dirname(HERE), "advanced-cpu-raytracing_b200"), HERE)


def _mutate(x, rng):
    kind = rng.randint(8)
    if kind == 0:
        return x[:rng.randint(10, len(x))]
    if kind == 1:
        tags = [m.span() for m in re.finditer(r"<[^>]+>", x)]
        a, b = tags[rng.randint(len(tags))]
        return x[:a] + x[b:]
    if kind == 2:
        i = rng.randint(len(x) - 40)
        return x[:i] + x[i + rng.randint(1, 40):]
    if kind == 3:
        nums = [m.span() for m in re.finditer(r"-?\d+\.?\d*", x)]
        a, b = nums[rng.randint(len(nums))]
        return x[:a] + ["nan", "inf", "-1", "999999999", "", "abc", "1e40", "0"][rng.randint(8)] + x[b:]
    if kind == 4:
        els = list(re.finditer(r"<(\w+)[^>]*>[^<]*</\1>", x))
        m = els[rng.randint(len(els))]
        return x[:m.start()] + x[m.end():]
    if kind == 5:
        ms = list(re.finditer(r"<(Material|Center|Textures|ImageId|Indices)>(\d+)", x))
        m = ms[rng.randint(len(ms))]
        return x[:m.start(2)] + str(rng.randint(50, 500)) + x[m.end(2):]
    if kind == 6:
        ms = list(re.finditer(r"(\d+) (\d+) (\d+)\n", x))
        m = ms[rng.randint(len(ms))]
        return x[:m.start()] + "%d %d %d\n" % (rng.randint(-5, 400), rng.randint(0, 3), rng.randint(1, 4)) + x[m.end():]
    i = rng.randint(len(x))
    return x[:i] + "<" + x[i:]


def test_mutated_scenes_load_or_fail_with_a_message_but_never_crash(tmp_path):
    from scenes_util import random_scene
    rng = np.random.RandomState(7)
    files = []
    for seed in range(4):
        x0 = open(random_scene(str(tmp_path / "s"), seed, textures=True, extras=True)).read()
        for k in range(40):
            p = str(tmp_path / "s" / ("m_%d_%d.xml" % (seed, k)))
            with open(p, "w") as f:
                f.write(_mutate(x0, rng))
            files.append(p)
    child = str(tmp_path / "child.py")
    with open(child, "w") as f:
        f.write(CHILD)
    r = subprocess.run([sys.executable, child] + files, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = r.stdout.decode(errors="replace")
    begun = [l.split()[1] for l in out.splitlines() if l.startswith("BEGIN")]
    assert r.returncode == 0, "the loader took the process down (rc %d) on %s" % (r.returncode, begun[-1] if begun else "?")
    outcomes = [l.split()[0] for l in out.splitlines() if l.startswith(("LOADED", "ERROR"))]
    assert len(outcomes) == len(files)
    assert outcomes.count("ERROR") >= 40 and outcomes.count("LOADED") >= 20          # both paths are exercised
    for l in out.splitlines():
        if l.startswith("ERROR"):
            assert "failed:" in l and len(l.split("failed:")[1].strip()) > 0, l            # a message, not an empty string
