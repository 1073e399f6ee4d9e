"""Output path of the host library: band-parallel PNG encoder, RGBE writer, asynchronous writer queue.
Replaces the main-thread stbi_write_png / stbi_write_hdr of the reference (main.cpp:186-195).  CPU only."""
import ctypes as C
import os
import time

import numpy as np
import pytest

from dtb200 import capi

PIL = pytest.importorskip("PIL.Image")


def _image(w, h, seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 255 // max(1, w - 1)), (yy * 255 // max(1, h - 1)), ((xx ^ yy) & 255)], axis=-1).astype(np.uint8)
    img[h // 3: h // 2] = rng.integers(0, 256, size=(h // 2 - h // 3, w, 3), dtype=np.uint8)       # an incompressible band
    return np.ascontiguousarray(img)


def _decode_rgbe(path):
    raw = open(path, "rb").read()
    head, _, rest = raw.partition(b"\n\n")
    assert head.startswith(b"#?RADIANCE") and b"32-bit_rle_rgbe" in head
    dims, _, body = rest.partition(b"\n")
    t = dims.split()
    h, w = int(t[1]), int(t[3])
    px = np.frombuffer(body, dtype=np.uint8).reshape(h, w, 4).astype(np.float64)
    scale = np.where(px[..., 3] > 0, np.ldexp(1.0, (px[..., 3] - 136).astype(int)), 0.0)
    return px[..., :3] * scale[..., None]


@pytest.mark.parametrize("w,h,threads", [(1, 1, 4), (37, 19, 3), (640, 360, 1), (640, 360, 8), (1920, 1080, 0)])
def test_parallel_png_decodes_to_the_same_pixels(tmp_path, w, h, threads):
    lib = capi.load_dthost()
    img = _image(w, h, seed=w + h)
    p = str(tmp_path / "par.png").encode()
    s = str(tmp_path / "ser.png").encode()
    assert lib.dth_write_png_parallel(p, w, h, img.ctypes.data, threads) == 0, lib.dth_last_error()
    assert lib.dth_write_png(s, w, h, img.ctypes.data) == 0
    a = np.asarray(PIL.open(p.decode()).convert("RGB"))
    b = np.asarray(PIL.open(s.decode()).convert("RGB"))
    assert a.shape == (h, w, 3) and (a == img).all() and (b == img).all()
    assert os.path.getsize(p.decode()) < 1.1 * os.path.getsize(s.decode()) + 4096       # band boundaries cost a few bytes, not a different compression


def test_writer_queue_is_asynchronous_and_complete(tmp_path):
    lib = capi.load_dthost()
    w, h = 1920, 1080
    frames = [_image(w, h, seed=k) for k in range(4)]
    hdr = (np.random.default_rng(5).random((h, w, 3)) * 40.0).astype(np.float32)
    hdr[0, 0] = 0.0
    wr = lib.dth_writer_create(2, 2)
    assert wr
    t0 = time.perf_counter()
    for k, f in enumerate(frames):
        assert lib.dth_writer_submit_png(wr, str(tmp_path / ("f%d.png" % k)).encode(), w, h, f.ctypes.data) == 0
        f[:] = 0                                   # the queue copied the pixels: the caller may reuse its buffer at once
    assert lib.dth_writer_submit_hdr(wr, str(tmp_path / "f.hdr").encode(), w, h, hdr.ctypes.data) == 0
    t_submit = time.perf_counter() - t0
    busy = C.c_double(0.0)
    assert lib.dth_writer_wait(wr, C.byref(busy)) == 0, lib.dth_last_error()
    t_total = time.perf_counter() - t0
    lib.dth_writer_destroy(wr)
    for k in range(4):
        assert (np.asarray(PIL.open(str(tmp_path / ("f%d.png" % k))).convert("RGB")) == _image(w, h, seed=k)).all()
    dec = _decode_rgbe(str(tmp_path / "f.hdr"))
    assert np.abs(dec - hdr).max() <= np.abs(hdr).max() / 128.0 and (dec[0, 0] == 0).all()      # 8-bit mantissa shared per pixel
    assert busy.value > 0.0
    assert t_submit < 0.5 * t_total, (t_submit, t_total)      # submit returned long before the files were encoded
    print("submit %.3f s, total %.3f s, encode time on workers %.3f s" % (t_submit, t_total, busy.value))


def test_writer_reports_unwritable_path(tmp_path):
    lib = capi.load_dthost()
    wr = lib.dth_writer_create(1, 1)
    img = _image(8, 8)
    assert lib.dth_writer_submit_png(wr, str(tmp_path / "no_such_dir" / "x.png").encode(), 8, 8, img.ctypes.data) == 0
    assert lib.dth_writer_wait(wr, None) != 0
    assert b"cannot write" in lib.dth_last_error()
    assert lib.dth_writer_submit_png(wr, str(tmp_path / "ok.png").encode(), 8, 8, img.ctypes.data) == 0
    assert lib.dth_writer_wait(wr, None) == 0       # the error was consumed; the queue keeps working
    lib.dth_writer_destroy(wr)
    assert lib.dth_writer_submit_png(None, b"x", 8, 8, img.ctypes.data) != 0 and lib.dth_writer_submit_png(None, None, 0, 0, None) != 0
