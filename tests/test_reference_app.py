"""The reference's own main.c / parser / writers, compiled untouched and linked against
libfmgi_cuda.so instead of global_illumination_cl.c + libOpenCL, with libpng replaced
by csrc/png_standin.c.  build/globalIllumination is built by `make app` where /root/reference
exists and travels to the GPU box as a prebuilt file."""
import ctypes as C
import hashlib
import os
import subprocess
import time
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

PKG = ROOT / "flatmatch-global-illumination_b200"
APP = PKG / "build" / "globalIllumination"
LUMA = (0.2126, 0.7152, 0.0722)


def run_app(tmp_path, env=None):
    (tmp_path / "tiles").mkdir(exist_ok=True)
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.perf_counter()
    r = subprocess.run([str(APP), str(GOLDEN / "example.png")], cwd=tmp_path, env=e, capture_output=True, text=True)
    return r, time.perf_counter() - t0


def png_lib():
    L = C.CDLL(str(PKG / "lib" / "libfmgi_png.so"))
    L.read_png_file.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.POINTER(C.c_uint8))]
    L.write_png_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return L


def test_png_standin_matches_pil(tmp_path):
    """png_helper.h:13-15 contract: 8-bit RGB / RGBA, decoded pixels equal to an independent decoder."""
    from PIL import Image

    L = png_lib()
    rng = np.random.default_rng(0)
    for channels, ctype, mode in ((3, 2, "RGB"), (4, 6, "RGBA")):
        for w, h in ((1, 1), (7, 5), (64, 33)):
            px = rng.integers(0, 256, (h, w, channels), dtype=np.uint8)
            ours = tmp_path / f"ours_{channels}_{w}.png"
            L.write_png_file(str(ours).encode(), w, h, ctype, px.ctypes.data)
            assert np.array_equal(np.asarray(Image.open(ours).convert(mode)), px)
            theirs = tmp_path / f"pil_{channels}_{w}.png"
            Image.fromarray(px, mode).save(theirs)           # PIL picks adaptive filters: exercises all five
            ww, hh, ct = C.c_int(), C.c_int(), C.c_int()
            buf = C.POINTER(C.c_uint8)()
            L.read_png_file(str(theirs).encode(), C.byref(ww), C.byref(hh), C.byref(ct), C.byref(buf))
            assert (ww.value, hh.value, ct.value) == (w, h, ctype)
            got = np.ctypeslib.as_array(buf, shape=(h, w, channels))
            assert np.array_equal(got, px)
    # the reference's sample layout decodes to the pixels the parser fixtures were made from
    ww, hh, ct = C.c_int(), C.c_int(), C.c_int()
    buf = C.POINTER(C.c_uint8)()
    L.read_png_file(str(GOLDEN / "example.png").encode(), C.byref(ww), C.byref(hh), C.byref(ct), C.byref(buf))
    assert (ww.value, hh.value) == (640, 440)
    ch = 3 if ct.value == 2 else 4
    got = np.ctypeslib.as_array(buf, shape=(440, 640, ch))
    assert np.array_equal(got[..., :3], np.asarray(Image.open(GOLDEN / "example.png").convert("RGB")))


def test_reference_app_writes_identical_json_and_fails_loudly_without_gpu(tmp_path, facts, fmgi):
    if not APP.exists():
        pytest.skip("build/globalIllumination not built (needs /root/reference at build time)")
    if fmgi.lib().fmgi_device_count() > 0:
        pytest.skip("a GPU is present; covered by the gpu test")
    r, _ = run_app(tmp_path)
    assert r.returncode == 1 and "[Err]" in r.stdout            # no CPU fallback behind the boundary
    assert hashlib.sha256((tmp_path / "geometry.json").read_bytes()).hexdigest() == facts["sha256_geometry_json"]
    assert hashlib.sha256((tmp_path / "collisionMap.json").read_bytes()).hexdigest() == facts["sha256_collision_map_json"]


def tone_map_tile(wall, atlas, spa):
    """main.c:68-79 normalisation + saveAs_core (rectangle.c:293-336) for one wall, tintExtra = 0."""
    b, tw, th = (int(x) for x in wall["lightmapSetup"][:3])
    lw = np.float32(np.sqrt(np.float32((wall["width"][:3].astype(np.float32) ** 2).sum())))
    lh = np.float32(np.sqrt(np.float32((wall["height"][:3].astype(np.float32) ** 2).sum())))
    tiles_per_sample = np.float32(np.float32(tw * th) / np.float32(np.float32(lw * lh) * np.float32(spa)))
    rgb = (atlas[b:b + tw * th, :3].astype(np.float32) * np.float32(0.35 * float(tiles_per_sample))).astype(np.float32)
    lum = (LUMA[0] * rgb[:, 0].astype(np.float64) + LUMA[1] * rgb[:, 1] + LUMA[2] * rgb[:, 2]).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        perc = (1 - np.exp((np.float32(-2) * lum).astype(np.float64))).astype(np.float32)
        out = rgb * (perc / lum)[:, None]
    px = np.clip(np.nan_to_num(out * np.float32(255), nan=0.0), 0, 255).astype(np.uint8)
    if wall["pos"][2] == 0 and wall["width"][2] == 0 and wall["height"][2] == 0:       # floor tint, rectangle.c:317-334
        px[:, 1] = (px[:, 1] * 0.95).astype(np.uint8)
        px[:, 2] = (px[:, 2] * 0.9).astype(np.uint8)
    return px.reshape(th, tw, 3), lum.reshape(th, tw)


@pytest.mark.gpu
def test_reference_app_end_to_end(tmp_path, facts, fmgi, scene):
    """example.png -> tiles/tile_N.png through the untouched main.c with the default 1e8 photons/m^2
    (1.54e9 photons, 8 bounces); tiles equal an independent bake + tone-map of the same photon set."""
    from PIL import Image

    if not APP.exists():
        pytest.fail("build/globalIllumination missing: run `make -C flatmatch-global-illumination_b200 app` where /root/reference exists")
    r, secs = run_app(tmp_path, {"FMGI_SEED": "1", "FMGI_STATS": "1"})
    assert r.returncode == 0, r.stdout + r.stderr
    print(f"example.png bake wall time {secs:.2f} s\\n" + "\\n".join(l for l in r.stdout.splitlines() if "[INF]" in l))
    assert hashlib.sha256((tmp_path / "geometry.json").read_bytes()).hexdigest() == facts["sha256_geometry_json"]
    assert hashlib.sha256((tmp_path / "collisionMap.json").read_bytes()).hexdigest() == facts["sha256_collision_map_json"]
    spa = 100_000_000                                   # main.c:58
    tex = fmgi.aligned_texels(scene.num_texels)
    st = fmgi.bake(fmgi.make_geometry(scene.walls, scene.windows, scene.lights, tex), spa, seed=1, max_depth=8)
    assert st["photons"] == sum(scene.photon_counts(spa))
    total = same = 0
    worst = 0
    for i, wall in enumerate(scene.walls):
        got = np.asarray(Image.open(tmp_path / "tiles" / f"tile_{i}.png").convert("RGB")).astype(np.int32)
        want, lum = tone_map_tile(wall, tex, spa)
        assert got.shape == want.shape
        lit = lum > 0
        d = np.abs(got - want.astype(np.int32))[lit]
        total += d.size
        same += int((d == 0).sum())
        worst = max(worst, int(d.max()) if d.size else 0)
    assert worst <= 1 and same / total > 0.995, (worst, same / total)
