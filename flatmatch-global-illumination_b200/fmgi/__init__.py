"""Python host-side mirror of libfmgi_cuda.so (include/fmgi.h).

This is plumbing for harnesses (tests/, bench.py): ctypes over the C ABI, numpy for host
buffers, optional torch tensors for device-resident atlases and streams.  All compute happens
in the CUDA library; there is no Python or CPU fallback — if the library is missing or no GPU
is visible, the calls raise.

Names follow the reference: ``Geometry`` (geometry.h:7-15), ``Rectangle`` records
(rectangle.h:19-26), ``perform_global_illumination_cl`` (global_illumination_cl.h:10).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent.parent
# FMGI_LIB selects another build of the library, e.g. lib/libfmgi_cuda_checked.so (bounds-checked kernels)
LIB_PATH = Path(os.environ.get("FMGI_LIB") or (_PKG / "lib" / "libfmgi_cuda.so"))

# rectangle.h:19-26 — 80 bytes, 16-byte aligned
RECT_DTYPE = np.dtype(
    {
        "names": ["pos", "width", "height", "n", "lightmapSetup"],
        "formats": [("<f4", 4), ("<f4", 4), ("<f4", 4), ("<f4", 4), ("<i4", 4)],
        "offsets": [0, 16, 32, 48, 64],
        "itemsize": 80,
    }
)

DEPOSIT_VEC4, DEPOSIT_SCALAR, DEPOSIT_WARP_AGG = 0, 1, 2
TIER_AUTO, TIER_SOUP, TIER_GRID, TIER_ROOMS = 0, 1, 2, 4


class Geometry(C.Structure):
    """geometry.h:7-15 (same bytes as fmgi_geometry)."""

    _fields_ = [
        ("windows", C.c_void_p), ("lights", C.c_void_p), ("walls", C.c_void_p), ("boxWalls", C.c_void_p),
        ("numWindows", C.c_int32), ("numLights", C.c_int32), ("numWalls", C.c_int32), ("numBoxWalls", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32),
        ("startingPositionX", C.c_float), ("startingPositionY", C.c_float),
        ("numTexels", C.c_int32), ("texels", C.c_void_p),
    ]


class Options(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("max_depth", C.c_int32), ("seed", C.c_uint32), ("num_gpus", C.c_int32),
        ("shard", C.c_int32), ("num_shards", C.c_int32), ("tier", C.c_int32), ("deposit", C.c_int32),
        ("device", C.c_int32), ("count_tests", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("photons", C.c_uint64), ("rays", C.c_uint64), ("deposits", C.c_uint64), ("mirror_bounces", C.c_uint64),
        ("rect_tests", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("trace_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("reduce_ms", C.c_double),
        ("total_ms", C.c_double),
        ("num_gpus", C.c_int32), ("tier", C.c_int32), ("num_sms", C.c_int32), ("sm_clock_khz", C.c_int32),
        ("init_ms", C.c_double), ("prepare_ms", C.c_double), ("grid_build_ms", C.c_double), ("upload_ms", C.c_double),
        ("pool_rays", C.c_int32), ("bounds_violations", C.c_int32),
    ]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_}
        if d["bounds_violations"] > 0:
            raise FmgiError(f"bounds-checked build: {d['bounds_violations']} index violations")
        return d


EXPORTS = [
    "performGlobalIlluminationCl", "fmgi_default_options", "fmgi_last_error", "fmgi_version", "fmgi_source_hash",
    "fmgi_device_count",
    "fmgi_release_cache", "fmgi_cached_bytes", "fmgi_tile_bytes", "fmgi_scene_tonemap", "fmgi_bake_tiles",
    "fmgi_ambient_occlusion", "fmgi_scene_ambient_occlusion", "fmgi_geosphere",
    "fmgi_bake", "fmgi_scene_create", "fmgi_scene_destroy", "fmgi_scene_trace", "fmgi_scene_sync",
    "fmgi_scene_photon_count", "fmgi_probe_closest_hit", "fmgi_probe_tile_ids", "fmgi_probe_philox",
    "fmgi_probe_sample_dirs", "fmgi_probe_paths", "fmgi_probe_deposit_peak", "fmgi_probe_philox2x32",
    "fmgi_probe_grid_table", "fmgi_tile_png_bytes", "fmgi_scene_tiles_png", "fmgi_bake_tiles_png",
]

_lib = None


class FmgiError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads lib/libfmgi_cuda.so (built by __graft_entry__.build() / make).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FmgiError(f"{LIB_PATH} not built - run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(str(LIB_PATH))
    L.fmgi_last_error.restype = C.c_char_p
    L.fmgi_version.restype = C.c_char_p
    L.fmgi_source_hash.restype = C.c_char_p
    L.fmgi_cached_bytes.restype = C.c_uint64
    L.fmgi_default_options.argtypes = [C.POINTER(Options)]
    L.performGlobalIlluminationCl.restype = None
    L.performGlobalIlluminationCl.argtypes = [C.POINTER(Geometry), C.c_int]
    L.fmgi_bake.argtypes = [C.POINTER(Geometry), C.c_int, C.POINTER(Options), C.POINTER(Stats)]
    L.fmgi_scene_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_int, C.c_int, C.POINTER(Options)]
    L.fmgi_scene_destroy.restype = None
    L.fmgi_scene_destroy.argtypes = [C.c_void_p]
    L.fmgi_scene_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Options), C.c_void_p]
    L.fmgi_scene_sync.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.fmgi_scene_photon_count.restype = C.c_uint64
    L.fmgi_scene_photon_count.argtypes = [C.c_void_p, C.c_int, C.POINTER(Options)]
    L.fmgi_tile_bytes.restype = C.c_uint64
    L.fmgi_tile_bytes.argtypes = [C.c_void_p, C.c_int]
    L.fmgi_scene_tonemap.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.fmgi_bake_tiles.argtypes = [C.POINTER(Geometry), C.c_int, C.POINTER(Options), C.c_int, C.c_void_p, C.POINTER(Stats)]
    L.fmgi_bake_tiles_png.argtypes = [C.POINTER(Geometry), C.c_int, C.POINTER(Options), C.c_int, C.c_void_p, C.POINTER(Stats)]
    L.fmgi_tile_png_bytes.restype = C.c_uint64
    L.fmgi_tile_png_bytes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.fmgi_scene_tiles_png.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.fmgi_ambient_occlusion.argtypes = [C.POINTER(Geometry), C.POINTER(Options)]
    L.fmgi_geosphere.argtypes = [C.c_int, C.c_void_p, C.c_int]
    L.fmgi_scene_ambient_occlusion.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.fmgi_probe_deposit_peak.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
    L.fmgi_probe_closest_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.fmgi_probe_tile_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.fmgi_probe_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.fmgi_probe_philox2x32.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.fmgi_probe_grid_table.restype = C.c_int64
    L.fmgi_probe_grid_table.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.fmgi_probe_sample_dirs.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_void_p]
    L.fmgi_probe_paths.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise FmgiError(f"fmgi error {rc}: {lib().fmgi_last_error().decode()}")


def options(**kw) -> Options:
    o = Options()
    lib().fmgi_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k}")
        setattr(o, k, v)
    return o


def aligned_rects(arr) -> np.ndarray:
    """16-byte-aligned copy of a Rectangle table (Rectangle is __attribute__((aligned(16))))."""
    arr = np.asarray(arr, dtype=RECT_DTYPE)
    n = len(arr)
    raw = np.zeros(80 * max(n, 1) + 16, dtype=np.uint8)
    off = (-raw.ctypes.data) % 16
    out = raw[off: off + 80 * n].view(RECT_DTYPE)
    out[...] = arr
    return out


def aligned_texels(num_texels: int) -> np.ndarray:
    """Zeroed (numTexels, 4) float32 atlas, 16-byte aligned like parseLayout.c:526-533 allocates it."""
    raw = np.zeros(16 * max(num_texels, 1) + 16, dtype=np.uint8)
    off = (-raw.ctypes.data) % 16
    return raw[off: off + 16 * num_texels].view("<f4").reshape(num_texels, 4)


def make_geometry(walls, windows, lights, texels, box_walls=None) -> Geometry:
    """A reference-layout Geometry over caller-owned, 16-byte-aligned numpy buffers."""
    g = Geometry()
    for name, arr in (("walls", walls), ("windows", windows), ("lights", lights)):
        assert arr.dtype == RECT_DTYPE and arr.ctypes.data % 16 == 0, name
    assert texels.dtype == np.float32 and texels.ndim == 2 and texels.shape[1] == 4 and texels.flags.c_contiguous
    g.walls, g.numWalls = walls.ctypes.data, len(walls)
    g.windows, g.numWindows = windows.ctypes.data, len(windows)
    g.lights, g.numLights = lights.ctypes.data, len(lights)
    if box_walls is not None and len(box_walls):
        g.boxWalls, g.numBoxWalls = box_walls.ctypes.data, len(box_walls)
    g.numTexels = texels.shape[0]
    g.texels = texels.ctypes.data
    return g


def perform_global_illumination_cl(geo: Geometry, num_samples_per_area: int) -> None:
    """The reference boundary: geo.texels += raw deposits (global_illumination_cl.h:10)."""
    lib().performGlobalIlluminationCl(C.byref(geo), int(num_samples_per_area))


def bake(geo: Geometry, num_samples_per_area: int, **opts) -> dict:
    """fmgi_bake: host-buffer bake with options; returns the counters."""
    o = options(**opts)
    st = Stats()
    _check(lib().fmgi_bake(C.byref(geo), int(num_samples_per_area), C.byref(o), C.byref(st)))
    return st.as_dict()


def bake_tiles(geo: Geometry, walls: np.ndarray, num_samples_per_area: int, tint_extra: int = 0, **opts):
    """fmgi_bake_tiles: bake, then normalise + tone-map + pack on the device (main.c:68-79 +
    rectangle.c:293-336).  Returns (packed RGB bytes in wall order, counters)."""
    o = options(**opts)
    st = Stats()
    n = int(lib().fmgi_tile_bytes(walls.ctypes.data, len(walls)))
    out = np.zeros(max(n, 1), dtype=np.uint8)
    _check(lib().fmgi_bake_tiles(C.byref(geo), int(num_samples_per_area), C.byref(o), int(tint_extra),
                                 out.ctypes.data, C.byref(st)))
    return out[:n], st.as_dict()


def tile_png_layout(walls: np.ndarray):
    """(total bytes, offsets[num_walls + 1]) of the PNG files fmgi_bake_tiles_png / fmgi_scene_tiles_png produce."""
    off = np.zeros(len(walls) + 1, dtype=np.uint64)
    n = int(lib().fmgi_tile_png_bytes(walls.ctypes.data, len(walls), off.ctypes.data))
    return n, off


def bake_tiles_png(geo: Geometry, walls: np.ndarray, num_samples_per_area: int, tint_extra: int = 0, **opts):
    """fmgi_bake_tiles_png: bake, tone-map and assemble one PNG file per wall on the device.
    Returns (buffer, offsets, counters); wall i's file is buffer[offsets[i]:offsets[i + 1]]."""
    o = options(**opts)
    st = Stats()
    n, off = tile_png_layout(walls)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    _check(lib().fmgi_bake_tiles_png(C.byref(geo), int(num_samples_per_area), C.byref(o), int(tint_extra),
                                     out.ctypes.data, C.byref(st)))
    return out[:n], off, st.as_dict()


def geosphere(iterations: int = 4) -> np.ndarray:
    n = lib().fmgi_geosphere(int(iterations), None, 0)
    out = np.zeros((n, 3), dtype=np.float32)
    lib().fmgi_geosphere(int(iterations), out.ctypes.data_as(C.c_void_p), n)
    return out


def ambient_occlusion(geo: Geometry, **opts) -> None:
    """fmgi_ambient_occlusion: performAmbientOcclusionNative (photonmap.c:480) on the GPU."""
    o = options(**opts)
    _check(lib().fmgi_ambient_occlusion(C.byref(geo), C.byref(o)))


class DeviceScene:
    """fmgi_scene: collider + emitter tables resident on one GPU."""

    def __init__(self, walls, windows, lights, num_texels: int, device: int = 0, tier: int = TIER_AUTO):
        self._h = C.c_void_p()
        self.walls = aligned_rects(walls)
        self.windows = aligned_rects(windows)
        self.lights = aligned_rects(lights)
        self.num_texels = int(num_texels)
        self.device = device
        o = options(device=device, tier=tier)
        _check(lib().fmgi_scene_create(C.byref(self._h), self.walls.ctypes.data, len(self.walls),
                                       self.windows.ctypes.data, len(self.windows), self.lights.ctypes.data,
                                       len(self.lights), self.num_texels, C.byref(o)))

    def close(self):
        if self._h:
            lib().fmgi_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def photon_count(self, spa: int, **opts) -> int:
        o = options(device=self.device, **opts)
        return int(lib().fmgi_scene_photon_count(self._h, int(spa), C.byref(o)))

    def trace(self, atlas_ptr: int, spa: int, stream: int = 0, **opts) -> None:
        """Enqueue the bake on a CUDA stream, accumulating into the device atlas at atlas_ptr."""
        o = options(device=self.device, **opts)
        _check(lib().fmgi_scene_trace(self._h, C.c_void_p(atlas_ptr), int(spa), C.byref(o), C.c_void_p(stream)))

    def sync(self) -> dict:
        st = Stats()
        _check(lib().fmgi_scene_sync(self._h, C.byref(st)))
        return st.as_dict()

    def tonemap(self, atlas_ptr: int, spa: int, rgb_ptr: int, tint_extra: int = 0, stream: int = 0) -> None:
        """fmgi_scene_tonemap: RAW device atlas -> packed RGB tiles on the device."""
        _check(lib().fmgi_scene_tonemap(self._h, C.c_void_p(atlas_ptr), int(spa), int(tint_extra),
                                        C.c_void_p(rgb_ptr), C.c_void_p(stream)))

    def tiles_png(self, atlas_ptr: int, spa: int, png_ptr: int, tint_extra: int = 0, stream: int = 0) -> None:
        """fmgi_scene_tiles_png: RAW device atlas -> one complete PNG file per wall, on the device."""
        _check(lib().fmgi_scene_tiles_png(self._h, C.c_void_p(atlas_ptr), int(spa), int(tint_extra),
                                          C.c_void_p(png_ptr), C.c_void_p(stream)))

    def tile_bytes(self) -> int:
        return int(lib().fmgi_tile_bytes(self.walls.ctypes.data, len(self.walls)))

    def grid_table(self) -> np.ndarray:
        """The device's floor-plan grid table T as (records, 8) uint32 words."""
        n = int(lib().fmgi_probe_grid_table(self._h, None, 0))
        if n < 0:
            raise FmgiError(lib().fmgi_last_error().decode())
        out = np.zeros((max(n, 1), 8), dtype=np.uint32)
        lib().fmgi_probe_grid_table(self._h, out.ctypes.data, n)
        return out[:n]

    # -- parity probes ---------------------------------------------------------------------------
    def closest_hit(self, origins, dirs):
        o = np.ascontiguousarray(origins, dtype=np.float32)
        d = np.ascontiguousarray(dirs, dtype=np.float32)
        n = len(o)
        idx = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float32)
        _check(lib().fmgi_probe_closest_hit(self._h, o.ctypes.data, d.ctypes.data, n, idx.ctypes.data, t.ctypes.data))
        return idx, t

    def tile_ids(self, rect_index, points):
        r = np.ascontiguousarray(rect_index, dtype=np.int32)
        p = np.ascontiguousarray(points, dtype=np.float32)
        out = np.empty(len(r), dtype=np.int32)
        _check(lib().fmgi_probe_tile_ids(self._h, r.ctypes.data, p.ctypes.data, len(r), out.ctypes.data))
        return out

    def paths(self, emitter_index: int, max_depth: int, seed: int, first: int, count: int):
        out = np.empty((count, max_depth), dtype=np.int32)
        _check(lib().fmgi_probe_paths(self._h, emitter_index, max_depth, seed, first, count, out.ctypes.data))
        return out


def philox(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    _check(lib().fmgi_probe_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data))
    return out


def philox2x32(ctr, key: int) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    out = np.empty(2, dtype=np.uint32)
    _check(lib().fmgi_probe_philox2x32(c.ctypes.data, int(key) & 0xFFFFFFFF, out.ctypes.data))
    return out


def deposit_peak(num_texels: int, num_deposits: int = 400_000_000, device: int = 0) -> float:
    """Deposits per second of the bare deposit instruction at uniform-random texels (fmgi_probe_deposit_peak)."""
    out = C.c_double(0.0)
    _check(lib().fmgi_probe_deposit_peak(C.c_uint64(int(num_texels)), C.c_uint64(int(num_deposits)), int(device),
                                         C.byref(out)))
    return out.value


def sample_dirs(normal, sky: bool, seed: int, n: int) -> np.ndarray:
    nn = np.ascontiguousarray(normal, dtype=np.float32)
    out = np.empty((n, 3), dtype=np.float32)
    _check(lib().fmgi_probe_sample_dirs(nn.ctypes.data, int(sky), seed, n, out.ctypes.data))
    return out
