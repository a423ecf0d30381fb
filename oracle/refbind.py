"""TEST INFRASTRUCTURE — ctypes bindings for the two CPU checkers.

* ``RefLib``    -> oracle/_ref/libfmgi_ref{,_depth}.so: the UNMODIFIED reference sources
                   (photonmap.c, rectangle.c, vector3_cl.c, parseLayout.c ...) compiled by
                   oracle/Makefile, driven through oracle/ref_shim.c.
* ``OracleLib`` -> oracle/libfmgi_oracle.so: our own C restatement (oracle/photon_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (flatmatch-global-illumination_b200/) never does.

Struct layouts mirror the reference: Rectangle rectangle.h:19-26 (80 B, 16-aligned),
Geometry geometry.h:7-15 (80 B on x86-64), Vector3 = cl_float4 vector3_cl.h:14.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"

# numpy view of one Rectangle (rectangle.h:19-26)
RECT_DTYPE = np.dtype(
    {
        "names": ["pos", "width", "height", "n", "lightmapSetup"],
        "formats": [("<f4", 4), ("<f4", 4), ("<f4", 4), ("<f4", 4), ("<i4", 4)],
        "offsets": [0, 16, 32, 48, 64],
        "itemsize": 80,
    }
)


class Geometry(C.Structure):
    """geometry.h:7-15."""

    _fields_ = [
        ("windows", C.c_void_p),
        ("lights", C.c_void_p),
        ("walls", C.c_void_p),
        ("boxWalls", C.c_void_p),
        ("numWindows", C.c_int32),
        ("numLights", C.c_int32),
        ("numWalls", C.c_int32),
        ("numBoxWalls", C.c_int32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("startingPositionX", C.c_float),
        ("startingPositionY", C.c_float),
        ("numTexels", C.c_int32),
        ("texels", C.c_void_p),
    ]


assert C.sizeof(Geometry) == 80


def _rects_from_ptr(ptr: int, n: int) -> np.ndarray:
    if n == 0:
        return np.zeros(0, dtype=RECT_DTYPE)
    buf = (C.c_char * (80 * n)).from_address(ptr)
    return np.frombuffer(buf, dtype=RECT_DTYPE, count=n).copy()


def aligned_rects(arr: np.ndarray) -> np.ndarray:
    """Copy of a RECT_DTYPE array whose storage is 16-byte aligned (Rectangle is aligned(16))."""
    n = len(arr)
    raw = np.zeros(80 * max(n, 1) + 16, dtype=np.uint8)
    off = (-raw.ctypes.data) % 16
    out = raw[off : off + 80 * n].view(RECT_DTYPE)
    out[...] = arr
    return out


def aligned_texels(num_texels: int) -> np.ndarray:
    raw = np.zeros(16 * max(num_texels, 1) + 16, dtype=np.uint8)
    off = (-raw.ctypes.data) % 16
    return raw[off : off + 16 * num_texels].view("<f4").reshape(num_texels, 4)


class Scene:
    """Plain-numpy copy of a Geometry: rect tables + atlas size.  Travels as an .npz fixture."""

    def __init__(self, walls, windows, lights, num_texels, box_walls=None, meta=None):
        self.walls = aligned_rects(np.asarray(walls, dtype=RECT_DTYPE))
        self.windows = aligned_rects(np.asarray(windows, dtype=RECT_DTYPE))
        self.lights = aligned_rects(np.asarray(lights, dtype=RECT_DTYPE))
        self.box_walls = aligned_rects(
            np.asarray(box_walls if box_walls is not None else np.zeros(0, RECT_DTYPE), dtype=RECT_DTYPE)
        )
        self.num_texels = int(num_texels)
        self.meta = dict(meta or {})

    # -- persistence -----------------------------------------------------------------
    def save(self, path):
        np.savez_compressed(
            path,
            walls=self.walls.view(np.uint8),
            windows=self.windows.view(np.uint8),
            lights=self.lights.view(np.uint8),
            box_walls=self.box_walls.view(np.uint8),
            num_texels=np.int64(self.num_texels),
            **{f"meta_{k}": np.asarray(v) for k, v in self.meta.items()},
        )

    @classmethod
    def load(cls, path) -> "Scene":
        z = np.load(path)
        meta = {k[5:]: z[k].tolist() for k in z.files if k.startswith("meta_")}
        return cls(
            z["walls"].view(RECT_DTYPE),
            z["windows"].view(RECT_DTYPE),
            z["lights"].view(RECT_DTYPE),
            int(z["num_texels"]),
            z["box_walls"].view(RECT_DTYPE),
            meta,
        )

    # -- helpers -----------------------------------------------------------------------
    def geometry(self, texels: np.ndarray) -> Geometry:
        """A reference-layout Geometry over this scene's tables and the given atlas."""
        assert texels.dtype == np.float32 and texels.shape == (self.num_texels, 4)
        assert texels.ctypes.data % 16 == 0
        g = Geometry()
        g.windows = self.windows.ctypes.data
        g.lights = self.lights.ctypes.data
        g.walls = self.walls.ctypes.data
        g.boxWalls = self.box_walls.ctypes.data
        g.numWindows, g.numLights = len(self.windows), len(self.lights)
        g.numWalls, g.numBoxWalls = len(self.walls), len(self.box_walls)
        g.width = int(self.meta.get("width", 0))
        g.height = int(self.meta.get("height", 0))
        g.startingPositionX = float(self.meta.get("startX", 0.0))
        g.startingPositionY = float(self.meta.get("startY", 0.0))
        g.numTexels = self.num_texels
        g.texels = texels.ctypes.data
        return g

    def emitters(self):
        """(rect, is_window) in the order the reference walks them (photonmap.c:412-431)."""
        return [(w, 1) for w in self.windows] + [(l, 0) for l in self.lights]

    def photon_counts(self, spa: int):
        """photonmap.c:414-418: area = |w|*|h| in float; N = (uint64)(int spa * float area)."""
        out = []
        for r, _ in self.emitters():
            w, h = r["width"], r["height"]
            lw = np.sqrt(np.float32(w[0] * w[0] + w[1] * w[1]) + np.float32(w[2] * w[2]), dtype=np.float32)
            lh = np.sqrt(np.float32(h[0] * h[0] + h[1] * h[1]) + np.float32(h[2] * h[2]), dtype=np.float32)
            area = np.float32(lw * lh)
            out.append(int(np.float32(np.float32(spa) * area)))
        return out

    def base_texel_mask(self) -> np.ndarray:
        """True for base-level (mip 0) texels — the only ones the photon modes write."""
        m = np.zeros(self.num_texels, dtype=bool)
        for r in self.walls:
            b, w, h = (int(x) for x in r["lightmapSetup"][:3])
            m[b : b + w * h] = True
        return m

    def normalisation(self, spa_total: float) -> np.ndarray:
        """Per-texel factor of main.c:68-79: 0.35 * tiles / (area * spa), base level only."""
        f = np.zeros(self.num_texels, dtype=np.float64)
        for r in self.walls:
            b, w, h = (int(x) for x in r["lightmapSetup"][:3])
            lw = float(np.linalg.norm(r["width"][:3].astype(np.float64)))
            lh = float(np.linalg.norm(r["height"][:3].astype(np.float64)))
            f[b : b + w * h] = 0.35 * (w * h) / (lw * lh * spa_total)
        return f


class RefLib:
    """The compiled reference.  depth != 8 needs the runtime-depth build."""

    def __init__(self, runtime_depth: bool = False):
        name = "libfmgi_ref_depth.so" if runtime_depth else "libfmgi_ref.so"
        path = REF_DIR / name
        if not path.exists():
            raise FileNotFoundError(f"{path} missing - run `make -C oracle ref` where /root/reference exists")
        self.path = path
        self.runtime_depth = runtime_depth
        L = self.lib = C.CDLL(str(path))
        L.fmgi_ref_parse_rgba.restype = C.POINTER(Geometry)
        L.fmgi_ref_parse_rgba.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]
        L.fmgi_ref_collision_map_json.restype = C.c_void_p
        L.fmgi_ref_collision_map_json.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.fmgi_ref_geometry_json.restype = C.c_void_p
        L.fmgi_ref_geometry_json.argtypes = [C.POINTER(Geometry)]
        L.fmgi_ref_free_string.argtypes = [C.c_void_p]
        L.fmgi_ref_free_geometry.argtypes = [C.POINTER(Geometry)]
        L.fmgi_ref_photonmap_native.restype = C.c_double
        L.fmgi_ref_photonmap_native.argtypes = [C.POINTER(Geometry), C.c_int, C.c_uint, C.c_int]
        for fn in (L.fmgi_ref_closest_hit_linear, L.fmgi_ref_closest_hit_bsp):
            fn.restype = None
            fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.fmgi_ref_tile_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.fmgi_ref_sample_dirs.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_int, C.c_void_p]
        L.performAmbientOcclusionNative.restype = None
        L.performAmbientOcclusionNative.argtypes = [C.POINTER(Geometry)]
        L.fmgi_ref_save_tile.restype = None
        L.fmgi_ref_save_tile.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        assert L.fmgi_ref_sizeof_rectangle() == 80 and L.fmgi_ref_sizeof_geometry() == 80
        assert L.fmgi_ref_sizeof_vector3() == 16

    # -- layout parsing ------------------------------------------------------------------
    def parse_rgba(self, pixels: np.ndarray, pixels_per_metre: float = 30.0, tile_size: float = 200.0):
        """pixels: (H, W) uint32 0xAABBGGRR.  Returns (Scene, geometry.json bytes)."""
        px = np.ascontiguousarray(pixels, dtype=np.uint32)
        h, w = px.shape
        with _quiet_stdout():
            g = self.lib.fmgi_ref_parse_rgba(px.ctypes.data, w, h, pixels_per_metre, tile_size)
        gc = g.contents
        sp = self.lib.fmgi_ref_geometry_json(g)
        gjson = C.string_at(sp)
        self.lib.fmgi_ref_free_string(sp)
        scene = Scene(
            _rects_from_ptr(gc.walls, gc.numWalls),
            _rects_from_ptr(gc.windows, gc.numWindows),
            _rects_from_ptr(gc.lights, gc.numLights),
            gc.numTexels,
            _rects_from_ptr(gc.boxWalls, gc.numBoxWalls),
            meta=dict(width=gc.width, height=gc.height, startX=gc.startingPositionX,
                      startY=gc.startingPositionY, scale=pixels_per_metre, tile_size=tile_size),
        )
        self.lib.fmgi_ref_free_geometry(g)
        return scene, gjson

    def collision_map_json(self, pixels: np.ndarray) -> bytes:
        px = np.ascontiguousarray(pixels, dtype=np.uint32)
        h, w = px.shape
        sp = self.lib.fmgi_ref_collision_map_json(px.ctypes.data, w, h)
        out = C.string_at(sp)
        self.lib.fmgi_ref_free_string(sp)
        return out

    # -- the oracle ------------------------------------------------------------------------
    def photonmap_native(self, scene: Scene, spa: int, seed: int = 1, depth: int = 8):
        """performPhotonMappingNative (photonmap.c:408) after srand(seed).
        Returns (raw atlas (numTexels,4) float32, seconds)."""
        tex = aligned_texels(scene.num_texels)
        g = scene.geometry(tex)
        with _quiet_stdout():
            secs = self.lib.fmgi_ref_photonmap_native(C.byref(g), int(spa), int(seed), int(depth))
        return tex, secs

    # -- probes ------------------------------------------------------------------------------
    def closest_hit_linear(self, walls, origins, dirs):
        walls = aligned_rects(walls)
        o = np.ascontiguousarray(origins, dtype=np.float32)
        d = np.ascontiguousarray(dirs, dtype=np.float32)
        n = len(o)
        idx = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float32)
        self.lib.fmgi_ref_closest_hit_linear(walls.ctypes.data, len(walls), o.ctypes.data, d.ctypes.data, n,
                                             idx.ctypes.data, t.ctypes.data)
        return idx, t

    def closest_hit_bsp(self, walls, origins, dirs):
        walls = aligned_rects(walls)
        o = np.ascontiguousarray(origins, dtype=np.float32)
        d = np.ascontiguousarray(dirs, dtype=np.float32)
        n = len(o)
        base = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float32)
        with _quiet_stdout():
            self.lib.fmgi_ref_closest_hit_bsp(walls.ctypes.data, len(walls), o.ctypes.data, d.ctypes.data, n,
                                              base.ctypes.data, t.ctypes.data)
        return base, t

    def tile_ids(self, rect, points):
        r = aligned_rects(np.asarray([rect], dtype=RECT_DTYPE))
        p = np.ascontiguousarray(points, dtype=np.float32)
        out = np.empty(len(p), dtype=np.int32)
        self.lib.fmgi_ref_tile_ids(r.ctypes.data, p.ctypes.data, len(p), out.ctypes.data)
        return out

    def sample_dirs(self, normal, sky: bool, seed: int, n: int):
        nn = np.ascontiguousarray(normal, dtype=np.float32)
        out = np.empty((n, 3), dtype=np.float32)
        self.lib.fmgi_ref_sample_dirs(nn.ctypes.data, int(sky), seed, n, out.ctypes.data)
        return out

    def ambient_occlusion_native(self, scene: Scene, texels: np.ndarray | None = None):
        """performAmbientOcclusionNative (photonmap.c:480): overwrites the base-level texels."""
        tex = aligned_texels(scene.num_texels) if texels is None else texels
        g = scene.geometry(tex)
        with _quiet_stdout():
            self.lib.performAmbientOcclusionNative(C.byref(g))
        return tex

    def geosphere4(self) -> np.ndarray:
        """The reference's 481-direction table (geoSphere.c:148), in its own order."""
        n = C.c_int.in_dll(self.lib, "geoSphere4NumVectors").value
        arr = (C.c_float * (4 * n)).in_dll(self.lib, "geoSphere4")
        return np.frombuffer(arr, dtype=np.float32).reshape(n, 4)[:, :3].copy()

    def save_tile(self, rect, normalised_texels: np.ndarray, tint_extra: int) -> np.ndarray:
        """saveAs (rectangle.c:338) on one wall: the RGB bytes it hands to write_png_file."""
        r = aligned_rects(np.asarray([rect], dtype=RECT_DTYPE))
        tw, th = int(rect["lightmapSetup"][1]), int(rect["lightmapSetup"][2])
        out = np.zeros(tw * th * 3, dtype=np.uint8)
        assert normalised_texels.dtype == np.float32 and normalised_texels.ctypes.data % 16 == 0
        self.lib.fmgi_ref_save_tile(r.ctypes.data, normalised_texels.ctypes.data, int(tint_extra), out.ctypes.data, out.size)
        return out


class _quiet_stdout:
    """The reference printf()s progress (photonmap.c:266-270, 398-404); keep test logs readable."""

    def __enter__(self):
        import sys

        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *a):
        # flush the C stdio buffer into /dev/null before restoring fd 1
        C.CDLL(None).fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._null)
        os.close(self._saved)


def load_layout_png(path) -> np.ndarray:
    """PNG -> (H, W) uint32 0xAABBGGRR, as loadImage produces (image.c:189-217)."""
    from PIL import Image

    im = np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint32)
    return im[..., 0] | (im[..., 1] << 8) | (im[..., 2] << 16) | (im[..., 3] << 24)


class _Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("photons", "rays", "deposits", "mirror_bounces", "rect_tests")]


class OracleLib:
    """oracle/libfmgi_oracle.so — the C restatement (photon_oracle.c)."""

    RNG_LIBC, RNG_PHILOX = 0, 1
    ACCEL_BSP, ACCEL_LINEAR = 0, 1

    def __init__(self):
        path = HERE / "libfmgi_oracle.so"
        if not path.exists():
            raise FileNotFoundError(f"{path} missing - run `make -C oracle oracle`")
        self.path = path
        L = self.lib = C.CDLL(str(path))
        self._libc = C.CDLL(None)
        L.orc_bake.restype = None
        L.orc_bake.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int,
                               C.POINTER(_Stats)]
        L.orc_trace_paths.restype = None
        L.orc_trace_paths.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                      C.c_uint64, C.c_int, C.c_void_p]
        L.orc_closest_hit.restype = None
        L.orc_closest_hit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p]
        L.orc_tile_id.restype = C.c_int
        L.orc_tile_id.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_photon_budget.restype = C.c_uint64
        L.orc_photon_budget.argtypes = [C.c_void_p, C.c_int]
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_philox2x32_10.restype = None
        L.orc_philox2x32_10.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_ambient_occlusion.restype = None
        L.orc_ambient_occlusion.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_tonemap_tiles.restype = None
        L.orc_tonemap_tiles.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]

    def bake(self, scene: Scene, spa: int, depth: int = 8, accel: int = 0, rng: int = 0, seed: int = 1,
             shard: int = 0, num_shards: int = 1, texels: np.ndarray | None = None):
        """Returns (raw atlas (numTexels,4) float32, stats dict).  rng=LIBC calls srand(seed) first."""
        tex = aligned_texels(scene.num_texels) if texels is None else texels
        st = _Stats()
        if rng == self.RNG_LIBC:
            self._libc.srand(C.c_uint(seed))
        self.lib.orc_bake(scene.walls.ctypes.data, len(scene.walls), scene.windows.ctypes.data,
                          len(scene.windows), scene.lights.ctypes.data, len(scene.lights),
                          tex.ctypes.data, int(spa), int(depth), accel, rng, seed, shard, num_shards,
                          C.byref(st))
        return tex, {n: getattr(st, n) for n, _ in _Stats._fields_}

    def trace_paths(self, scene: Scene, emitter_index: int, depth: int, seed: int, first: int, count: int):
        rect, is_window = scene.emitters()[emitter_index]
        r = aligned_rects(np.asarray([rect], dtype=RECT_DTYPE))
        out = np.empty((count, depth), dtype=np.int32)
        self.lib.orc_trace_paths(scene.walls.ctypes.data, len(scene.walls), r.ctypes.data, is_window,
                                 emitter_index, depth, seed, first, count, out.ctypes.data)
        return out

    def closest_hit(self, walls, origins, dirs, accel: int = 1):
        walls = aligned_rects(walls)
        o = np.ascontiguousarray(origins, dtype=np.float32)
        d = np.ascontiguousarray(dirs, dtype=np.float32)
        idx = np.empty(len(o), dtype=np.int32)
        t = np.empty(len(o), dtype=np.float32)
        self.lib.orc_closest_hit(walls.ctypes.data, len(walls), accel, o.ctypes.data, d.ctypes.data, len(o),
                                 idx.ctypes.data, t.ctypes.data)
        return idx, t

    def tile_ids(self, rect, points):
        r = aligned_rects(np.asarray([rect], dtype=RECT_DTYPE))
        p = np.ascontiguousarray(points, dtype=np.float32)
        return np.array([self.lib.orc_tile_id(r.ctypes.data, p[i].ctypes.data) for i in range(len(p))],
                        dtype=np.int32)

    def ambient_occlusion(self, scene: Scene, dirs: np.ndarray, accel: int = 0, texels: np.ndarray | None = None):
        tex = aligned_texels(scene.num_texels) if texels is None else texels
        d = np.ascontiguousarray(dirs, dtype=np.float32)
        self.lib.orc_ambient_occlusion(scene.walls.ctypes.data, len(scene.walls), tex.ctypes.data, d.ctypes.data,
                                       len(d), accel)
        return tex

    def tonemap_tiles(self, scene: Scene, raw_texels: np.ndarray, spa: int, tint_extra: int = 0) -> np.ndarray:
        """main.c:68-79 + saveAs_core for every wall: concatenated RGB bytes in wall order."""
        n = int(sum(int(w["lightmapSetup"][1]) * int(w["lightmapSetup"][2]) for w in scene.walls))
        out = np.zeros(3 * n, dtype=np.uint8)
        t = np.ascontiguousarray(raw_texels, dtype=np.float32)
        self.lib.orc_tonemap_tiles(scene.walls.ctypes.data, len(scene.walls), t.ctypes.data, int(spa), int(tint_extra),
                                   out.ctypes.data)
        return out

    def philox(self, ctr, key):
        c = np.asarray(ctr, dtype=np.uint32)
        k = np.asarray(key, dtype=np.uint32)
        out = np.empty(4, dtype=np.uint32)
        self.lib.orc_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        return out

    def philox2x32(self, ctr, key: int):
        c = np.asarray(ctr, dtype=np.uint32)
        out = np.empty(2, dtype=np.uint32)
        self.lib.orc_philox2x32_10(c.ctypes.data, int(key) & 0xFFFFFFFF, out.ctypes.data)
        return out
