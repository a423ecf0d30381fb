"""Host-side lightmap layout helpers mirroring the reference's scene builder.

parseLayout refuses atlases above 1e9 bytes (parseLayout.c:519-524) and TILE_SIZE is hard-coded
in main.c:44, so the high-resolution configuration of BASELINE.json (4x texel density, ~2 GB
atlas) needs a harness-built Geometry: `retile` re-derives every wall's tile grid and atlas
offset for another texel density with the reference's own policy.
"""
from __future__ import annotations

import numpy as np


def tile_counts(width_len: np.float32, height_len: np.float32, tile_size: float):
    """createRectangleV's policy (rectangle.c:24-42), float32 arithmetic: start at 1x1, double the
    axis with the lower texels/m until tiles / area >= TILE_SIZE."""
    tw, th = 1, 1
    w, h = np.float32(width_len), np.float32(height_len)
    area = np.float32(w * h)
    ts = np.float32(tile_size)
    cur = np.float32(np.float32(tw * th) / area)
    while cur < ts:
        if np.float32(np.float32(tw) / w) < np.float32(np.float32(th) / h):
            tw *= 2
        else:
            th *= 2
        cur = np.float32(np.float32(tw * th) / area)
    return tw, th


def mipmap_texels(tw: int, th: int) -> int:
    """getNumMipmapTexels (rectangle.c:166-192): the full chain down to 1x1."""
    n = tw * th
    while tw > 1 or th > 1:
        if tw > 1:
            tw //= 2
        if th > 1:
            th //= 2
        n += tw * th
    return n


def _length(v) -> np.float32:
    v = v.astype(np.float32)
    return np.sqrt(np.float32(np.float32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]), dtype=np.float32)


def retile(walls: np.ndarray, tile_size: float):
    """Copy of the wall table with lightmapSetup recomputed for `tile_size` texels per m^2
    (tile counts per rectangle.c:24-42, bases per parseLayout.c:512-517).  Returns (walls, numTexels)."""
    out = walls.copy()
    base = 0
    for i in range(len(out)):
        tw, th = tile_counts(_length(out["width"][i][:3]), _length(out["height"][i][:3]), tile_size)
        out["lightmapSetup"][i] = (base, tw, th, 0)
        base += mipmap_texels(tw, th)
    return out, base
