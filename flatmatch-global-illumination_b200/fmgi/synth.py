"""Synthetic floor-plan layouts in the reference's colour code (parseLayout.c:15-24).

BASELINE.json configs[2] asks for a "synthetic 4000x4000 px generated multi-room layout (~20k
rectangles)".  This module draws one: a seeded grid of rectangular rooms with 4-px walls, door gaps
between neighbours, green window runs on the rooms that touch the outside, and windowless
interior rooms (which the reference's parser lights with auto-placed ceiling lights,
parseLayout.c:298-341).  The output is the pixel array the reference's loadImage() would produce
from a PNG (image.c:189-217): uint32 words 0xAABBGGRR.  Turning pixels into rectangles stays the
job of the reference's own, untouched parseLayout().
"""
from __future__ import annotations

import numpy as np

WALL = 0xFF000000
EMPTY = 0xFFFFFFFF
OUTSIDE = 0xFF7F7F7F
DOOR = 0xFFDFDFDF
WINDOW = 0xFF00FF00


def make_layout(size_px: int = 4000, seed: int = 1, room_min_m: float = 3.0, room_max_m: float = 6.0,
                pixels_per_metre: float = 30.0, wall_px: int = 4, margin_px: int = 8,
                door_px: int = 27, window_frac: float = 0.5) -> np.ndarray:
    """Returns a (size_px, size_px) uint32 layout.  Deterministic for a given argument tuple."""
    rng = np.random.default_rng(seed)
    img = np.full((size_px, size_px), OUTSIDE, dtype=np.uint32)
    lo, hi = margin_px, size_px - margin_px

    def cuts():
        """Wall centre lines along one axis: room widths uniform in [room_min, room_max] metres."""
        pos, out = lo, [lo]
        while True:
            step = int(rng.uniform(room_min_m, room_max_m) * pixels_per_metre)
            if pos + step + int(room_min_m * pixels_per_metre) > hi - wall_px:
                out.append(hi - wall_px)
                return out
            pos += step
            out.append(pos)

    xs, ys = cuts(), cuts()
    # everything inside the outer wall starts as wall, rooms are carved out
    img[lo:hi, lo:hi] = WALL
    nx, ny = len(xs) - 1, len(ys) - 1
    for j in range(ny):
        for i in range(nx):
            img[ys[j] + wall_px: ys[j + 1], xs[i] + wall_px: xs[i + 1]] = EMPTY

    def door(a0, a1):
        """Position of a door gap inside the open span [a0, a1)."""
        span = a1 - a0
        w = min(door_px, max(span - 6, 1))
        s = a0 + int(rng.integers(3, max(span - w - 2, 4)))
        return s, min(s + w, a1 - 1)

    # doors between horizontally / vertically adjacent rooms
    for j in range(ny):
        for i in range(nx):
            if i + 1 < nx and rng.random() < 0.7:
                s, e = door(ys[j] + wall_px, ys[j + 1])
                img[s:e, xs[i + 1]: xs[i + 1] + wall_px] = DOOR
            if j + 1 < ny and rng.random() < 0.7:
                s, e = door(xs[i] + wall_px, xs[i + 1])
                img[ys[j + 1]: ys[j + 1] + wall_px, s:e] = DOOR

    # windows in the outer wall of perimeter rooms: a green run replacing the wall pixels, so that
    # it has OUTSIDE on one side and EMPTY on the other (parseLayout.c:102-113 creates the emitter
    # on the OUTSIDE<->WINDOW edge)
    def window(a0, a1):
        span = a1 - a0
        w = max(int(span * window_frac), 4)
        s = a0 + (span - w) // 2
        return s, s + w

    for i in range(nx):
        s, e = window(xs[i] + wall_px, xs[i + 1])
        if rng.random() < 0.8:
            img[ys[0]: ys[0] + wall_px, s:e] = WINDOW
        if rng.random() < 0.8:
            img[ys[ny]: ys[ny] + wall_px, s:e] = WINDOW
    for j in range(ny):
        s, e = window(ys[j] + wall_px, ys[j + 1])
        if rng.random() < 0.8:
            img[s:e, xs[0]: xs[0] + wall_px] = WINDOW
        if rng.random() < 0.8:
            img[s:e, xs[nx]: xs[nx] + wall_px] = WINDOW
    return img


def to_rgb(img: np.ndarray) -> np.ndarray:
    """(H, W, 3) uint8 view for writing the layout as a PNG."""
    out = np.empty(img.shape + (3,), dtype=np.uint8)
    out[..., 0] = img & 0xFF
    out[..., 1] = (img >> 8) & 0xFF
    out[..., 2] = (img >> 16) & 0xFF
    return out
