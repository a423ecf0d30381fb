#!/usr/bin/env python3
"""Per-section instruction / lane / stall-sample breakdown of a k_trace capture taken with
`ncu --set full --import-source on` (build has -lineinfo).

    python profiles/sections_ncu.py gpurun_out/prof.ncu-rep [rays in the captured launch]

Sections are found by marker comments / function heads in the CUDA sources, so the table follows the code.
"""
import csv
import io
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "flatmatch-global-illumination_b200" / "csrc"

MARKERS = {
    "trace_kernels.cuh": [
        (r"^// ---- closest hit against the shared-memory soup", "soup scan"),
        (r"^// ---- closest hit through the floor-plan grid", "grid misc / helpers"),
        (r"void planes\(", "planes"),
        (r"void walk\(", "walk setup"),
        (r"int2? ?range = |ldg256\(p\.grid_table \+ 2 \* ci", "walk loop"),
        (r"int finish\(", "finish"),
        (r"int closest_hit_grid\(", "closest-hit call"),
        (r"int closest_hit_soup_planes\(", "soup + planes"),
        (r"int rooms_locate\(", "rooms locate"),
        (r"int rooms_start\(", "rooms start"),
        (r"float rooms_inv\(", "rooms walk"),
        (r"unsigned rooms_grid_lookup\(", "rooms face grid"),
        (r"while \(\(code >> kRoomKindShift\) == 0u\) \{", "rooms face grid"),
        (r"all lanes whose walk ended, together", "rooms step outcome"),
        (r"^// The whole walk of one ray", "rooms whole walk"),
        (r"int tile_index\(", "tile index"),
        (r"void sample_hemisphere\(", "sampler"),
        (r"void deposit\(", "deposit"),
    ],
    "trace_core.cuh": [
        (r"find_emitter\(", "find emitter"),
        (r"stage_soup\(", "stage soup"),
        (r"__global__ void .*k_trace|k_trace\(const TraceParams", "init"),
        (r"---- A\. refill", "A refill"),
        (r"---- P\. one Philox", "P philox call"),
        (r"---- S\. new direction", "S direction / emission"),
        (r"---- C\. closest hit", "C closest-hit call"),
        (r"---- D\. bounce", "D bounce"),
        (r"photonmap\.c:251", "deposit"),
        (r"---- counters", "counters"),
    ],
}


def section_table(fname):
    path = CSRC / fname
    if not path.exists() or fname not in MARKERS:
        return None
    bounds = []
    lines = path.read_text().splitlines()
    for pat, name in MARKERS[fname]:
        for i, l in enumerate(lines, 1):
            if re.search(pat, l):
                bounds.append((i, name))
                break
    bounds.sort()
    return bounds


def main():
    rep = sys.argv[1]
    rays = float(sys.argv[2]) if len(sys.argv) > 2 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    f = lambda s: int(s) if s.lstrip("-").isdigit() else 0
    cur, hdr, agg, tables = None, None, {}, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            tables.setdefault(cur, section_table(cur))
            continue
        if len(r) > 2 and r[0] == "Line No":
            hdr = r
            continue
        if not hdr or len(r) < 10 or not r[0].isdigit():
            continue
        ln = int(r[0])
        d = dict(zip(hdr[4:], r[4:]))
        name = cur
        if tables.get(cur):
            name = cur + ": (head)"
            for b, n in tables[cur]:
                if ln >= b:
                    name = n
        a = agg.setdefault(name, [0, 0, 0, 0])
        a[0] += f(d["Instructions Executed"]); a[1] += f(d["Thread Instructions Executed"])
        a[2] += f(d["# Samples"]); a[3] += f(d.get("stall_long_sb", "0"))
    tot = sum(a[0] for a in agg.values()); tots = max(1, sum(a[2] for a in agg.values()))
    print(f"{'section':26s} {'warp inst':>9s} {'lanes':>6s} {'samples':>8s} {'long_sb':>8s}" + ("  warp-inst/ray" if rays else ""))
    for c, a in sorted(agg.items(), key=lambda x: -x[1][0]):
        if a[0] * 500 < tot:
            continue
        extra = f"  {a[0] / rays:8.2f}" if rays else ""
        print(f"{c:26s} {100 * a[0] / tot:8.1f}% {a[1] / max(a[0], 1):6.1f} {100 * a[2] / tots:7.1f}% {100 * a[3] / tots:7.1f}%{extra}")
    print(f"total warp instructions {tot}" + (f" = {tot / rays:.1f} per ray" if rays else ""))


if __name__ == "__main__":
    main()
