set -u
mkdir -p gpurun_out
g++ -O2 -std=c++17 -ffp-contract=off -I include tests/cpu/rooms_check.cpp flatmatch-global-illumination_b200/csrc/rooms_build.cpp -o /tmp/rooms_check -lpthread
python - <<'PY'
import sys, numpy as np
sys.path.insert(0,'oracle'); sys.path.insert(0,'flatmatch-global-illumination_b200')
import refbind
sc = refbind.Scene.load('tests/golden/synth4000_scene.npz')
with open('/tmp/scene.bin','wb') as f:
    np.array([len(sc.walls),len(sc.windows),len(sc.lights)],dtype='<i4').tofile(f)
    for t in (sc.walls,sc.windows,sc.lights): np.ascontiguousarray(t).tofile(f)
PY
nproc
for t in 1 2 4 8 16; do echo threads $t; FMGI_BUILD_THREADS=$t FMGI_ROOMS_TIMING=1 /tmp/rooms_check /tmp/scene.bin 1000 2>&1 | grep "^\[rooms\|^rooms"; done
FMGI_ROOMS_TIMING=1 python bench.py --no-cpu --no-app --no-secondary --steps 3 --warmup 2 --e2e-steps 3 --workload synth4000_1e9x4 2>gpurun_out/rb.err | tail -1 > gpurun_out/rb_synth.json
grep rooms gpurun_out/rb.err | tail -3
python - <<'PY'
import json
d = json.loads(open("gpurun_out/rb_synth.json").read())
print("value %.4g kernel_ms %.3f e2e %.4g" % (d["value"], d["kernel_ms_per_step"], d["e2e"]["value"]), d["e2e"])
PY
