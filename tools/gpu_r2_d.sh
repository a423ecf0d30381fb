set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -k "synth4000" > gpurun_out/r2d_synth4000.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_synth4000.log
grep -E "synth4000 tile|passed|failed|Error|assert" gpurun_out/r2d_synth4000.log | cut -c1-900
python -m pytest tests -m gpu -q > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_tests.log
tail -4 gpurun_out/r2d_tests.log
