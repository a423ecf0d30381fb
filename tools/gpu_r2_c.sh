set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
tail -15 gpurun_out/r2c_tests.log
