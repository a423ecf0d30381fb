# The memcheck stand-in (compute-sanitizer is closed on the pool): the GPU suite and the small all-kernels exercise
# against the bounds-checked build of the library.  Any index outside its table raises FmgiError in the Python mirror.
set -u
mkdir -p gpurun_out
export FMGI_LIB=$PWD/flatmatch-global-illumination_b200/lib/libfmgi_cuda_checked.so
python -m pytest tests -m gpu -q > gpurun_out/r2_checked_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_checked_tests.log
tail -5 gpurun_out/r2_checked_tests.log
python profiles/sanitize_small.py > gpurun_out/r2_checked_small.log 2>&1; echo "rc=$?" >> gpurun_out/r2_checked_small.log
tail -4 gpurun_out/r2_checked_small.log
FMGI_POOL_K=4 python profiles/sanitize_small.py > gpurun_out/r2_checked_small_pool.log 2>&1; echo "rc=$?" >> gpurun_out/r2_checked_small_pool.log
tail -2 gpurun_out/r2_checked_small_pool.log
