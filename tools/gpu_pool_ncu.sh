set -u
mkdir -p gpurun_out
for k in 2 4; do
FMGI_POOL_K=$k ncu --set full --import-source on --clock-control none -k regex:k_trace_pool -s 1 -c 1 -o gpurun_out/prof_pool_k$k -f python bench.py --no-cpu --no-app --no-secondary --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/pool_ncu_k$k.log 2>&1
done
ls -la gpurun_out/prof_pool*
