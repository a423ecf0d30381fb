set -u
mkdir -p gpurun_out
for k in ${STEPS:-2}; do
FMGI_ROOM_STEPS=$k ncu --set full --import-source on --clock-control none -k regex:k_trace -s 2 -c 1 -o gpurun_out/prof_rooms_example_k$k -f python bench.py --no-cpu --no-app --no-secondary --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/rooms_ncu_k$k.log 2>&1
done
ls -la gpurun_out/prof_rooms*
