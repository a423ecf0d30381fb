set -u
APP=$PWD/flatmatch-global-illumination_b200/build/globalIllumination
PNG=$PWD/tests/golden/example.png
mkdir -p /tmp/t1/tiles && cd /tmp/t1
for i in 1 2 3; do
  s=$(date +%s.%N)
  FMGI_STATS=2 CUDA_VISIBLE_DEVICES=0 $APP $PNG > out.txt 2>&1; rc=$?
  e=$(date +%s.%N)
  echo "run $i rc=$rc wall $(echo "$e - $s" | bc) s"
  grep -E "breakdown|exit_begin" out.txt
done
ls tiles | wc -l
