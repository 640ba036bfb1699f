mkdir -p gpurun_out
C2="bench_data/c2_hexagon_room.mcrtpack --sqrtspp 8 --modes f64 --reps 2 --pool 33554432"
V3="bench_data/v3_spaceship.mcrtpack.xz --width 1920 --height 1080 --sqrtspp 6 --modes f64 --reps 2 --pool 33554432"
for g in default 32 64 128; do
  if [ $g = default ]; then unset MCRT_L2_FETCH; else export MCRT_L2_FETCH=$g; fi
  echo "== L2 fetch granularity $g"
  timeout 200 python tools/probe.py $C2 2>&1 | grep -E "rep1|granularity" | cut -c1-330
  timeout 200 python tools/probe.py $V3 2>&1 | grep -E "rep1" | cut -c1-330
done
export MCRT_L2_FETCH=32
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_extend|k_shadow|k_shade" -s 9 -c 6 python bench.py --child-render --workload c2 --sqrtspp 4 2>&1 | grep -E "k_extend|k_shadow|k_shade<|dram__bytes|gpu__time" | cut -c1-150
