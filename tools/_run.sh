mkdir -p gpurun_out
rm -f gpurun_out/probe_fast.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5
C2="bench_data/c2_hexagon_room.mcrtpack --sqrtspp 8"
V3="bench_data/v3_spaceship.mcrtpack.xz --width 1920 --height 1080 --sqrtspp 6"
for L in 0 4 2; do
  export MCRT_BVH4_MAX_LEAF=$L
  echo "== max leaf $L"
  timeout 300 python tools/probe_fast.py $C2 --tag "c2_leaf$L" --skip-exact-render --skip-trace 2>&1 | grep -E "render"
  timeout 400 python tools/probe_fast.py $V3 --tag "v3_leaf$L" $( [ $L != 4 ] && echo --skip-exact-render --skip-trace ) 2>&1 | grep -E "render|trace|fast vs"
done
unset MCRT_BVH4_MAX_LEAF
echo "== v3 dynamic_fetch=0"
timeout 400 python tools/probe_fast.py $V3 --tag "v3_nodyn" --opt dynamic_fetch=0 --skip-exact-render --skip-trace 2>&1 | grep -E "render"
echo "== c2 dynamic_fetch=1"
timeout 400 python tools/probe_fast.py $C2 --tag "c2_dyn" --opt dynamic_fetch=1 --skip-exact-render --skip-trace 2>&1 | grep -E "render"
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_c2_a.json 2> gpurun_out/r2_bench_c2_a.err
tail -c 6000 gpurun_out/r2_bench_c2_a.json; tail -5 gpurun_out/r2_bench_c2_a.err
P="python tools/probe.py bench_data/v3_spaceship.mcrtpack.xz --width 1920 --height 1080 --sqrtspp 3 --modes f64 --reps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 3 -c 1 -o gpurun_out/r2_extend_v3_dyn $P > gpurun_out/ncu2.log 2>&1
