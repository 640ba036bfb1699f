mkdir -p gpurun_out
rm -f gpurun_out/probe_fast.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_big_scenes.py -m gpu -q 2>&1 | tail -4
C2="bench_data/c2_hexagon_room.mcrtpack --sqrtspp 8"
V3="bench_data/v3_spaceship.mcrtpack.xz --width 1920 --height 1080 --sqrtspp 6"
V5="bench_data/v5_lego_bulldozer.mcrtpack.xz --width 1920 --height 1080 --sqrtspp 4"
for v in "" _ns _ls; do
  export MCRT_LIB=$PWD/monte-carlo-ray-tracer_b200/libmcrt_b200$v.so
  echo "== variant '$v'"
  timeout 300 python tools/probe_fast.py $C2 --tag "c2$v" --skip-exact-render --skip-trace 2>&1 | grep -E "render"
  timeout 400 python tools/probe_fast.py $V3 --tag "v3$v" --skip-exact-render --skip-trace 2>&1 | grep -E "render"
  timeout 400 python tools/probe_fast.py $V5 --tag "v5$v" --skip-exact-render --skip-trace 2>&1 | grep -E "render"
done
