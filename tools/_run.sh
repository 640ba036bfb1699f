mkdir -p gpurun_out
rm -f gpurun_out/probe_fast.jsonl
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
C2="bench_data/c2_hexagon_room.mcrtpack --sqrtspp 8"
V3="bench_data/v3_spaceship.mcrtpack.xz --width 1920 --height 1080 --sqrtspp 6"
timeout 300 python tools/probe_fast.py $C2 --tag c2_occ 2>&1 | grep -E "render|fast vs"
timeout 400 python tools/probe_fast.py $V3 --tag v3_occ 2>&1 | grep -E "render|fast vs"
for v in _s5 _s6; do
  export MCRT_LIB=$PWD/monte-carlo-ray-tracer_b200/libmcrt_b200$v.so
  echo "== variant $v"
  timeout 300 python tools/probe_fast.py $C2 --tag "c2$v" --skip-exact-render --skip-trace 2>&1 | grep -E "render"
done
