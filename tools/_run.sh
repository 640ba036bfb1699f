mkdir -p gpurun_out
echo "== adapter test, pinned seed + 6 random seeds (photon mapper)"
EXE=monte-carlo-ray-tracer_b200/host/_build/test_gpu_integrators
MCRT_SEED=0x12345678 timeout 200 $EXE oracle/_ref/scenes hexagon_room.json 1 24 16 2 20000 2>&1 | grep -E "OK|MISMATCH|error"
MCRT_SEED=0x12345678 timeout 200 $EXE oracle/_ref/scenes hexagon_room.json 0 24 16 2 2>&1 | grep -E "OK|MISMATCH|error"
for sd in 1 2 3 4 5 6; do MCRT_SEED=$sd timeout 200 $EXE oracle/_ref/scenes hexagon_room.json 1 24 16 2 20000 2>&1 | grep -E "Mapper" ; done
C2="bench_data/c2_hexagon_room.mcrtpack --sqrtspp 8 --modes f64 --reps 2"
V3="bench_data/v3_spaceship.mcrtpack.xz --width 1920 --height 1080 --sqrtspp 6 --modes f64 --reps 2"
for pool in 16777216 33554432; do for bps in 4 8 16; do
  echo "== pool $pool bps $bps"
  timeout 200 python tools/probe.py $C2 --pool $pool --bps $bps 2>&1 | grep rep1 | cut -c1-110
  timeout 200 python tools/probe.py $V3 --pool $pool --bps $bps 2>&1 | grep rep1 | cut -c1-110
done; done
