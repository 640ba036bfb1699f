set -x
mkdir -p gpurun_out
timeout 600 python tools/diag_band.py 2>&1 | tail -40 > gpurun_out/diag_band.txt
cat gpurun_out/diag_band.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_big_scenes.py::test_c2_benchmark_rows_match_reference 2>&1 | tail -25 > gpurun_out/r2_pytest_2.txt
cat gpurun_out/r2_pytest_2.txt
P="python tools/probe.py bench_data/v3_spaceship.mcrtpack.xz --width 1920 --height 1080 --sqrtspp 3 --modes f64 --reps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 3 -c 1 -o gpurun_out/r2_shade_v3 $P > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 3 -c 1 -o gpurun_out/r2_extend_v3 $P > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
