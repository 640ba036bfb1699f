mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_pytest_final.txt; cat gpurun_out/r2_pytest_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --no-profile --no-cpu-baseline 2> gpurun_out/r2_bench_final2.err | grep '^{' > gpurun_out/r2_bench_final2.json; head -c 300 gpurun_out/r2_bench_final2.json; echo; tail -3 gpurun_out/r2_bench_final2.err
