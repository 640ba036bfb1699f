timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/all_tests.log 2>&1; tail -3 gpurun_out/all_tests.log
python bench.py > gpurun_out/bench_lite.json 2> gpurun_out/bench_lite.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_lite.json').read().strip().splitlines()[-1])
print('C2 f64:', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],3), d['roofline']['stage_ms_per_step'])"
python bench.py --precision f32 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 f32', round(d['value'],1), d['roofline']['stage_ms_per_step'])"
for prec in f64 f32; do python bench.py --workload c3 --sqrtspp 8 --no-cpu-baseline --steps 2 --warmup 3 --precision $prec 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3 $prec', round(d['value'],1), round(d['roofline']['frac'],3), d['roofline']['stage_ms_per_step'])"; done
timeout 300 python tools/probe_emit.py bench_data/c4_water_caustics_small_maps.mcrtpack.xz --emissions 1e6 --sqrtspp 4 --modes f64,f32 > gpurun_out/emit_probe.log 2>&1; tail -6 gpurun_out/emit_probe.log
