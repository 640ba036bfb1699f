timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/all_tests.log 2>&1; tail -3 gpurun_out/all_tests.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('FINAL C2 f64:', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],3), d['roofline']['stage_ms_per_step'], d['cpu_baseline']['value'], d['clocks'])"
python bench.py --precision f32 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_final_f32.json; python -c "
import json
d=json.loads(open('gpurun_out/bench_final_f32.json').read().strip().splitlines()[-1])
print('FINAL C2 f32:', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
