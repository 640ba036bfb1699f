timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/all_tests.log 2>&1; tail -3 gpurun_out/all_tests.log
for v in "" _nopf; do
  for rep in 1 2; do
  MCRT_LIB=$PWD/monte-carlo-ray-tracer_b200/libmcrt_b200$v.so python bench.py --sqrtspp 8 --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 f64 variant[$v]', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(x,1) for k,x in d['roofline']['stage_ms_per_step'].items()})"
  done
  MCRT_LIB=$PWD/monte-carlo-ray-tracer_b200/libmcrt_b200$v.so python bench.py --workload c3 --sqrtspp 8 --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3 f64 variant[$v]', round(d['value'],1), {k: round(x,1) for k,x in d['roofline']['stage_ms_per_step'].items()})"
  MCRT_LIB=$PWD/monte-carlo-ray-tracer_b200/libmcrt_b200$v.so python bench.py --sqrtspp 8 --no-cpu-baseline --steps 3 --warmup 3 --precision f32 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 f32 variant[$v]', round(d['value'],1), {k: round(x,1) for k,x in d['roofline']['stage_ms_per_step'].items()})"
done
