mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gputests.txt
cat gpurun_out/gputests.txt
timeout 600 python bench.py > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err; tail -c 600 gpurun_out/bench_b1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['ms_per_step'], d.get('cpu_baseline'), d['roofline']['ms_per_call'], d['roofline']['frac'])
print(json.dumps(d.get('breakdown'), indent=None)[:3000])
PY
