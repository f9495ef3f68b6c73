#!/bin/bash
# dress rehearsal of the driver's round-end GPU tier: tests, smoke(), reference arm, default bench
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
t0=$SECONDS; python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $? wall $((SECONDS-t0)) s"; cut -c1-300 gpurun_out/bench_ref.json; echo
t0=$SECONDS; python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $? wall $((SECONDS-t0)) s"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print("value", d['value'], d['unit'], "ms/step", d['ms_per_step'], "steps", d['steps'], "warmup", d['warmup'])
print("roofline", d['roofline'])
print("e2e", d['e2e'], "cpu", d['cpu_baseline']['value'], d['cpu_baseline']['cores'], "clocks", d['clocks'], "launches", d['gpu_launches'])
for k, v in d['extras'].items():
    print(k, json.dumps(v)[:1500])
PY
