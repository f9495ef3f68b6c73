#!/bin/bash
mkdir -p gpurun_out/ab
X=$PWD/bts-fully-tf_b200/lib/libbtslpg_x.so
BTSLPG_LIB=$X python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/ab/pytest_x.log 2>&1; echo "pytest x exit $?"; tail -3 gpurun_out/ab/pytest_x.log
for i in 1 2; do
timeout 300 python bench.py --skip-cpu --skip-e2e > gpurun_out/ab/bench_v5_$i.json 2> gpurun_out/ab/bench_v5_$i.err; echo "v5 exit $?"
BTSLPG_LIB=$X timeout 300 python bench.py --skip-cpu --skip-e2e > gpurun_out/ab/bench_x_$i.json 2> gpurun_out/ab/bench_x_$i.err; echo "x exit $?"
done
timeout 300 python bench.py --skip-cpu --skip-e2e --dtype bf16 > gpurun_out/ab/bench_v5_bf16.json 2>/dev/null
BTSLPG_LIB=$X timeout 300 python bench.py --skip-cpu --skip-e2e --dtype bf16 > gpurun_out/ab/bench_x_bf16.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['extras']['per_pass'])
    except Exception as e: print(f, 'ERR', e)
PY
