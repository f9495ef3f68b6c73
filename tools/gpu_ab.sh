#!/bin/bash
# A/B session: parity tests on the default build, then the LPG microbench (f32 + bf16) for the default
# build and every experiment build lib/libbtslpg_<tag>.so named in $TAGS.
mkdir -p gpurun_out/ab
LIBDIR=$PWD/bts-fully-tf_b200/lib
python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/ab/pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/ab/pytest.log
for tag in default $TAGS; do
  if [ $tag = default ]; then unset BTSLPG_LIB; else export BTSLPG_LIB=$LIBDIR/libbtslpg_$tag.so; fi
  timeout 300 python bench.py --skip-cpu --skip-e2e --skip-decoder > gpurun_out/ab/bench_${tag}_f32.json 2> gpurun_out/ab/bench_${tag}_f32.err; echo "$tag f32 exit $?"
  timeout 300 python bench.py --skip-cpu --skip-e2e --skip-decoder --dtype bf16 > gpurun_out/ab/bench_${tag}_bf16.json 2> gpurun_out/ab/bench_${tag}_bf16.err; echo "$tag bf16 exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pp=d['extras']['per_pass']
        print("%-44s %8.1f GB/s  fwd %6.2f us (%.3f)  bwd %6.2f us (%.3f)" % (f.split('/')[-1], d['value'], pp['fwd']['us'], pp['fwd']['frac_of_peak'], pp['bwd']['us'], pp['bwd']['frac_of_peak']))
    except Exception as e: print(f, 'ERR', e)
PY
