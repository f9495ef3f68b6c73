#!/bin/bash
# bf16-focused A/B: GPU parity tests of the default build, then bf16 (and f32) bench lines for default + $TAGS, head sweep bf16 for both
mkdir -p gpurun_out/ab
LIBDIR=$PWD/bts-fully-tf_b200/lib
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
for tag in default $TAGS; do
  if [ $tag = default ]; then unset BTSLPG_LIB; else export BTSLPG_LIB=$LIBDIR/libbtslpg_$tag.so; fi
  for i in 1 2; do
    timeout 300 python bench.py --skip-cpu --skip-e2e --skip-decoder --dtype bf16 > gpurun_out/ab/bench_${tag}_bf16_$i.json 2>/dev/null
  done
  python tools/sweep_head.py > gpurun_out/ab/sweep_head_$tag.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab/bench_*bf16_?.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); pp=d['extras']['per_pass']
    print("%-36s %8.1f GB/s  fwd %6.2f us (%.3f)  bwd %6.2f us (%.3f)" % (f.split('/')[-1], d['value'], pp['fwd']['us'], pp['fwd']['frac_of_peak'], pp['bwd']['us'], pp['bwd']['frac_of_peak']))
import os
tags=[os.path.basename(f)[11:-5] for f in sorted(glob.glob('gpurun_out/ab/sweep_head_*.json'))]
pts={t:{(p['dtype'],p['enc'],p['kernel']):p for p in json.load(open('gpurun_out/ab/sweep_head_%s.json'%t))['points']} for t in tags}
for k in pts[tags[0]]:
    if k[0]=='bf16': print("%-5s %-12s %-18s" % k, "  ".join("%s %7.2f us %.3f" % (t, pts[t][k]['us'], pts[t][k]['frac']) for t in tags))
PY
