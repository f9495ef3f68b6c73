#!/bin/bash
# full GPU suite on the default build, then the tail micro-benchmark for the default and the $TAGS experiment builds
mkdir -p gpurun_out/ab
LIBDIR=$PWD/bts-fully-tf_b200/lib
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
for tag in default $TAGS; do
  if [ $tag = default ]; then unset BTSLPG_LIB; else export BTSLPG_LIB=$LIBDIR/libbtslpg_$tag.so; fi
  python tools/bench_tail.py --skip-cpu --skip-literal --skip-concat > gpurun_out/ab/tail_$tag.json 2>/dev/null
  python tools/bench_tail.py --skip-cpu --skip-literal --skip-concat --height 352 --width 1216 > gpurun_out/ab/tail_kitti_$tag.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab/tail_*.json')):
    d=json.load(open(f))
    print("%-40s" % f.split('/')[-1], " ".join("%s %.2f us (%.3f)" % (k, d[k]['us'], d[k]['frac_of_peak']) for k in ('silog_fwd','silog_bwd','eval_metrics')))
PY
