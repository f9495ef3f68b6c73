#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -8 gpurun_out/pytest_gpu.log
python tools/profile_tail_convs.py > gpurun_out/tail_convs.json 2> gpurun_out/tail_convs.err; echo "tail convs exit $?"
timeout 900 python tools/bench_decoder.py --steps 5 --warmup 3 > gpurun_out/decoder_fused_n1.jsonl 2> gpurun_out/decoder_fused_n1.err; echo "decoder exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/tail_convs.json'))
for k,v in d.items():
    if 'ours_fused_backward' in v: print(k, "cuDNN bwd kernels:", [t for t,n in v['kernels_us'][:4]], "ours:", v['ours_fused_backward'])
for l in open('gpurun_out/decoder_fused_n1.jsonl'):
    x=json.loads(l); print("cfg", x['config'], x['value'], "img/s", x['ms_per_step'], "ms")
PY
