#!/bin/bash
# GPU session: all GPU tests, decoder profile, decoder benches
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -6 gpurun_out/pytest_gpu.log
python tools/profile_decoder.py > gpurun_out/profile_decoder.json 2> gpurun_out/profile_decoder.err; echo "profile exit $?"
timeout 900 python tools/bench_decoder.py --steps 5 --warmup 2 > gpurun_out/decoder_fused_n1.jsonl 2> gpurun_out/decoder_fused_n1.err; echo "decoder exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/profile_decoder.json'))
for tag,v in d.items():
    print("==", tag, "GPU total %.1f ms" % (v['gpu_us_total']/1e3), v.get('tail_ms'))
    for r in v['top'][:12]:
        print("   %8.1f us %5.1f%% x%-4d %s" % (r['us'], 100*r['share'], r['calls'], r['name'][:80]))
for l in open('gpurun_out/decoder_fused_n1.jsonl'):
    x=json.loads(l); print("cfg", x['config'], x['value'], "img/s", x['ms_per_step'], "ms")
PY
