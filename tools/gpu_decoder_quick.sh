#!/bin/bash
# quick decoder session: decoder + slice + depthconv tests, step profile, config-3 bench
mkdir -p gpurun_out
python -m pytest tests/test_decoder_gpu.py tests/test_concat_gpu.py tests/test_guard_bands_gpu.py -k "decoder or concat" -m gpu -q --timeout 600 > gpurun_out/pytest_dec.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_dec.log
python tools/profile_decoder.py > gpurun_out/profile_decoder.json 2> gpurun_out/profile_decoder.err; echo "profile exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/profile_decoder.json'))
for tag,v in d.items():
    print("==", tag, "GPU total %.1f ms" % (v['gpu_us_total']/1e3))
    for r in v['top'][:22]:
        print("   %8.1f us %5.1f%% x%-4d %s" % (r['us'], 100*r['share'], r['calls'], r['name'][:100]))
PY
timeout 600 python tools/bench_decoder.py --config 3 --steps 5 --warmup 3 > gpurun_out/decoder_cfg3.jsonl 2> gpurun_out/decoder_cfg3.err; echo "decoder exit $?"; cut -c1-330 gpurun_out/decoder_cfg3.jsonl
