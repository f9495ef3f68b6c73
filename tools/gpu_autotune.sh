timeout 900 python tools/bench_decoder.py --steps 5 --warmup 3 > gpurun_out/decoder_autotune.jsonl 2> gpurun_out/decoder_autotune.err; echo "exit $?"
timeout 900 python tools/bench_decoder.py --steps 5 --warmup 3 --no-cudnn-autotune > gpurun_out/decoder_heur.jsonl 2>/dev/null
python - <<'PY'
import json
for f in ("decoder_autotune","decoder_heur"):
    for l in open('gpurun_out/%s.jsonl'%f):
        x=json.loads(l); print(f, "cfg", x['config'], x['value'], "img/s", x['ms_per_step'], "ms")
PY
