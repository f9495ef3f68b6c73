python tools/profile_decoder.py > gpurun_out/profile_decoder.json 2> gpurun_out/profile_decoder.err; echo "profile exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/profile_decoder.json'))
v=d['train_352x1216_b16']
print("train GPU total %.1f ms" % (v['gpu_us_total']/1e3))
for r in v['top'][:14]:
    print("   %8.1f us %5.1f%% x%-4d %s" % (r['us'], 100*r['share'], r['calls'], r['name'][:110]))
PY
