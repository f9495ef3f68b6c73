for pb in 64 8; do
timeout 600 python tools/bench_decoder.py --config 3 --steps 10 --warmup 3 --per-gpu-batch $pb > gpurun_out/dec_graph_$pb.jsonl 2> gpurun_out/dec_graph_$pb.err; echo "graph b$pb exit $?"; tail -2 gpurun_out/dec_graph_$pb.err
timeout 600 python tools/bench_decoder.py --config 3 --steps 10 --warmup 3 --per-gpu-batch $pb --no-graph > gpurun_out/dec_eager_$pb.jsonl 2>/dev/null
done
timeout 600 python bench.py --skip-cpu --skip-e2e --steps 50 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ("dec_graph_64","dec_eager_64","dec_graph_8","dec_eager_8"):
    for l in open('gpurun_out/%s.jsonl'%f):
        x=json.loads(l); print(f, x['value'], "img/s", x['ms_per_step'], "ms graph=", x.get('cuda_graph'))
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1]); print(d['extras'].get('decoder_config3'))
PY
