#!/bin/bash
# multi-GPU rehearsal: NCCL trainer test, training configs and the driver's bench launch at N = $1 ranks
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
python -m pytest tests/test_trainer_gpu.py tests/test_optim_gpu.py tests/test_wgrad_gpu.py -q -m gpu --timeout 900 > gpurun_out/pytest_train_n$N.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_train_n$N.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/bench_decoder.py --config 5 4 3 --steps 5 --warmup 3 > gpurun_out/dec_n$N.jsonl 2> gpurun_out/dec_n$N.err; echo "decoder exit $?"; cut -c1-1500 gpurun_out/dec_n$N.jsonl; tail -5 gpurun_out/dec_n$N.err | cut -c1-400
t0=$SECONDS; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $? wall $((SECONDS-t0)) s"; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print("value", d['value'], "e2e", d['e2e'])
for k in ('decoder_config3','decoder_config5','decoder_config4'):
    print(k, json.dumps(d['extras'].get(k))[:1200])
PY
