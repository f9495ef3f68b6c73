#!/bin/bash
# final multi-GPU measurement: decoder configs and the driver's bench launch at N = $1 ranks (add "tests" as $2 for the NCCL parity tests)
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
if [ "$2" = "tests" ]; then
python -m pytest tests/test_trainer_gpu.py tests/test_optim_gpu.py -q -m gpu --timeout 900 > gpurun_out/pytest_train_n$N.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_train_n$N.log | cut -c1-300
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/bench_decoder.py --config 5 4 3 --steps 5 --warmup 3 > gpurun_out/dec_n$N.jsonl 2> gpurun_out/dec_n$N.err; echo "decoder exit $?"
t0=$SECONDS; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $? wall $((SECONDS-t0)) s"
