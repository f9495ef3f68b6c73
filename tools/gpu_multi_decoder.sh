#!/bin/bash
# decoder configs only at N = $1 ranks (the cheap refresh of the scaling table)
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/bench_decoder.py --config 5 4 3 --steps 5 --warmup 3 > gpurun_out/dec_n$N.jsonl 2> gpurun_out/dec_n$N.err; echo "decoder exit $?"
