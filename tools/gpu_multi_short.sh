#!/bin/bash
# multi-GPU session without the reference arm: N = number of visible GPUs
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_f32_n$N.json 2> gpurun_out/bench_f32_n$N.err; echo "bench n$N exit $?"; cut -c1-400 gpurun_out/bench_f32_n$N.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_decoder.py --steps 5 --warmup 2 > gpurun_out/decoder_fused_n$N.jsonl 2> gpurun_out/decoder_fused_n$N.err; echo "decoder n$N exit $?"; cut -c1-300 gpurun_out/decoder_fused_n$N.jsonl
