#!/bin/bash
# multi-GPU session: N = number of visible GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_f32_n$N.json 2> gpurun_out/bench_f32_n$N.err; echo "bench n$N exit $?"; cat gpurun_out/bench_f32_n$N.json | head -c 1500; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref n$N exit $?"; cat gpurun_out/bench_ref_n$N.json | head -c 600; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_decoder.py --steps 5 --warmup 2 > gpurun_out/decoder_fused_n$N.jsonl 2> gpurun_out/decoder_fused_n$N.err; echo "decoder n$N exit $?"; cat gpurun_out/decoder_fused_n$N.jsonl
