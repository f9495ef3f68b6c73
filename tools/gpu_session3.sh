bash tools/gpu_head.sh
timeout 900 python tools/bench_decoder.py --steps 5 --warmup 2 > gpurun_out/decoder_fused_n1.jsonl 2> gpurun_out/decoder_fused_n1.err; echo "decoder exit $?"; cut -c1-420 gpurun_out/decoder_fused_n1.jsonl
