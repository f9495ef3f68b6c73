#!/bin/bash
# tests + smoke + decoder configs + launch list of the default bench + full ncu capture of the tcgen05 iconv1 kernel
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python tools/bench_decoder.py --config 3 5 4 --steps 5 --warmup 3 > gpurun_out/dec_n1.jsonl 2> gpurun_out/dec_n1.err; echo "decoder exit $?"; cut -c1-400 gpurun_out/dec_n1.jsonl
CMD="python bench.py --steps 20 --warmup 3 --skip-decoder --skip-cpu"
$CMD > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
CMD2="python tools/prof_iconv.py 32"
$CMD2 > gpurun_out/iconv_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"iconv1_fwd" -s 2 -c 1 -f -o gpurun_out/prof_iconv $CMD2 > gpurun_out/ncu_iconv.log 2>&1
echo "ncu iconv exit $?"
