#!/bin/bash
# Short refresh of the 1-GPU evidence after a kernel change: the bench line, the FAST captures and the stereo kernels' capture.
out=${1:-gpurun_out/r2zz}
mkdir -p $out
python bench.py --steps 20 --warmup 5 > $out/bench_1gpu.json 2> $out/own.err; echo own rc=$?
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency --no-copy-probe"
$CMD > $out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv $CMD > $out/ncu_launch.log 2>&1; echo launches rc=$?
P="python tools/prof_extract.py 320 2"
$P > $out/plain2.log 2>&1 && ncu --set full --metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum --clock-control none --import-source on -k regex:k_fast_strips -s 2 -c 2 -f -o $out/fast_strips $P > $out/ncu_fast.log 2>&1; echo fast rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_stereo" -s 3 -c 3 -f -o $out/stereo $P > $out/ncu_stereo.log 2>&1; echo stereo rc=$?
