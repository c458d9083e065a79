#!/bin/bash
# One-GPU evidence run for profiles/ (round 2): bench lines of every BASELINE config, the reference arm, the C++ shim latency,
# the ncu launch list of the bench command and --set full captures of the dominant kernel (its plain and its shifted launch of one
# 640-image batch) and of the other kernels of a batch.
out=${1:-gpurun_out/r2z}
mkdir -p $out
python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_reference_arm.json 2> $out/ref.err; echo ref rc=$?
python bench.py --steps 20 --warmup 5 > $out/bench_1gpu.json 2> $out/own.err; echo own rc=$?
for c in c1 c3 c5a c5b; do python bench.py --config $c --steps 10 --warmup 3 > $out/bench_1gpu_$c.json 2> $out/$c.err; echo $c rc=$?; done
python tools/shim_latency.py > $out/shim_latency.json 2> $out/shim.err; echo shim rc=$?
python tools/kfdb_bench.py 2>/dev/null > $out/kfdb_latency.jsonl
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency --no-copy-probe"
$CMD > $out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv $CMD > $out/ncu_launch.log 2>&1; echo launches rc=$?
P="python tools/prof_extract.py 320 2"
$P > $out/plain2.log 2>&1 && ncu --set full --metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum --clock-control none --import-source on -k regex:k_fast_strips -s 2 -c 2 -f -o $out/fast_strips $P > $out/ncu_fast.log 2>&1; echo fast rc=$?
$P > $out/plain3.log 2>&1 && ncu --set full --clock-control none -k regex:"k_resize_tiles|k_import|k_fill|k_blur|k_describe|k_octree|k_stereo" -s 16 -c 16 -f -o $out/others $P > $out/ncu_others.log 2>&1; echo others rc=$?
