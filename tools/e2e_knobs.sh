#!/bin/bash
# e2e leg of the short bench under other pipeline depths / chunk counts (tuning helper)
out=${1:-gpurun_out/e2e_knobs.txt}
run() {
  python bench.py --steps 10 --warmup 3 --no-hamming --no-cpu-baseline --no-latency --no-copy-probe 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', 'device', round(d['value']), 'e2e', round(d['e2e']['value']), 'single_call', round(d['e2e']['single_call_value']))"
}
BENCH_PIPELINE_DEPTH=2 run depth2 | tee -a "$out"
BENCH_PIPELINE_DEPTH=3 run depth3 | tee -a "$out"
BENCH_PIPELINE_DEPTH=4 run depth4 | tee -a "$out"
BENCH_PIPELINE_DEPTH=3 ORBGPU_CHUNKS=1 run depth3_chunks1 | tee -a "$out"
