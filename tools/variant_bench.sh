#!/bin/bash
# Runs the short device-resident bench for the default library and every gpurun_variants/lib_*.so (tuning helper).
out=${1:-gpurun_out/variants.txt}
mkdir -p "$(dirname "$out")"
run() {
  python bench.py --steps 10 --warmup 3 --no-hamming --no-cpu-baseline --no-latency --no-copy-probe 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', round(d['value']), 'fps', round(d['ms_per_step'],3), 'ms', 'e2e', round(d['e2e']['value']), {k:round(v,3) for k,v in d['stage_ms'].items()})"
}
run default | tee "$out"
for f in $(ls gpurun_variants/lib_*.so 2>/dev/null); do ORBGPU_LIB=$PWD/$f run "$(basename $f)" | tee -a "$out"; done
