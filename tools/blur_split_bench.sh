mkdir -p gpurun_out/r2r
run() { python bench.py --steps 10 --warmup 3 --no-hamming --no-cpu-baseline --no-latency --no-copy-probe 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', round(d['value']), 'fps', round(d['ms_per_step'],3), 'ms', 'e2e', round(d['e2e']['value']), 'sum of stages', round(sum(d['stage_ms'].values()),3))"; }
(timeout 200 python -m pytest tests/test_gpu_extract.py tests/test_gpu_stereo.py -x -q -m gpu 2>&1 | tail -3) > gpurun_out/r2r/parity.txt
run split | tee gpurun_out/r2r/blur_split.txt
ORBGPU_NO_BLUR_SPLIT=1 run nosplit | tee -a gpurun_out/r2r/blur_split.txt
run split | tee -a gpurun_out/r2r/blur_split.txt
ORBGPU_NO_OVERLAP=1 run no_overlap_at_all | tee -a gpurun_out/r2r/blur_split.txt
cat gpurun_out/r2r/parity.txt
