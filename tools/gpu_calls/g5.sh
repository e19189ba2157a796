set -x
mkdir -p gpurun_out
( cd tools/ubench && ./heapx ) 2>&1 | grep variant | tail -4
timeout 900 python -m pytest tests/test_gpu_beam.py tests/test_gpu_full.py -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; tail -5 gpurun_out/pytest_gpu.txt
for v in "" nocandb owner; do
  lib=julius_b200/libjb200${v:+_$v}.so
  JB200_LIB=$PWD/$lib timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_v_${v:-base}.json 2> gpurun_out/bench_v_${v:-base}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v_${v:-base}.json').read().strip().splitlines()[-1]); r=d['roofline']
print('${v:-base}', round(d['value']), r['kernel_ms'], r['beam_phase_cycles_per_frame'], d['decoded_ok'])
PY
done
JB200_HEAP_SINGLE=3 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_v_warp4.json 2> gpurun_out/bench_v_warp4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v_warp4.json').read().strip().splitlines()[-1]); r=d['roofline']
print('warp4', round(d['value']), r['kernel_ms'], r['beam_phase_cycles_per_frame'], d['decoded_ok'])
PY
