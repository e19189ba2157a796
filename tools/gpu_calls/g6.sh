set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host.py tests/test_gpu_beam.py -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; tail -8 gpurun_out/pytest_gpu.txt
for v in "" p5spec pf pf_p5; do
  lib=julius_b200/libjb200${v:+_$v}.so
  JB200_LIB=$PWD/$lib timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_v_${v:-base}.json 2> gpurun_out/bench_v_${v:-base}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v_${v:-base}.json').read().strip().splitlines()[-1]); r=d['roofline']
print('${v:-base}', round(d['value']), r['kernel_ms'], r['beam_phase_cycles_per_frame'], d['decoded_ok'])
PY
done
( time timeout 1200 python bench.py --steps 3 --warmup 2 ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 2500 gpurun_out/bench_default.json; tail -4 gpurun_out/bench_default.err
