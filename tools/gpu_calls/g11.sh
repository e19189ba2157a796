set -x
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/refcuda.txt 2>&1
import sys, json; sys.path.insert(0,'.')
import bench
try:
    print(json.dumps(bench.reference_cuda_dnn("dnn20k", 300), indent=1))
except Exception as e:
    print("FAILED", e)
PY
cat gpurun_out/refcuda.txt | tail -20
for v in "" hints; do
  lib=julius_b200/libjb200${v:+_$v}.so
  JB200_LIB=$PWD/$lib timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_v_${v:-base}.json 2> gpurun_out/bench_v_${v:-base}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v_${v:-base}.json').read().strip().splitlines()[-1]); r=d['roofline']
print('${v:-base}', round(d['value']), r['kernel_ms'], r['beam_phase_cycles_per_frame'], d['decoded_ok'])
PY
done
