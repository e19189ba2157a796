set -x
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/shim_leg.json 2>&1
import sys, json; sys.path.insert(0,'.')
import bench
print(json.dumps(bench.shim_leg("tri20k", 64, 1000, 32), indent=1))
PY
cat gpurun_out/shim_leg.json
for mode in "" JB200_NO_L2_WINDOW=1; do
  env $mode timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_l2_${mode:-window}.json 2> gpurun_out/bench_l2.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_l2_${mode:-window}.json').read().strip().splitlines()[-1]); r=d['roofline']
print('${mode:-window}', round(d['value']), r['kernel_ms'], r['beam_phase_cycles_per_frame'], d['decoded_ok'])
PY
done
timeout 900 python bench.py --workload dnn60k_mp --steps 2 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_dnn60k.json 2> gpurun_out/bench_dnn60k.err; tail -c 1500 gpurun_out/bench_dnn60k.json; tail -3 gpurun_out/bench_dnn60k.err
timeout 600 python bench.py --workload tri20k_mp --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_mp.json 2> gpurun_out/bench_mp.err; tail -c 1300 gpurun_out/bench_mp.json
