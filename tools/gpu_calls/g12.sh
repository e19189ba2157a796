set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/pytest_gpu.txt 2>&1; tail -30 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
( time timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-legs ) > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 2500 gpurun_out/bench_quick.json; tail -4 gpurun_out/bench_quick.err
python - <<'PY' > gpurun_out/refcuda.txt 2>&1
import sys, json; sys.path.insert(0,'.')
import bench
try:
    print(json.dumps(bench.reference_cuda_dnn("dnn20k", 300), indent=1))
except Exception as e:
    print("FAILED", e)
PY
tail -20 gpurun_out/refcuda.txt
