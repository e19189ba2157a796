set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_full.py tests/test_gpu_beam.py -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; tail -6 gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --workload dnn60k_mp --steps 2 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_dnn60k.json 2> gpurun_out/bench_dnn60k.err; tail -c 1800 gpurun_out/bench_dnn60k.json; tail -3 gpurun_out/bench_dnn60k.err
JB200_HEAP_CACHE=0 timeout 900 python bench.py --workload dnn60k_mp --steps 1 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_dnn60k_nocache.json 2> gpurun_out/bench_dnn60k_nocache.err; tail -c 600 gpurun_out/bench_dnn60k_nocache.json
