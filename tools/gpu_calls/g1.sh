set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( cd tools/ubench && ./heapx ) > gpurun_out/heapx.txt 2>&1; tail -12 gpurun_out/heapx.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -5 gpurun_out/pytest_gpu.txt
JB200_ENABLE_GRAMMAR=1 timeout 300 python -m pytest tests/test_gpu_beam.py -m gpu -k grammar -q > gpurun_out/pytest_grammar.txt 2>&1; tail -15 gpurun_out/pytest_grammar.txt
JB200_GPU_EXTRA_CASES=1 timeout 300 python -m pytest tests/test_gpu_beam.py -m gpu -k cpu_pinned -q > gpurun_out/pytest_extra.txt 2>&1; tail -15 gpurun_out/pytest_extra.txt
JB200_DNN_KERNEL=2 timeout 120 python -m pytest tests/test_gpu_dnn.py -m gpu -q > gpurun_out/pytest_dnn_cluster.txt 2>&1; tail -8 gpurun_out/pytest_dnn_cluster.txt
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err; tail -c 1500 gpurun_out/bench_pipe.json
JB200_HEAP_SINGLE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; tail -c 600 gpurun_out/bench_single.json
JB200_HEAP_SINGLE=2 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_pipe_generic.json 2> gpurun_out/bench_pipe_generic.err; tail -c 600 gpurun_out/bench_pipe_generic.json
timeout 600 python bench.py --workload tri20k_mp --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_mp.json 2> gpurun_out/bench_mp.err; tail -c 600 gpurun_out/bench_mp.json
