set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; tail -15 gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --workload dnn60k_mp --steps 2 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_dnn60k.json 2> gpurun_out/bench_dnn60k.err; tail -c 3000 gpurun_out/bench_dnn60k.json; tail -3 gpurun_out/bench_dnn60k.err
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_tri20k_b.json 2> gpurun_out/bench_tri20k_b.err; tail -c 1800 gpurun_out/bench_tri20k_b.json
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs --utts 444 > gpurun_out/bench_tri20k_444.json 2> gpurun_out/bench_tri20k_444.err; tail -c 1500 gpurun_out/bench_tri20k_444.json
timeout 600 python bench.py --workload tri20k_gbeam --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs --utts 512 > gpurun_out/bench_gbeam512.json 2> gpurun_out/bench_gbeam512.err; tail -c 1500 gpurun_out/bench_gbeam512.json
