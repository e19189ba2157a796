set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; tail -15 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
( time timeout 1200 python bench.py --steps 3 --warmup 2 ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 4000 gpurun_out/bench_default.json; tail -4 gpurun_out/bench_default.err
JB200_NO_CLOSED_FORM=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_noclosed.json 2> gpurun_out/bench_noclosed.err; tail -c 1200 gpurun_out/bench_noclosed.json
timeout 600 python bench.py --workload tri20k_mp --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_mp.json 2> gpurun_out/bench_mp.err; tail -c 1500 gpurun_out/bench_mp.json
timeout 600 python bench.py --workload dnn20k --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_dnn20k.json 2> gpurun_out/bench_dnn20k.err; tail -c 2500 gpurun_out/bench_dnn20k.json; tail -3 gpurun_out/bench_dnn20k.err
JB200_DNN_KERNEL=2 timeout 600 python bench.py --workload dnn20k --steps 3 --warmup 2 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_dnn20k_k2.json 2> gpurun_out/bench_dnn20k_k2.err; tail -c 1500 gpurun_out/bench_dnn20k_k2.json; tail -3 gpurun_out/bench_dnn20k_k2.err
