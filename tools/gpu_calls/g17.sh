set -x
mkdir -p gpurun_out
( time timeout 400 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_r02c_tri20k.json 2> gpurun_out/bench_r02c_tri20k.err; tail -c 600 gpurun_out/bench_r02c_tri20k.json; tail -4 gpurun_out/bench_r02c_tri20k.err
( time timeout 300 python bench.py --workload dnn20k --steps 3 --warmup 2 --no-extra-legs ) > gpurun_out/bench_r02c_dnn20k.json 2> gpurun_out/bench_r02c_dnn20k.err; tail -c 900 gpurun_out/bench_r02c_dnn20k.json; tail -3 gpurun_out/bench_r02c_dnn20k.err
for w in tri20k_mp tri20k_gbeam; do timeout 200 python bench.py --workload $w --steps 3 --warmup 2 --no-extra-legs --no-cpu-baseline > gpurun_out/bench_r02c_$w.json 2> gpurun_out/bench_r02c_$w.err; tail -c 300 gpurun_out/bench_r02c_$w.json; tail -2 gpurun_out/bench_r02c_$w.err; done
for f in test_gpu_gmm test_gpu_dnn test_gpu_sweep test_gpu_full test_gpu_host; do ( time timeout 420 python -m pytest tests/$f.py -m gpu -q --durations=5 ) > gpurun_out/pytest_$f.txt 2>&1; tail -12 gpurun_out/pytest_$f.txt; done
