set -x
mkdir -p gpurun_out
timeout 300 python tools/refcuda_debug.py dnn20k 100 > gpurun_out/refcuda_debug.txt 2>&1; tail -40 gpurun_out/refcuda_debug.txt
timeout 200 python bench.py --workload tri20k_mp --steps 3 --warmup 2 --no-extra-legs --no-cpu-baseline > gpurun_out/bench_r02d_tri20k_mp.json 2> gpurun_out/bench_r02d_tri20k_mp.err; tail -c 400 gpurun_out/bench_r02d_tri20k_mp.json; tail -2 gpurun_out/bench_r02d_tri20k_mp.err
( time timeout 300 python -m pytest tests/test_gpu_beam.py tests/test_gpu_stream.py tests/test_gpu_gmm.py -m gpu -q ) > gpurun_out/pytest_c.txt 2>&1; tail -6 gpurun_out/pytest_c.txt
( time timeout 200 python -m pytest tests/test_gpu_host.py -m gpu -q -k user_defined ) > gpurun_out/pytest_d.txt 2>&1; tail -12 gpurun_out/pytest_d.txt
