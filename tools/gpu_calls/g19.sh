set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
timeout 300 python tools/refcuda_debug.py dnn20k 100 > gpurun_out/refcuda_debug.txt 2>&1; tail -8 gpurun_out/refcuda_debug.txt
( time timeout 400 python bench.py --workload dnn20k --steps 3 --warmup 2 --no-extra-legs ) > gpurun_out/bench_r02d_dnn20k.json 2> gpurun_out/bench_r02d_dnn20k.err; tail -c 1500 gpurun_out/bench_r02d_dnn20k.json; tail -3 gpurun_out/bench_r02d_dnn20k.err
