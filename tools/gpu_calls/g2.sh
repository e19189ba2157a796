set -x
mkdir -p gpurun_out
( cd tools/ubench && ./heapx ) > gpurun_out/heapx.txt 2>&1; tail -8 gpurun_out/heapx.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; tail -15 gpurun_out/pytest_gpu.txt
( time timeout 900 python bench.py --steps 3 --warmup 2 ) > gpurun_out/bench_tri20k.json 2> gpurun_out/bench_tri20k.err; tail -c 2500 gpurun_out/bench_tri20k.json; tail -4 gpurun_out/bench_tri20k.err
timeout 600 python bench.py --workload tri20k_mp --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_mp.json 2> gpurun_out/bench_mp.err; tail -c 900 gpurun_out/bench_mp.json
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -4 gpurun_out/bench_ref.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beam_kernel -s 1 -c 1 -o gpurun_out/prof_beam python bench.py --no-cpu-baseline --steps 1 --warmup 1 --frames 300 > gpurun_out/ncu_beam.log 2>&1; tail -3 gpurun_out/ncu_beam.log; ls -la gpurun_out/
