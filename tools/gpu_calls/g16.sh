set -x
mkdir -p gpurun_out
( time timeout 300 python tools/exp_pipeline.py tri20k 3 p32:444:32,p48:444:48,p64:444:64,p96:444:96 ) > gpurun_out/exp_slices.txt 2> gpurun_out/exp_slices.err; cat gpurun_out/exp_slices.txt; tail -3 gpurun_out/exp_slices.err
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_r02b_tri20k.json 2> gpurun_out/bench_r02b_tri20k.err; tail -c 1800 gpurun_out/bench_r02b_tri20k.json; tail -4 gpurun_out/bench_r02b_tri20k.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/ncu_launches_r02b.csv python bench.py --no-cpu-baseline --no-extra-legs --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:beam_kernel -s 1 -c 1 -o gpurun_out/prof_beam_r02b python bench.py --no-cpu-baseline --no-extra-legs --steps 1 --warmup 1 --frames 300 --pipe-frames 0 --utts 592 > gpurun_out/ncu_beam.log 2>&1; tail -2 gpurun_out/ncu_beam.log
ls -la gpurun_out
