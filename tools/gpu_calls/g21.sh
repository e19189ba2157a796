set -x
mkdir -p gpurun_out
( time JB200_CHECK_HEAP=1 timeout 300 python -m pytest tests/test_gpu_beam.py tests/test_gpu_stream.py -m gpu -q ) > gpurun_out/pytest_g.txt 2>&1; tail -6 gpurun_out/pytest_g.txt
( time timeout 300 python -m pytest tests/test_gpu_full.py -m gpu -q ) > gpurun_out/pytest_h.txt 2>&1; tail -6 gpurun_out/pytest_h.txt
( time timeout 300 python tools/exp_pipeline.py tri20k 3 base,no_relocate,b3_pipe32 ) > gpurun_out/exp_reloc2.txt 2> gpurun_out/exp_reloc2.err; cat gpurun_out/exp_reloc2.txt; tail -3 gpurun_out/exp_reloc2.err
( time timeout 400 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_r02e_tri20k.json 2> gpurun_out/bench_r02e_tri20k.err; tail -c 700 gpurun_out/bench_r02e_tri20k.json; tail -4 gpurun_out/bench_r02e_tri20k.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/ncu_launches_r02e.csv python bench.py --no-cpu-baseline --no-extra-legs --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:beam_kernel -s 1 -c 1 -o gpurun_out/prof_beam_r02e python bench.py --no-cpu-baseline --no-extra-legs --steps 1 --warmup 1 --frames 300 --pipe-frames 0 --utts 592 > gpurun_out/ncu_beam.log 2>&1; tail -2 gpurun_out/ncu_beam.log
( time timeout 300 python bench.py --workload dnn20k --steps 3 --warmup 2 --no-extra-legs ) > gpurun_out/bench_r02e_dnn20k.json 2> gpurun_out/bench_r02e_dnn20k.err; tail -c 1200 gpurun_out/bench_r02e_dnn20k.json; tail -3 gpurun_out/bench_r02e_dnn20k.err
