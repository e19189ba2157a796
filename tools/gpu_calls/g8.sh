set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; tail -6 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
( time timeout 1500 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_r02_tri20k.json 2> gpurun_out/bench_r02_tri20k.err; tail -c 1500 gpurun_out/bench_r02_tri20k.json; tail -4 gpurun_out/bench_r02_tri20k.err
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; tail -c 600 gpurun_out/bench_r02_reference.json; tail -4 gpurun_out/bench_r02_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/ncu_launches_r02.csv python bench.py --no-cpu-baseline --no-extra-legs --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beam_kernel -s 1 -c 1 -o gpurun_out/prof_beam_r02 python bench.py --no-cpu-baseline --no-extra-legs --steps 1 --warmup 1 --frames 300 > gpurun_out/ncu_beam.log 2>&1; tail -2 gpurun_out/ncu_beam.log
timeout 900 ncu --set full --clock-control none -k regex:gmm_score -s 1 -c 1 -o gpurun_out/prof_gmm_r02 python bench.py --no-cpu-baseline --no-extra-legs --steps 1 --warmup 1 --frames 300 > gpurun_out/ncu_gmm.log 2>&1; tail -2 gpurun_out/ncu_gmm.log
for w in tri20k_mp dnn20k; do timeout 900 python bench.py --workload $w --steps 3 --warmup 2 --no-extra-legs > gpurun_out/bench_r02_$w.json 2> gpurun_out/bench_r02_$w.err; tail -c 400 gpurun_out/bench_r02_$w.json; done
timeout 900 python bench.py --workload tri20k_gbeam --utts 512 --steps 3 --warmup 2 --no-extra-legs --no-cpu-baseline > gpurun_out/bench_r02_tri20k_gbeam.json 2> gpurun_out/bench_r02_tri20k_gbeam.err; tail -c 400 gpurun_out/bench_r02_tri20k_gbeam.json
ls -la gpurun_out
