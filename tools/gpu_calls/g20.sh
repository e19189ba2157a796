set -x
mkdir -p gpurun_out
( time JB200_CHECK_HEAP=1 timeout 300 python -m pytest tests/test_gpu_beam.py tests/test_gpu_stream.py -m gpu -q -x ) > gpurun_out/pytest_e.txt 2>&1; tail -8 gpurun_out/pytest_e.txt
( time timeout 300 python -m pytest tests/test_gpu_full.py tests/test_gpu_sweep.py -m gpu -q ) > gpurun_out/pytest_f.txt 2>&1; tail -8 gpurun_out/pytest_f.txt
( time timeout 300 python tools/exp_pipeline.py tri20k 3 base,no_relocate,b3_pipe32 ) > gpurun_out/exp_reloc.txt 2> gpurun_out/exp_reloc.err; cat gpurun_out/exp_reloc.txt; tail -3 gpurun_out/exp_reloc.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
timeout 300 python tools/refcuda_debug.py dnn20k 100 > gpurun_out/refcuda_debug.txt 2>&1; tail -6 gpurun_out/refcuda_debug.txt
