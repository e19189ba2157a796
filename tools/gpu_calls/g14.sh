set -x
mkdir -p gpurun_out
( time timeout 420 python tools/exp_pipeline.py tri20k 3 base,host_numbering,b3_nopipe,b3_pipe64,b3_pipe125 ) > gpurun_out/exp_pipeline2.txt 2> gpurun_out/exp_pipeline2.err; cat gpurun_out/exp_pipeline2.txt; tail -5 gpurun_out/exp_pipeline2.err
( time timeout 500 python -m pytest tests/test_gpu_stream.py tests/test_gpu_beam.py -m gpu -q --durations=8 ) > gpurun_out/pytest_a.txt 2>&1; tail -25 gpurun_out/pytest_a.txt
( time timeout 300 python -m pytest tests/test_gpu_host.py -m gpu -q --durations=8 -k "frame_by_frame or progressive or linked_in" ) > gpurun_out/pytest_b.txt 2>&1; tail -25 gpurun_out/pytest_b.txt
