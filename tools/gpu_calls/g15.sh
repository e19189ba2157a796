set -x
mkdir -p gpurun_out
( time JB200_LIB=$PWD/julius_b200/libjb200_mb5.so EXP_MAX_UTTS=740 timeout 500 python tools/exp_pipeline.py tri20k 3 r5_nopipe:740:0,r4_nopipe:592:0,r4_pipe64:592:64,r4_pipe32:592:32,r4_pipe125:592:125 ) > gpurun_out/exp_mb5.txt 2> gpurun_out/exp_mb5.err; cat gpurun_out/exp_mb5.txt; tail -5 gpurun_out/exp_mb5.err
( time JB200_LIB=$PWD/julius_b200/libjb200_mb5.so timeout 300 python -m pytest tests/test_gpu_beam.py tests/test_gpu_stream.py -m gpu -q -x ) > gpurun_out/pytest_mb5.txt 2>&1; tail -5 gpurun_out/pytest_mb5.txt
