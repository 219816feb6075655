set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -m gpu -x -q > gpurun_out/r02_gputest_7.log 2>&1; tail -3 gpurun_out/r02_gputest_7.log
python tools/unet_gemm_dbg.py 8 64 > gpurun_out/r02_gemm_dbg_b8_v7.txt 2> gpurun_out/r02_gemm_dbg_b8_v7.err
python bench.py --steps 2 --warmup 3 --profile-out gpurun_out/r02_unet_ops_v7.json > gpurun_out/r02_bench_v7.json 2> gpurun_out/r02_bench_v7.err; cat gpurun_out/r02_bench_v7.json | cut -c1-400
