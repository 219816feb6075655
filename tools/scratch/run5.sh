mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_clip.py tests/test_gpu_vae.py -m gpu -x -q > gpurun_out/r02_gputest_8.log 2>&1; tail -3 gpurun_out/r02_gputest_8.log
python tools/unet_gemm_dbg.py 8 64 > gpurun_out/r02_gemm_dbg_b8_v9.txt 2> gpurun_out/r02_gemm_dbg_b8_v9.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02_unet_ops_v9.json > gpurun_out/r02_bench_v9.json 2> gpurun_out/r02_bench_v9.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v9.json')); print(d['value'], d['unet_step_ms'], d['roofline']['frac'], d['kernel_families_ms_per_unet_call'])"
