mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_clip.py tests/test_gpu_vae.py -m gpu -x -q > gpurun_out/r02_gputest_9.log 2>&1; tail -15 gpurun_out/r02_gputest_9.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02_unet_ops_v10.json > gpurun_out/r02_bench_v10.json 2> gpurun_out/r02_bench_v10.err
PBE_LN_FOLD=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02_unet_ops_v10_nofold.json > gpurun_out/r02_bench_v10_nofold.json 2> gpurun_out/r02_bench_v10_nofold.err
for f in r02_bench_v10 r02_bench_v10_nofold; do python -c "
import json; d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['unet_step_ms'], d['roofline']['frac'], d['kernel_families_ms_per_unet_call'])"; done
tail -3 gpurun_out/r02_bench_v10.err
