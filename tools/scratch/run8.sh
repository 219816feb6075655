mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_clip.py tests/test_gpu_vae.py -m gpu -x -q > gpurun_out/r02_gputest_10.log 2>&1; tail -15 gpurun_out/r02_gputest_10.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02_unet_ops_v12.json > gpurun_out/r02_bench_v12.json 2> gpurun_out/r02_bench_v12.err
PBE_LN_FOLD=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02_unet_ops_v12_nofold.json > gpurun_out/r02_bench_v12_nofold.json 2> gpurun_out/r02_bench_v12_nofold.err
PBE_B200_LIB=tools/bin/libpbe_v7.so python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02_unet_ops_v12_v7.json > gpurun_out/r02_bench_v12_v7.json 2> gpurun_out/r02_bench_v12_v7.err
for f in r02_bench_v12 r02_bench_v12_nofold r02_bench_v12_v7; do python -c "
import json; d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['unet_step_ms']['p50'], d['roofline']['frac'], d['kernel_families_ms_per_unet_call'])"; done
