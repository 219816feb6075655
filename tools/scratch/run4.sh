mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err
PBE_GEMM_PAIR_MIN_K=5 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02_unet_ops_pair5.json > gpurun_out/ab_pair5.json 2> gpurun_out/ab_pair5.err
PBE_GEMM_PAIR_MIN_K=5 python tools/unet_gemm_dbg.py 8 64 > gpurun_out/r02_gemm_dbg_b8_pair5.txt 2> gpurun_out/r02_gemm_dbg_b8_pair5.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ab_default2.json 2> gpurun_out/ab_default2.err
for f in ab_default ab_pair5 ab_default2; do python -c "
import json,sys; d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['unet_step_ms']['p50'], d['kernel_families_ms_per_unet_call'])"; done
