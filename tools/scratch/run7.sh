mkdir -p gpurun_out
for v in v5 v7 cur v5 cur; do echo "== $v"; PBE_B200_LIB=tools/bin/libpbe_$v.so python tools/gemm_probe.py geglu proj_out nores ffout 2>&1 | grep -v "^\[pbe\]"; done > gpurun_out/r02_probe_ab.txt 2>&1
cat gpurun_out/r02_probe_ab.txt
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02_unet_ops_v11.json > gpurun_out/r02_bench_v11.json 2> gpurun_out/r02_bench_v11.err
PBE_LN_FOLD=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02_unet_ops_v11_nofold.json > gpurun_out/r02_bench_v11_nofold.json 2> gpurun_out/r02_bench_v11_nofold.err
PBE_B200_LIB=tools/bin/libpbe_v7.so python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02_unet_ops_v11_v7.json > gpurun_out/r02_bench_v11_v7.json 2> gpurun_out/r02_bench_v11_v7.err
for f in r02_bench_v11 r02_bench_v11_nofold r02_bench_v11_v7; do python -c "
import json; d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['unet_step_ms']['p50'], d['roofline']['frac'], d['kernel_families_ms_per_unet_call'])"; done
