set -x
mkdir -p gpurun_out
python tools/unet_gemm_dbg.py 8 64 > gpurun_out/r02_gemm_dbg_b8.txt 2> gpurun_out/r02_gemm_dbg_b8.err
python tools/unet_gemm_dbg.py 1 64 > gpurun_out/r02_gemm_dbg_b1.txt 2> gpurun_out/r02_gemm_dbg_b1.err
python tools/ncu_target.py 1 64 > gpurun_out/ncu_target_b1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b1_v6.csv python tools/ncu_target.py 1 64 > gpurun_out/ncu_b1.log 2>&1
tail -3 gpurun_out/ncu_b1.log
