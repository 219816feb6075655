mkdir -p gpurun_out
python tools/unet_gemm_dbg.py 8 64 > gpurun_out/r02_gemm_dbg_b8_v8.txt 2> gpurun_out/r02_gemm_dbg_b8_v8.err
