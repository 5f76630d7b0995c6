timeout 120 tools/mma_bench.bin 2>&1 | tee gpurun_out/r02zb_mma_bench.txt
SPCL_B200_LIB=$PWD/variants/trace.so timeout 300 python tools/gpu_trace.py 16384 128 0 both 2>&1 | sed -n "/bwd_kernel/,\$p" > gpurun_out/r02zb_trace_wide.log; sed -n 1,3p gpurun_out/r02zb_trace_wide.log; sed -n 20,30p gpurun_out/r02zb_trace_wide.log
