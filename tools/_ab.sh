timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_distributed_gpu.py -m gpu -q -x -k "bf16 or cfg3 or positive or stats_parts or optional or hostfeed or raw or graph" 2>&1 | tail -3
SPCL_B200_LIB=$PWD/variants/trace.so timeout 300 python tools/gpu_trace_sp.py self 2>&1 | tail -8 | tee gpurun_out/r02zj_trace_sp_self.txt
SPCL_B200_LIB=$PWD/variants/trace.so timeout 300 python tools/gpu_trace_sp.py slice 2>&1 | tail -16 | tee gpurun_out/r02zj_trace_sp_slice.txt
timeout 300 python tools/gpu_fwd_parts.py 2>&1 | tee gpurun_out/r02zj_parts.log
