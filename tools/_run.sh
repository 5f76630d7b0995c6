bash tools/gpu_round.sh v2c quick > gpurun_out/v2c_round.log 2>&1
tail -12 gpurun_out/v2c_round.log | cut -c1-1000; tail -10 gpurun_out/v2c_trace.log
