#!/usr/bin/env bash
# Quick GPU session: full GPU test suite + smoke + op timings.  tools/gpu_quick.sh TAG
set -u
TAG="${1:-r02}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
python tools/gpu_time.py 16384 128 self > gpurun_out/${TAG}_time.log 2>&1
python tools/gpu_time.py 16384 128 slice >> gpurun_out/${TAG}_time.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_time.log
