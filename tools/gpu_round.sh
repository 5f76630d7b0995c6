#!/usr/bin/env bash
# One GPU-box session: tests, smoke, bench, ncu launch list and full captures.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
TAG="${1:-r01}"
python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bwd_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_bwd \
    python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_ncu_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 6 -c 2 -f -o gpurun_out/${TAG}_prof_fwd \
    python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_ncu_fwd.log 2>&1
fi
tail -3 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_bench.json | cut -c1-600
