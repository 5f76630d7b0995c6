#!/usr/bin/env bash
# One GPU-box session: tests, smoke, bench, ncu launch list and full captures.  Logs -> gpurun_out/.
#   tools/gpu_round.sh TAG [quick]
set -u
mkdir -p gpurun_out
TAG="${1:-r01}"
QUICK="${2:-}"
python -m pytest tests -m gpu -q --maxfail=25 -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
python tools/gpu_exp.py 16384 128 0 > gpurun_out/${TAG}_exp.log 2>&1
python tools/gpu_trace.py 16384 128 > gpurun_out/${TAG}_trace.log 2>&1
if [ "$QUICK" == "quick" ]; then
python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
else
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bwd_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_bwd \
    python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_ncu_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stats_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_fwd \
    python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_ncu_fwd.log 2>&1
python tools/gpu_dense_bench.py > gpurun_out/${TAG}_dense_bench.json 2> gpurun_out/${TAG}_dense_bench.err
SPCL_DENSE_TMA=0 python tools/gpu_dense_bench.py > gpurun_out/${TAG}_dense_bench_ldg.json 2>> gpurun_out/${TAG}_dense_bench.err
SPCL_DENSE_TMA=1 python tools/gpu_dense_bench.py > gpurun_out/${TAG}_dense_bench_tma.json 2>> gpurun_out/${TAG}_dense_bench.err
for W in fwd bwd; do
ncu --set full --clock-control none --import-source on -k regex:pool_rows_${W} -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_dense_${W} \
    python tools/gpu_dense_bench.py > gpurun_out/${TAG}_ncu_dense_${W}.log 2>&1
done
python tools/cfg5_step.py > gpurun_out/${TAG}_cfg5.json 2> gpurun_out/${TAG}_cfg5.err
SPCL_DENSE_TMA=1 timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -q > gpurun_out/${TAG}_pytest_tma.log 2>&1; echo "pytest(tma) rc=$?" | tee -a gpurun_out/${TAG}_pytest_tma.log
fi
tail -4 gpurun_out/${TAG}_pytest_tma.log 2>/dev/null; cat gpurun_out/${TAG}_dense_bench_ldg.json gpurun_out/${TAG}_dense_bench_tma.json 2>/dev/null | cut -c1-700
tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_exp.log; cat gpurun_out/${TAG}_bench.json | cut -c1-900
