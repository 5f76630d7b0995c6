#!/usr/bin/env bash
# One GPU-box session: tests, smoke, bench (+ reference arm), ncu launch list and full captures.  Logs -> gpurun_out/.
#   tools/gpu_round.sh TAG [quick]
set -u
mkdir -p gpurun_out
TAG="${1:-r02}"
QUICK="${2:-}"
python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
if [ "$QUICK" == "quick" ]; then
python bench.py --steps 20 --warmup 5 --quick > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
else
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bwd_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_bwd \
    python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${TAG}_ncu_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stats_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_fwd \
    python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${TAG}_ncu_fwd.log 2>&1
python tools/gpu_dense_bench.py > gpurun_out/${TAG}_dense_bench.json 2> gpurun_out/${TAG}_dense_bench.err
python tools/gpu_aux_bench.py > gpurun_out/${TAG}_aux_bench.json 2> gpurun_out/${TAG}_aux_bench.err
SPCL_AUX_ONCE=1 ncu --set full --clock-control none -k regex:"l2norm|prepare|raw_bwd" -c 40 -f -o gpurun_out/${TAG}_prof_aux \
    python tools/gpu_aux_bench.py > gpurun_out/${TAG}_ncu_aux.log 2>&1
for W in fwd bwd; do
ncu --set full --clock-control none --import-source on -k regex:pool_rows_${W} -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_dense_${W} \
    python tools/gpu_dense_bench.py > gpurun_out/${TAG}_ncu_dense_${W}.log 2>&1
done
python tools/cfg5_step.py > gpurun_out/${TAG}_cfg5.json 2> gpurun_out/${TAG}_cfg5.err
python tools/decoder_step.py > gpurun_out/${TAG}_decoder.json 2> gpurun_out/${TAG}_decoder.err
python tools/gpu_fwd_parts.py > gpurun_out/${TAG}_parts.log 2>&1
# summarise on the box (text travels back; the reports of one session exceed gpurun's 64 MiB return limit)
SPCL_PROFILES_OUT=gpurun_out/profiles_${TAG} python tools/summarize_profiles.py ${TAG} > gpurun_out/${TAG}_summarize.log 2>&1
rm -f gpurun_out/${TAG}_prof_aux.ncu-rep gpurun_out/${TAG}_prof_dense_fwd.ncu-rep gpurun_out/${TAG}_prof_dense_bwd.ncu-rep
du -sh gpurun_out
fi
tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench.json | cut -c1-700
