cat > /tmp/raw.py <<'PY'
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import torch, gpu_small
print(gpu_small.raw_group_times(256, 256))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_group -s 20 -c 1 -f -o gpurun_out/r02zd_prof_fused python /tmp/raw.py > gpurun_out/r02zd_ncu_fused.log 2>&1
ls -la gpurun_out/r02zd_prof_fused.ncu-rep
