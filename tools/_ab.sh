SPCL_FUSED_STAMP=1 SPCL_B200_LIB=$PWD/variants/stamp.so timeout 300 python - <<'PY' 2>&1 | grep "spcl fused" | tail -4
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import torch, gpu_small
print(gpu_small.raw_group_times(256, 256))
torch.cuda.synchronize()
PY
timeout 300 python tools/gpu_small_profile.py 2>&1 | grep -v "^$" | cut -c1-150 | sed -n 1,12p
timeout 300 python tools/gpu_small.py 2>&1 | grep "cfg"
