timeout 300 python tools/gpu_small_profile.py grouped 2>&1 | grep -v "^$" | cut -c1-160 | sed -n 1,40p
echo ==== staged
SPCL_FUSED_SMALL=0 timeout 300 python tools/gpu_small_profile.py grouped 2>&1 | grep -v "^$" | cut -c1-160 | sed -n 1,30p
