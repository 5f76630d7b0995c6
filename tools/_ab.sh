python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf16 or cfg3 or positive or stats_parts" 2>&1 | tail -2
for V in notry cur tryw notry cur; do
  if [ $V == cur ]; then unset SPCL_B200_LIB; else export SPCL_B200_LIB=$PWD/variants/$V.so; fi
  python tools/gpu_time.py 16384 128 self
done
unset SPCL_B200_LIB; python tools/gpu_time.py 16384 128 slice
