# Scratch script of the last gpurun call (development): A/B of library variants built into variants/*.so with
#   SPCL_OUT=$PWD/variants/x.so SPCL_NVCC_EXTRA="-DSOME_SWITCH=1" bash self-paced-contrastive-learning_b200/csrc/build.sh
# and selected per process with SPCL_B200_LIB.  Example:
for V in cur; do
  if [ $V == cur ]; then unset SPCL_B200_LIB; else export SPCL_B200_LIB=$PWD/variants/$V.so; fi
  timeout 300 python tools/gpu_time.py 16384 128 self
done
