python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/gpu_exp.py 16384 128 0
python bench.py --steps 30 --warmup 5 --skip-cpu-baseline | cut -c1-1200
