python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 --skip-cpu-baseline | cut -c1-1000
