python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 2>&1 | tail -5
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2
nproc
