python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --config c4 --steps 10 2>&1 | tail -1
