python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --config c5 --steps 20 2>&1 | tail -1
python bench.py --config c5 --steps 20 --nenv 65536 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-900
