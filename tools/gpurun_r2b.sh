set -x
python tools/exp_pgs.py c3 2>&1 | tail -20
python tools/exp_pgs.py c4 2>&1 | tail -20
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
