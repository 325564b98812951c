set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
EXP_CAPS=100 python tools/exp_pgs.py c4 2>&1 | tail -6
