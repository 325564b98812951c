python -m pytest tests -m gpu -x -q 2>&1 | tail -2
EXP_CAPS=100 python tools/exp_pgs.py c3 2>&1 | tail -1 | cut -c1-250
EXP_CAPS=100 python tools/exp_pgs.py c5 2>&1 | tail -1 | cut -c1-250
EXP_CAPS=100 python tools/exp_pgs.py c4 2>&1 | tail -1 | cut -c1-250
