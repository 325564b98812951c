python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/exp_graph.py c3 2>&1 | tail -1
B2_NO_ORDER_FORK=1 python tools/exp_graph.py c3 2>&1 | tail -1
python tools/exp_graph.py c5 2>&1 | tail -1
B2_NO_ORDER_FORK=1 python tools/exp_graph.py c5 2>&1 | tail -1
EXP_CAPS=100 python tools/exp_pgs.py c3 2>&1 | tail -1 | cut -c1-250
