set -x
python tools/exp_tc.py 2>&1 | tail -12
B2_TC_PROJECT=1 python -m pytest tests -m gpu -x -q -k "c4 or forward_matches or golden" 2>&1 | tail -5
