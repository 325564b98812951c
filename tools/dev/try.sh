EXP_CAPS=100,50,30,10 python tools/exp_pgs.py c4 2>&1 | tail -9 | cut -c1-120
