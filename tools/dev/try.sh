python -m pytest tests -m gpu -x -q 2>&1 | tail -2
EXP_CAPS=100 python tools/exp_pgs.py c3 2>&1 | tail -1 | cut -c1-250
EXP_CAPS=100 python tools/exp_pgs.py c5 2>&1 | tail -1 | cut -c1-250
EXP_CAPS=100 python tools/exp_pgs.py c4 2>&1 | tail -1 | cut -c1-250
python bench.py --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; python tools/bench_line.py gpurun_out/bench_c3.json | cut -c1-330
