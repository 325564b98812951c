set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-other-configs > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 600 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "n", d["n_gpus"])
print("obs_exchange", d.get("obs_exchange"))
PY
