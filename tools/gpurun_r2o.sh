set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 50 --warmup 5 --no-other-configs > gpurun_out/bench_4gpu_b.json 2> /dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_4gpu_b.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "n", d["n_gpus"])
print("obs_exchange", d.get("obs_exchange"))
PY
