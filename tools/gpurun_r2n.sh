set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
tail -c 400 gpurun_out/bench_4gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_4gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "n", d["n_gpus"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
print("obs_exchange", d.get("obs_exchange"))
for c, v in d.get("configs", {}).items(): print(c, v["value"], v["ms_per_step"])
PY
