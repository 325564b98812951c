for v in "" "B2_BENCH_NO_SAMPLER=1" "B2_BENCH_SMI_MS=1000"; do
env $v python bench.py --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'ms', round(d['ms_per_step'],3), 'blocks', [round(x,2) for x in d['block_ms_min_med_max']], {k: round(v,3) for k,v in d['roofline']['kernel_ms_all'].items() if v>0.05}, 'clk', d['clocks'], 'exch', d['obs_exchange']['fused_peer_store_ms_per_step'] if d.get('obs_exchange') else None)
"
done
