time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_full_n8_v2.json 2> gpurun_out/bench_full_n8_v2.err || { tail -40 gpurun_out/bench_full_n8_v2.err; exit 1; }
python -c "
import json
txt=open('gpurun_out/bench_full_n8_v2.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k in ('epa','self_collision','pipeline'): print(k, d[k]['value'], d[k].get('stage_ms_rank0'))
bp=d['broad_phase']['dense']; print({x:bp[x] for x in bp if 'ms' in x or 'pairs_per_s' in x})
"
