time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err || tail -20 gpurun_out/bench_n4.err
python -c "
import json
txt=open('gpurun_out/bench_n4.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['value_including_host_packing'])
for k in ('epa','self_collision','pipeline','six_type_mix'): print(k, d[k]['value'], d[k].get('stage_ms_rank0'))
bp=d['broad_phase']['dense']; print({x:bp[x] for x in bp if x.endswith('_ms') or 'pairs_per_s' in x}, bp.get('fused_gather',{}).get('query_and_gather_ms'))
"
