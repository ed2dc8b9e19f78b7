time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2_v4.json 2> gpurun_out/bench_n2_v4.err || tail -20 gpurun_out/bench_n2_v4.err
head -c 100 gpurun_out/bench_n2_v4.json; echo
python -c "
import json
txt=open('gpurun_out/bench_n2_v4.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['value_including_host_packing'])
for k in ('epa','self_collision','pipeline','six_type_mix'): print(k, d[k]['value'], d[k].get('stage_ms_rank0'))
"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-160
