set -x
time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_full_n8.json 2> gpurun_out/bench_full_n8.err || { tail -40 gpurun_out/bench_full_n8.err; exit 1; }
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_full_n8.json') if l.startswith('{')][-1])
print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()})
for k in ('e2e','broad_phase','epa','self_collision','pipeline'): print(k, json.dumps(d[k])[:2600])
"
