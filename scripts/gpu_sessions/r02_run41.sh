python -m pytest tests/test_norm_gpu.py tests/test_gjk_gpu.py tests/test_epa_gpu.py -x -q 2>&1 | grep -E "^E|passed|failed" | head
python scripts/gjk_c1.py 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2_v3.json 2> gpurun_out/bench_n2_v3.err || tail -20 gpurun_out/bench_n2_v3.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_v3.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in ('epa','self_collision','pipeline'): print(k, d[k]['value'], d[k].get('stage_ms_rank0'))
bp=d['broad_phase']['dense']; print({x:bp[x] for x in bp if 'ms' in x or 'pairs_per_s' in x})
"
