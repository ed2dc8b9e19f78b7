timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | head
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
export D3D_SAN_COLLIDERS=200 D3D_SAN_PAIRS=600
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize.py > gpurun_out/r02_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_memcheck.txt
tail -3 gpurun_out/r02_memcheck.txt
export D3D_SAN_COLLIDERS=80 D3D_SAN_PAIRS=150 D3D_GJK_SPLIT_MIN=1
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python scripts/sanitize.py > gpurun_out/r02_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_racecheck.txt
tail -3 gpurun_out/r02_racecheck.txt
unset D3D_SAN_COLLIDERS D3D_SAN_PAIRS D3D_GJK_SPLIT_MIN
time timeout 900 python bench.py > gpurun_out/bench_full_n1_v6.json 2> gpurun_out/bench_full_n1_v6.err || tail -30 gpurun_out/bench_full_n1_v6.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full_n1_v6.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['value_including_host_packing'], d['gpu_launches'], d['clocks'])
for k in ('epa','self_collision','pipeline','six_type_mix','libccd','hydroelastic'):
    if k in d: print(k, d[k]['value'], d[k].get('parity_on_cpu_sample'), d[k].get('stage_ms_rank0'))
print(d['parity_on_cpu_sample'], d['cpu_baseline'])
"
