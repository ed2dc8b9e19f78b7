set -x
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_epa<" -s 1 -c 1 -o gpurun_out/r02_epa_c5 python scripts/r02_dev.py pipe > gpurun_out/ncu_epa_c5.log 2>&1
tail -3 gpurun_out/ncu_epa_c5.log
python bench.py --no-extra --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['clocks'])"
