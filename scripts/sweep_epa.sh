#!/bin/bash
for so in scripts/lib_epa*.so; do
  echo "== $so"
  D3D_B200_LIB=$PWD/$so python bench.py --workload epa --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
