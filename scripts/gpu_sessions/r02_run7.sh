set -x
for v in "" scripts/lib_gen4.so; do D3D_B200_LIB=$v D3D_SKIP_PAIRS=1 python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from distance3d_b200 import gjk, random as R
n = 1 << 21
for names, hv in ((R.PRIMITIVES + ("mesh",), (10, 10)), (("mesh", "box", "capsule"), (8, 40))):
    dc = R.random_collider_set_device(3, 2 * n, names=names, hull_vertices=hv)
    pairs = torch.arange(2 * n, dtype=torch.int32, device="cuda").reshape(n, 2)
    out = gjk.gjk_distance_batch(dc, pairs)
    for _ in range(2): gjk.gjk_distance_batch(dc, pairs, out=out)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gjk.gjk_distance_batch(dc, pairs, out=out)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get("D3D_B200_LIB") or "default", names, "%.4e pairs/s" % (n / (e0.elapsed_time(e1) / 5) * 1e3))
PY
done 2>&1 | grep pairs/s | tee gpurun_out/r02_gjk_generic4.txt
ncu --set full --clock-control none --import-source on -k regex:k_gjk_thread -c 1 -o gpurun_out/r02_gjk_v8 python scripts/gjk_c1.py 1048576 > gpurun_out/ncu_gjk.log 2>&1
tail -3 gpurun_out/ncu_gjk.log
