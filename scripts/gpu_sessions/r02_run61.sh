timeout 300 python scripts/gjk_c1.py 2>&1 | tail -2
D3D_N=1048576 timeout 300 python scripts/gjk_per_type.py sphere,ellipsoid,capsule,cylinder,box 2>&1 | tail -2
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
