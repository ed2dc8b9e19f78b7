timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | head
export D3D_SAN_COLLIDERS=200 D3D_SAN_PAIRS=600
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize.py > gpurun_out/r02_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_memcheck.txt
tail -3 gpurun_out/r02_memcheck.txt
export D3D_SAN_COLLIDERS=80 D3D_SAN_PAIRS=150 D3D_GJK_SPLIT_MIN=1
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python scripts/sanitize.py > gpurun_out/r02_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_racecheck.txt
tail -3 gpurun_out/r02_racecheck.txt
