set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_fused_gather.py 2>&1 | tail -25 | tee gpurun_out/r02_fused_gather_n2.txt
