mkdir -p gpurun_out
EPI_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/timeline.py 10m 1 > gpurun_out/r2n_tl.log 2>&1
grep -v commit gpurun_out/timeline_r1.txt | head -90
