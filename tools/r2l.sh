mkdir -p gpurun_out
EPI_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/timeline.py 10m 2 > gpurun_out/r2l.log 2>&1
grep "^r0" gpurun_out/r2l.log | head -64
