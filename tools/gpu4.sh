export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 256 3 > gpurun_out/dist256.log 2>&1; echo "rc $?" >> gpurun_out/dist256.log
tail -25 gpurun_out/dist256.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py 1024 2 > gpurun_out/dist1024.log 2>&1; echo "rc $?" >> gpurun_out/dist1024.log
tail -12 gpurun_out/dist1024.log
