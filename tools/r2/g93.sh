cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/ddp_grad_check.py 2>&1 | tail -4
PWC_WGRAD_STREAM=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/ddp_grad_check.py 2>&1 | tail -2
