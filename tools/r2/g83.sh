cd /root/repo
timeout 300 python tools/halo_narrow_bench.py 2>&1 | tail -8
