cd /root/repo
timeout 300 python tools/dual_stream_once.py 8 300 2>&1 | tail -8
