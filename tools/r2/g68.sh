cd /root/repo
./tools/bin/store_bw 2>&1 | tail -12
