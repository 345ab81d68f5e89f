cd /root/repo
timeout 60 python tools/halo_narrow_dbg.py 16 16 16 224 512 2>&1 | tail -14
timeout 60 python tools/halo_narrow_dbg.py 32 32 16 112 256 2>&1 | tail -14
