cd /root/repo
timeout 600 compute-sanitizer --tool memcheck python tools/halo_narrow_dbg.py 32 32 16 96 256 2>&1 | grep -v "^   \|halo dbg" | head -40
