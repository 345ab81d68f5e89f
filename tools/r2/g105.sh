cd /root/repo
for g in 148 0 138 128 112 148 0; do echo "PWC_CV_GRID=$g"; PWC_CV_GRID=$g timeout 120 python tools/roofline_once.py 8 2>&1 | tail -1; done
for g in 148 0; do echo "B=16 PWC_CV_GRID=$g"; PWC_CV_GRID=$g timeout 120 python tools/roofline_once.py 16 2>&1 | tail -1; done
