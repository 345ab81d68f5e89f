cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "cost_volume or cv" 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume_quad -s 10 -c 1 -f -o gpurun_out/r2_cv_quad_tmaout python tools/roofline_once.py 8 > gpurun_out/r2_ncu_cv_tmaout.log 2>&1; tail -2 gpurun_out/r2_ncu_cv_tmaout.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cost_volume_quad -c 60 --csv --log-file gpurun_out/r2_launches_cv_roofline_tmaout.csv python tools/roofline_once.py 8 > /dev/null 2>&1; tail -2 gpurun_out/r2_launches_cv_roofline_tmaout.csv | cut -c1-200
python tools/roofline_once.py 8 2>&1 | tail -2
