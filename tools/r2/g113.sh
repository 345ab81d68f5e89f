cd /root/repo
PWC_TC_KSPLIT=1 PWC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv3x3_tc_f16 -c 15 --csv --log-file gpurun_out/r2_launches_f16_ksplit.csv python tools/fwd_once.py > /dev/null 2>&1
PWC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv3x3_tc_f16 -c 15 --csv --log-file gpurun_out/r2_launches_f16_noksplit.csv python tools/fwd_once.py > /dev/null 2>&1
python - <<'PY'
import csv
for n in ("ksplit","noksplit"):
    lines=[l for l in open(f'gpurun_out/r2_launches_f16_{n}.csv') if l.startswith('"')]
    rows=list(csv.DictReader(lines))
    print(n, [(r['Grid Size'], round(float(r['Metric Value'].replace(',',''))/ (1000 if r['Metric Unit']=='ns' else 1),1)) for r in rows[10:15]])
PY
for v in 0 1 0 1; do
  if [ $v = 1 ]; then unset PWC_TC_KSPLIT; else export PWC_TC_KSPLIT=1; fi
  timeout 300 python bench.py --no-train --no-cpu-baseline --min-seconds 1.0 2>/dev/null | tail -1 | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print('NO_KSPLIT=', os.environ.get('PWC_TC_NO_KSPLIT'), 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'burst', round(d['burst_value'],1), d['probe']['sha256_16'], d['clocks']['sm_mhz'])"
done
