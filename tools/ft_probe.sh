#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export YL_FLAG=tma
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "flag_tma filter: $(timeout 120 python tools/run_stage.py filter --iters 20)"
TAG=flag_tma timeout 120 python tools/step_probe.py
echo "--- groups, YL_FLAG=tma"
timeout 600 python tools/groups_probe.py
unset YL_FLAG
echo "--- groups, YL_FLAG=ldg"
GROUPS=1,2,4 timeout 600 python tools/groups_probe.py
YL_FLAG=tma N=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 40 --csv --log-file gpurun_out/ft_launches.csv python tools/step_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/ft_launches.csv')))
for r in rows[-6:]:
    if len(r)>14: print(r[4][:50], r[-1])
PY
