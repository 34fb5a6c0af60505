#!/bin/bash
# A/B of the warp-specialised front end (YL_FILTER=ws) against the split form; run on the GPU box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export YL_FILTER=ws
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "ws_default: $(timeout 120 python tools/run_stage.py filter --iters 20)"
TAG=ws_default timeout 120 python tools/step_probe.py
for v in yolov4_b200/variants/lib_*.so; do
  echo "$(basename $v): $(YL_LIB=$PWD/$v timeout 120 python tools/run_stage.py filter --iters 20)"
  TAG=$(basename $v) YL_LIB=$PWD/$v timeout 120 python tools/step_probe.py
done
unset YL_FILTER
YL_FILTER=ws N=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 40 --csv --log-file gpurun_out/ws_launches.csv python tools/step_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/ws_launches.csv')))
for r in rows[-7:]:
    if len(r)>14: print(r[4][:50], r[-1])
PY
YL_FILTER=ws N=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_filter_raw_ws -s 2 -c 1 -o gpurun_out/ws_full python tools/step_probe.py > gpurun_out/ws_ncu.log 2>&1
tail -2 gpurun_out/ws_ncu.log
timeout 600 python -m pytest tests/test_exchange_multigpu.py -m gpu -x -q 2>&1 | tail -30
