#!/bin/bash
# A/B of the warp-specialised front end (YL_FILTER=ws) against the split form; run on the GPU box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export YL_FILTER=ws
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
echo "ws_default: $(timeout 120 python tools/run_stage.py filter --iters 20)"
for v in yolov4_b200/variants/lib_*.so; do
  echo "$(basename $v): $(YL_LIB=$PWD/$v timeout 120 python tools/run_stage.py filter --iters 20)"
done
unset YL_FILTER
echo "split: $(timeout 120 python tools/run_stage.py filter --iters 20)"
