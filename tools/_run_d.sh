mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "decode or postprocess or fused" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
timeout 600 python tools/fuzz_parity.py 40 9 | tail -1 >> gpurun_out/d_pytest.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_decode -c 60 --csv --log-file gpurun_out/d_launches.csv python tools/kernel_times.py > gpurun_out/d_ncu.log 2>&1
tail -n 4 gpurun_out/d_pytest.log
