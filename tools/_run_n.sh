mkdir -p gpurun_out
(TAG=pscan timeout 200 python tools/step_probe.py; TAG=pscan timeout 200 python tools/step_probe.py) > gpurun_out/n_step.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_alt_paths.py -x -q -k "postprocess or fused or nms or oversized or head or full_size or detect or kernel_forms or fuzz_parity" > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
cat gpurun_out/n_step.log; tail -n 3 gpurun_out/n_pytest.log
