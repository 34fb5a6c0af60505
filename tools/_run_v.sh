mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/v_smoke.log
tail -n 3 gpurun_out/v_pytest.log gpurun_out/v_smoke.log
