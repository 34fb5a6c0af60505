mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; echo "ref rc=$?" >> gpurun_out/f_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 40 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/f_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_gather_rows -s 3 -c 1 -o gpurun_out/f_k_gather_rows -f python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/f_ncu_gather.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/f_all_launches.csv python tools/kernel_times.py > gpurun_out/f_ncu_all.log 2>&1
tail -n 2 gpurun_out/f_pytest.log gpurun_out/f_smoke.log gpurun_out/f_bench.err gpurun_out/f_ref.err; wc -l gpurun_out/f_bench.json gpurun_out/f_ref.json
