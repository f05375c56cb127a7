set +e
mkdir -p gpurun_out
( time PE_CHECK_ONLY=tc3s timeout 120 python tests/tcp_gpu_check.py f5 f7 ) > gpurun_out/s2_check.log 2>&1; echo "check rc=$?" >> gpurun_out/s2_rc.txt
( time timeout 200 python -m pytest tests/test_gpu_tcs.py -x -q ) > gpurun_out/s2_pytest_tcs.log 2>&1; echo "pytest_tcs rc=$?" >> gpurun_out/s2_rc.txt
( time timeout 150 python bench.py --engine tc3s --steps 1000 ) > gpurun_out/s2_bench_tc3s.json 2> gpurun_out/s2_bench_tc3s.err; echo "bench_tc3s rc=$?" >> gpurun_out/s2_rc.txt
cat gpurun_out/s2_rc.txt; tail -4 gpurun_out/s2_check.log; tail -3 gpurun_out/s2_pytest_tcs.log
