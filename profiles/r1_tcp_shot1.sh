set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/s1_gpu.txt 2>&1
( time timeout 200 python tests/tcp_gpu_check.py f5 f7 ) > gpurun_out/s1_check.log 2>&1; echo "check rc=$?" >> gpurun_out/s1_rc.txt
( time timeout 240 python -m pytest tests/test_gpu_tcp.py -x -q ) > gpurun_out/s1_pytest_tcp.log 2>&1; echo "pytest_tcp rc=$?" >> gpurun_out/s1_rc.txt
( time timeout 150 python bench.py --engine tc3p --steps 1000 ) > gpurun_out/s1_bench_tc3p.json 2> gpurun_out/s1_bench_tc3p.err; echo "bench_tc3p rc=$?" >> gpurun_out/s1_rc.txt
( time timeout 60 python tests/tcp_gpu_check.py prof ) > gpurun_out/s1_prof.log 2>&1; echo "prof rc=$?" >> gpurun_out/s1_rc.txt
( time timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tc or f7 or ragged or deterministic or refeed or adam" ) > gpurun_out/s1_pytest_parity.log 2>&1; echo "pytest_parity rc=$?" >> gpurun_out/s1_rc.txt
cat gpurun_out/s1_rc.txt
tail -3 gpurun_out/s1_pytest_tcp.log
