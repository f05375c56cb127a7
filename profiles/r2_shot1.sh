set +e
mkdir -p gpurun_out
rm -f gpurun_out/a_*
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
( timeout 200 python tests/probe_umma.py ) > gpurun_out/a_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/a_rc.txt
( PE_TEST_TC4=1 timeout 300 python -m pytest tests/test_gpu_tc4_forward.py -q -x --timeout 120 ) > gpurun_out/a_tc4fwd.log 2>&1; echo "tc4fwd rc=$?" >> gpurun_out/a_rc.txt
( PE_TEST_TC4=1 timeout 300 python -m pytest tests/test_gpu_tc4_forward.py -q --timeout 120 ) > gpurun_out/a_tc4fwd_all.log 2>&1; echo "tc4fwd_all rc=$?" >> gpurun_out/a_rc.txt
( PE_TEST_TC4=1 timeout 400 python -m pytest tests/test_gpu_tc4.py -q --timeout 120 ) > gpurun_out/a_tc4.log 2>&1; echo "tc4 rc=$?" >> gpurun_out/a_rc.txt
( PE_CHECK_ONLY=tc3s,tc4 timeout 300 python tests/tcp_gpu_check.py f5 ) > gpurun_out/a_ab.log 2>&1; echo "ab rc=$?" >> gpurun_out/a_rc.txt
cat gpurun_out/a_rc.txt; tail -30 gpurun_out/a_tc4fwd.log; tail -30 gpurun_out/a_tc4.log; tail -12 gpurun_out/a_ab.log
