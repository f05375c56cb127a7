set +e
mkdir -p gpurun_out
rm -f gpurun_out/b_*
( timeout 300 python tests/probe_umma.py ) > gpurun_out/b_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/b_rc.txt
( PE_TEST_TC4=1 timeout 300 python -m pytest tests/test_gpu_tc4_forward.py -q -s --timeout 120 ) > gpurun_out/b_tc4fwd.log 2>&1; echo "tc4fwd rc=$?" >> gpurun_out/b_rc.txt
cat gpurun_out/b_rc.txt; tail -22 gpurun_out/b_probe.log; grep -E "rel err|passed|failed|Error" gpurun_out/b_tc4fwd.log | head -40
