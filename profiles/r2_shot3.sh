set +e
mkdir -p gpurun_out
rm -f gpurun_out/c_* gpurun_out/tcf_check.jsonl
( PE_TEST_TC4=1 timeout 150 python -m pytest tests/test_gpu_tc4_forward.py -q -s --timeout 100 ) > gpurun_out/c_fwd.log 2>&1; echo "fwd rc=$?" >> gpurun_out/c_rc.txt
( timeout 200 python -m pytest tests/test_gpu_tcf.py -q --timeout 100 ) > gpurun_out/c_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/c_rc.txt
( timeout 200 python tests/tcf_gpu_check.py f5 prof ) > gpurun_out/c_check.log 2>&1; echo "check rc=$?" >> gpurun_out/c_rc.txt
( timeout 150 python tests/gpu_refgold_report.py tcf ) > gpurun_out/c_refgold.jsonl 2> gpurun_out/c_refgold.err; echo "refgold rc=$?" >> gpurun_out/c_rc.txt
cat gpurun_out/c_rc.txt; grep -E "rel err|passed|failed" gpurun_out/c_fwd.log | head -20; tail -25 gpurun_out/c_tcf.log; tail -12 gpurun_out/c_check.log | cut -c1-900
