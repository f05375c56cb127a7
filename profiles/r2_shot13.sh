set +e
mkdir -p gpurun_out
rm -f gpurun_out/m_* gpurun_out/tcf_check.jsonl
( timeout 300 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_tcs.py tests/test_gpu_parity.py -q -x --timeout 100 -k "tcf or not tc3s" ) > gpurun_out/m_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/m_rc.txt
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 prof ) > gpurun_out/m_check.log 2>&1; echo "check rc=$?" >> gpurun_out/m_rc.txt
( timeout 200 python tests/bench_lbfgs_phases.py ) > gpurun_out/m_lbfgs_phases.txt 2>&1; echo "lbfgs rc=$?" >> gpurun_out/m_rc.txt
cat gpurun_out/m_rc.txt; tail -3 gpurun_out/m_tcf.log; grep -E "ms_per_step" gpurun_out/m_check.log | cut -c1-300; cat gpurun_out/m_lbfgs_phases.txt
