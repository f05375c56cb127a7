set +e
mkdir -p gpurun_out
rm -f gpurun_out/h_* gpurun_out/tcf_check.jsonl
( timeout 300 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_tcs.py -q -x --timeout 100 -k "tcf or not tc3s" ) > gpurun_out/h_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/h_rc.txt
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 prof ) > gpurun_out/h_check.log 2>&1; echo "check rc=$?" >> gpurun_out/h_rc.txt
( timeout 150 python tests/gpu_refgold_report.py tcf ) > gpurun_out/h_refgold.jsonl 2> gpurun_out/h_refgold.err; echo "refgold rc=$?" >> gpurun_out/h_rc.txt
cat gpurun_out/h_rc.txt; tail -5 gpurun_out/h_tcf.log; grep -E "ms_per_step|grad_rel" gpurun_out/h_check.log | cut -c1-420
