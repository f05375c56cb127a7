set +e
mkdir -p gpurun_out
rm -f gpurun_out/j_* gpurun_out/tcf_check.jsonl
( timeout 300 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_tcs.py -q -x --timeout 100 -k "tcf or not tc3s" ) > gpurun_out/j_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/j_rc.txt
( PE_CHECK_ENGINES=tcf,tcf16 timeout 200 python tests/tcf_gpu_check.py f5 f7 prof ) > gpurun_out/j_check.log 2>&1; echo "check rc=$?" >> gpurun_out/j_rc.txt
( timeout 150 python tests/gpu_refgold_report.py tcf ) > gpurun_out/j_refgold.jsonl 2> gpurun_out/j_refgold.err; echo "refgold rc=$?" >> gpurun_out/j_rc.txt
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:resid_tcf -s 3 -c 1 -o gpurun_out/j_tcf_full -f python tests/ncu_target.py tcf 6 ) > gpurun_out/j_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/j_rc.txt
cat gpurun_out/j_rc.txt; tail -5 gpurun_out/j_tcf.log; tail -3 gpurun_out/j_ncu.log; grep -E "ms_per_step" gpurun_out/j_check.log | cut -c1-300
