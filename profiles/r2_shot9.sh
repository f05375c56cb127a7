set +e
mkdir -p gpurun_out
rm -f gpurun_out/i_* gpurun_out/tcf_check.jsonl
( timeout 300 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_tcs.py -q -x --timeout 100 -k "tcf or not tc3s" ) > gpurun_out/i_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/i_rc.txt
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 prof ) > gpurun_out/i_check.log 2>&1; echo "check rc=$?" >> gpurun_out/i_rc.txt
( PE_LIB_PATH=$PWD/pinn_elastodynamics_b200/libpinn_elasto_ew8.so PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 prof ) > gpurun_out/i_check_ew8.log 2>&1; echo "check8 rc=$?" >> gpurun_out/i_rc.txt
( timeout 150 python tests/gpu_refgold_report.py tcf ) > gpurun_out/i_refgold.jsonl 2> gpurun_out/i_refgold.err; echo "refgold rc=$?" >> gpurun_out/i_rc.txt
cat gpurun_out/i_rc.txt; tail -5 gpurun_out/i_tcf.log; grep -E "ms_per_step" gpurun_out/i_check.log gpurun_out/i_check_ew8.log | cut -c1-300
