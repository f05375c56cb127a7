set +e
mkdir -p gpurun_out
rm -f gpurun_out/l_* gpurun_out/tcf_check.jsonl
( timeout 300 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_tcs.py -q -x --timeout 100 -k "tcf or not tc3s" ) > gpurun_out/l_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/l_rc.txt
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 prof ) > gpurun_out/l_check.log 2>&1; echo "check rc=$?" >> gpurun_out/l_rc.txt
( PE_LIB_PATH=$PWD/pinn_elastodynamics_b200/libpinn_elasto_ew8.so PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 ) > gpurun_out/l_check_ew8.log 2>&1; echo "check8 rc=$?" >> gpurun_out/l_rc.txt
cat gpurun_out/l_rc.txt; tail -3 gpurun_out/l_tcf.log; grep -E "ms_per_step" gpurun_out/l_check.log gpurun_out/l_check_ew8.log | cut -c1-300
