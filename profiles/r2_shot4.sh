set +e
mkdir -p gpurun_out
rm -f gpurun_out/d_* gpurun_out/tcf_check.jsonl
( timeout 200 python -m pytest tests/test_gpu_tcf.py -q --timeout 100 ) > gpurun_out/d_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/d_rc.txt
( PE_CHECK_ENGINES=tc3s,tcf timeout 200 python tests/tcf_gpu_check.py f5 prof ) > gpurun_out/d_check.log 2>&1; echo "check rc=$?" >> gpurun_out/d_rc.txt
( PE_LIB_PATH=$PWD/pinn_elastodynamics_b200/libpinn_elasto_ew8.so PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 prof ) > gpurun_out/d_check_ew8.log 2>&1; echo "check8 rc=$?" >> gpurun_out/d_rc.txt
cat gpurun_out/d_rc.txt; tail -5 gpurun_out/d_tcf.log; grep ms_per_step gpurun_out/d_check.log gpurun_out/d_check_ew8.log | cut -c1-300
