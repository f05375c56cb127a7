set +e
mkdir -p gpurun_out
rm -f gpurun_out/r_* gpurun_out/tcf_check.jsonl
( timeout 300 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_tcs.py -q -x --timeout 100 -k "tcf or not tc3s" ) > gpurun_out/r_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/r_rc.txt
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 prof ) > gpurun_out/r_check.log 2>&1; echo "check rc=$?" >> gpurun_out/r_rc.txt
( PE_LIB_PATH=$PWD/pinn_elastodynamics_b200/libpinn_elasto_iss1.so PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 ) > gpurun_out/r_check_iss1.log 2>&1; echo "check1 rc=$?" >> gpurun_out/r_rc.txt
( PE_BFGS_TRACE=1 timeout 300 python bench.py --config 5 --steps 200 --no-cpu-baseline --no-e2e ) > gpurun_out/r_bench5.json 2> gpurun_out/r_bench5.err; echo "bench5 rc=$?" >> gpurun_out/r_rc.txt
cat gpurun_out/r_rc.txt; tail -3 gpurun_out/r_tcf.log; grep -E "ms_per_step" gpurun_out/r_check.log gpurun_out/r_check_iss1.log | cut -c1-300; grep "bfgs trace" gpurun_out/r_bench5.err; cut -c1-250 gpurun_out/r_bench5.json
