set +e
mkdir -p gpurun_out
rm -f gpurun_out/p_* gpurun_out/tcf_check.jsonl
( timeout 300 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_tcs.py tests/test_gpu_lbfgs.py -q -x --timeout 100 -k "tcf or not tc3s" ) > gpurun_out/p_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/p_rc.txt
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 prof ) > gpurun_out/p_check.log 2>&1; echo "check rc=$?" >> gpurun_out/p_rc.txt
( timeout 200 python tests/bench_lbfgs_phases.py ) > gpurun_out/p_lbfgs_phases.txt 2>&1; echo "lbfgs rc=$?" >> gpurun_out/p_rc.txt
( timeout 300 python bench.py --config 5 --steps 200 --no-cpu-baseline ) > gpurun_out/p_bench5.json 2> gpurun_out/p_bench5.err; echo "bench5 rc=$?" >> gpurun_out/p_rc.txt
cat gpurun_out/p_rc.txt; tail -3 gpurun_out/p_tcf.log; grep -E "ms_per_step" gpurun_out/p_check.log | cut -c1-300; grep direction gpurun_out/p_lbfgs_phases.txt; cut -c1-300 gpurun_out/p_bench5.json
