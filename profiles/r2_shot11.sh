set +e
mkdir -p gpurun_out
rm -f gpurun_out/k_* gpurun_out/tcf_check.jsonl
( timeout 300 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_tcs.py -q -x --timeout 100 -k "tcf or not tc3s" ) > gpurun_out/k_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/k_rc.txt
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 prof ) > gpurun_out/k_check.log 2>&1; echo "check rc=$?" >> gpurun_out/k_rc.txt
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:resid_tcf -s 3 -c 1 -o gpurun_out/k_tcf_full -f python tests/ncu_target.py tcf 6 ) > gpurun_out/k_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/k_rc.txt
( timeout 300 python bench.py --steps 300 ) > gpurun_out/k_bench2.json 2> gpurun_out/k_bench2.err; echo "bench2 rc=$?" >> gpurun_out/k_rc.txt
( timeout 300 python bench.py --config 3 --steps 100 ) > gpurun_out/k_bench3.json 2> gpurun_out/k_bench3.err; echo "bench3 rc=$?" >> gpurun_out/k_rc.txt
( timeout 300 python bench.py --config 5 --steps 200 ) > gpurun_out/k_bench5.json 2> gpurun_out/k_bench5.err; echo "bench5 rc=$?" >> gpurun_out/k_rc.txt
cat gpurun_out/k_rc.txt; tail -3 gpurun_out/k_tcf.log; grep -E "ms_per_step" gpurun_out/k_check.log | cut -c1-300; cut -c1-400 gpurun_out/k_bench2.json gpurun_out/k_bench3.json gpurun_out/k_bench5.json; tail -3 gpurun_out/k_bench3.err gpurun_out/k_bench5.err
