set +e
mkdir -p gpurun_out
rm -f gpurun_out/z_* gpurun_out/tcf_check.jsonl
( timeout 700 python -m pytest tests -m gpu -x -q --timeout 200 ) > gpurun_out/z_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_rc.txt
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 prof ) > gpurun_out/z_check.log 2>&1; echo "check rc=$?" >> gpurun_out/z_rc.txt
( timeout 400 python bench.py --steps 600 ) > gpurun_out/z_bench2.json 2> gpurun_out/z_bench2.err; echo "bench2 rc=$?" >> gpurun_out/z_rc.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/z_rc.txt
cat gpurun_out/z_rc.txt; tail -n 3 gpurun_out/z_pytest_gpu.log; grep -h -E "ms_per_step" gpurun_out/z_check.log | grep tcf | cut -c1-170; cut -c1-220 gpurun_out/z_bench2.json; tail -n 2 gpurun_out/z_smoke.log
