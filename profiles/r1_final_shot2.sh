set +e
mkdir -p gpurun_out
rm -f gpurun_out/g_*.log gpurun_out/g_rc.txt
( time timeout 120 python tests/gpu_refgold_report.py ) > gpurun_out/g_refgold_report.jsonl 2> gpurun_out/g_refgold_report.err; echo "report rc=$?" >> gpurun_out/g_rc.txt
( time timeout 400 python -m pytest tests -q -m gpu ) > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_rc.txt
cat gpurun_out/g_rc.txt; tail -6 gpurun_out/g_pytest.log
