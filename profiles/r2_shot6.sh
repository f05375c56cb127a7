set +e
mkdir -p gpurun_out
rm -f gpurun_out/f_*
( timeout 200 python tests/probe_umma_timing.py ) > gpurun_out/f_umma_timing.txt 2>&1; echo "timing rc=$?" >> gpurun_out/f_rc.txt
( timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 ) > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_rc.txt
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_rc.txt
cat gpurun_out/f_rc.txt; cat gpurun_out/f_umma_timing.txt; tail -15 gpurun_out/f_pytest.log; tail -3 gpurun_out/f_smoke.log
