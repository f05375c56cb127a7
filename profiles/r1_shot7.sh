set +e
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_reference_golden.py tests/test_gpu_tcs.py -q -k "reference_source or engine_auto" ) > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_pytest.log
tail -40 gpurun_out/s7_pytest.log
