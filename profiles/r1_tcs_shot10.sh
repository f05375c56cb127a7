set +e
mkdir -p gpurun_out
( PE_LIB_PATH=$PWD/pinn_elastodynamics_b200/libpinn_elasto_ew12.so timeout 60 python -m pytest tests/test_gpu_tcs.py tests/test_gpu_reference_golden.py -x -q ) > gpurun_out/s10_pytest_ew12.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s10_pytest_ew12.log
tail -6 gpurun_out/s10_pytest_ew12.log
