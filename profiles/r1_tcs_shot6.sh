set +e
mkdir -p gpurun_out
rm -f gpurun_out/tcp_check.jsonl gpurun_out/s6.log
for v in base x5; do
  echo "=== $v" >> gpurun_out/s6.log
  ( PE_LIB_PATH=$PWD/pinn_elastodynamics_b200/libpinn_elasto_$v.so PE_CHECK_ONLY=tc3s timeout 100 python tests/tcp_gpu_check.py f5 f7 ) 2>&1 | grep -E '"tc3s"|Error|error|Traceback' >> gpurun_out/s6.log
  echo "rc=$?" >> gpurun_out/s6.log
done
( PE_LIB_PATH=$PWD/pinn_elastodynamics_b200/libpinn_elasto_x5.so timeout 200 python -m pytest tests/test_gpu_tcs.py -q -k "not bit_identical" ) > gpurun_out/s6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s6.log
tail -5 gpurun_out/s6_pytest.log >> gpurun_out/s6.log
cat gpurun_out/s6.log | cut -c1-420
