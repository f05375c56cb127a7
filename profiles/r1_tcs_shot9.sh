set +e
mkdir -p gpurun_out
rm -f gpurun_out/tcp_check.jsonl gpurun_out/s9.log
echo "=== default lib (TCS_EW=8)" >> gpurun_out/s9.log
( PE_CHECK_ONLY=tc3s timeout 50 python tests/tcp_gpu_check.py f5 f7 ) 2>&1 | grep -E '"tc3s"|Error|error|Traceback' >> gpurun_out/s9.log
echo "=== ew12 (TCS_EW=12)" >> gpurun_out/s9.log
( PE_LIB_PATH=$PWD/pinn_elastodynamics_b200/libpinn_elasto_ew12.so PE_CHECK_ONLY=tc3s timeout 50 python tests/tcp_gpu_check.py f5 f7 ) 2>&1 | grep -E '"tc3s"|Error|error|Traceback' >> gpurun_out/s9.log
echo "rc=$?" >> gpurun_out/s9.log
cut -c1-330 gpurun_out/s9.log
