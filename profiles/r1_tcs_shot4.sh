set +e
mkdir -p gpurun_out
rm -f gpurun_out/tcp_check.jsonl
for v in base x2 x3 x23; do
  echo "=== $v" >> gpurun_out/s4.log
  ( PE_LIB_PATH=$PWD/pinn_elastodynamics_b200/libpinn_elasto_$v.so PE_CHECK_ONLY=tc3s timeout 100 python tests/tcp_gpu_check.py f5 f7 ) 2>&1 | grep -E '"tc3s"|Error|error|Traceback' >> gpurun_out/s4.log
  echo "rc=$?" >> gpurun_out/s4.log
done
cat gpurun_out/s4.log | cut -c1-400
