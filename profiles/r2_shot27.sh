set +e
mkdir -p gpurun_out
rm -f gpurun_out/a2_* gpurun_out/tcf_check.jsonl
( timeout 300 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_reference_golden.py -q -x --timeout 100 ) > gpurun_out/a2_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/a2_rc.txt
P=$PWD/pinn_elastodynamics_b200
for v in prev new base prev new; do
  case $v in
    new) lib=$P/libpinn_elasto.so;;
    *) lib=$P/libpinn_elasto_$v.so;;
  esac
  ( PE_LIB_PATH=$lib PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 ) >> gpurun_out/a2_check_$v.log 2>&1; echo "check $v rc=$?" >> gpurun_out/a2_rc.txt
done
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py prof ) > gpurun_out/a2_prof.log 2>&1; echo "prof rc=$?" >> gpurun_out/a2_rc.txt
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout 200 ) > gpurun_out/a2_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/a2_rc.txt
cat gpurun_out/a2_rc.txt; tail -n 3 gpurun_out/a2_tcf.log; tail -n 3 gpurun_out/a2_parity.log; grep -H -E "ms_per_step" gpurun_out/a2_check_*.log | grep tcf | cut -c1-170
