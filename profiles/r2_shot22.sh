set +e
mkdir -p gpurun_out
rm -f gpurun_out/v_* gpurun_out/tcf_check.jsonl
( timeout 300 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_reference_golden.py -q -x --timeout 100 ) > gpurun_out/v_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/v_rc.txt
P=$PWD/pinn_elastodynamics_b200
for v in adjall new fu2 ew16 base new; do
  case $v in
    new) lib=$P/libpinn_elasto.so;;
    *) lib=$P/libpinn_elasto_$v.so;;
  esac
  ( PE_LIB_PATH=$lib PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 ) >> gpurun_out/v_check_$v.log 2>&1; echo "check $v rc=$?" >> gpurun_out/v_rc.txt
done
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py prof ) > gpurun_out/v_prof.log 2>&1; echo "prof rc=$?" >> gpurun_out/v_rc.txt
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/v_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/v_ncu_launches.log 2>&1; echo "launches rc=$?" >> gpurun_out/v_rc.txt
cat gpurun_out/v_rc.txt; tail -n 3 gpurun_out/v_tcf.log; grep -H -E "ms_per_step" gpurun_out/v_check_*.log | grep tcf | cut -c1-170
