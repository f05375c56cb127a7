set +e
mkdir -p gpurun_out
rm -f gpurun_out/u_* gpurun_out/tcf_check.jsonl
( timeout 300 python -m pytest tests/test_gpu_tcf.py tests/test_gpu_reference_golden.py -q -x --timeout 100 ) > gpurun_out/u_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/u_rc.txt
P=$PWD/pinn_elastodynamics_b200
for v in base new nopdl nohint adjall base new; do
  case $v in
    new) lib=$P/libpinn_elasto.so; pdl=1;;
    nopdl) lib=$P/libpinn_elasto.so; pdl=0;;
    *) lib=$P/libpinn_elasto_$v.so; pdl=1;;
  esac
  ( PE_PDL=$pdl PE_LIB_PATH=$lib PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 ) >> gpurun_out/u_check_$v.log 2>&1; echo "check $v rc=$?" >> gpurun_out/u_rc.txt
done
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py prof ) > gpurun_out/u_prof.log 2>&1; echo "prof rc=$?" >> gpurun_out/u_rc.txt
( timeout 300 python bench.py --steps 300 ) > gpurun_out/u_bench2.json 2> gpurun_out/u_bench2.err; echo "bench2 rc=$?" >> gpurun_out/u_rc.txt
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/u_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/u_ncu_launches.log 2>&1; echo "launches rc=$?" >> gpurun_out/u_rc.txt
( timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 ) > gpurun_out/u_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_rc.txt
cat gpurun_out/u_rc.txt; tail -n 3 gpurun_out/u_tcf.log; grep -H -E "ms_per_step" gpurun_out/u_check_*.log | grep tcf | cut -c1-170; tail -n 4 gpurun_out/u_pytest_gpu.log; cut -c1-300 gpurun_out/u_bench2.json
