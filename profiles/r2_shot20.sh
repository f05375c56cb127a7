set +e
mkdir -p gpurun_out
rm -f gpurun_out/t_* gpurun_out/tcf_check.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/t_gpu.txt
( timeout 300 python -m pytest tests/test_gpu_tcf.py -q -x --timeout 100 ) > gpurun_out/t_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/t_rc.txt
for v in base "" ew16; do
  lib=$PWD/pinn_elastodynamics_b200/libpinn_elasto${v:+_$v}.so
  ( PE_LIB_PATH=$lib PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 ) > gpurun_out/t_check_${v:-new}.log 2>&1; echo "check ${v:-new} rc=$?" >> gpurun_out/t_rc.txt
done
( PE_LIB_PATH=$PWD/pinn_elastodynamics_b200/libpinn_elasto_base.so PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 ) > gpurun_out/t_check_base2.log 2>&1
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py prof ) > gpurun_out/t_prof.log 2>&1; echo "prof rc=$?" >> gpurun_out/t_rc.txt
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:resid_tcf -s 3 -c 1 -o gpurun_out/t_tcf_full -f python tests/ncu_target.py tcf 6 ) > gpurun_out/t_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/t_rc.txt
( timeout 200 compute-sanitizer --tool memcheck python tests/sanitize_target.py tcf ) > gpurun_out/t_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/t_rc.txt
( timeout 300 compute-sanitizer --tool racecheck python tests/sanitize_target.py tcf ) > gpurun_out/t_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/t_rc.txt
( timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 ) > gpurun_out/t_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_rc.txt
cat gpurun_out/t_rc.txt; tail -3 gpurun_out/t_tcf.log; grep -h -E "ms_per_step" gpurun_out/t_check_*.log | cut -c1-200; tail -4 gpurun_out/t_memcheck.log gpurun_out/t_racecheck.log; tail -4 gpurun_out/t_pytest_gpu.log
