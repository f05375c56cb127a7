set +e
mkdir -p gpurun_out
rm -f gpurun_out/w_* gpurun_out/tcf_check.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/w_gpu.txt
( timeout 700 python -m pytest tests -m gpu -x -q --timeout 200 ) > gpurun_out/w_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w_rc.txt
( timeout 400 python bench.py ) > gpurun_out/w_bench2.json 2> gpurun_out/w_bench2.err; echo "bench2 rc=$?" >> gpurun_out/w_rc.txt
( timeout 400 python bench.py --impl reference --steps 8 --warmup 3 ) > gpurun_out/w_bench_ref.json 2> gpurun_out/w_bench_ref.err; echo "benchref rc=$?" >> gpurun_out/w_rc.txt
for c in 3 4 5; do
( timeout 400 python bench.py --config $c --steps 300 ) > gpurun_out/w_bench$c.json 2> gpurun_out/w_bench$c.err; echo "bench$c rc=$?" >> gpurun_out/w_rc.txt
done
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 prof ) > gpurun_out/w_check.log 2>&1; echo "check rc=$?" >> gpurun_out/w_rc.txt
( PE_LIB_PATH=$PWD/pinn_elastodynamics_b200/libpinn_elasto_base.so PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py f5 f7 ) > gpurun_out/w_check_base.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:resid_tcf -s 3 -c 1 -o gpurun_out/w_tcf_full -f python tests/ncu_target.py tcf 6 ) > gpurun_out/w_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/w_rc.txt
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/w_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/w_ncu_launches.log 2>&1; echo "launches rc=$?" >> gpurun_out/w_rc.txt
cat gpurun_out/w_rc.txt; tail -n 3 gpurun_out/w_pytest_gpu.log; grep -h -E "ms_per_step" gpurun_out/w_check.log gpurun_out/w_check_base.log | grep tcf | cut -c1-170; for f in gpurun_out/w_bench*.json; do cut -c1-220 $f; done
