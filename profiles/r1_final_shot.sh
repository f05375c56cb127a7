set +e
mkdir -p gpurun_out
rm -f gpurun_out/f_*.log gpurun_out/f_rc.txt
( time timeout 400 python -m pytest tests -x -q -m gpu ) > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_rc.txt
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_rc.txt
( time timeout 200 python bench.py ) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_rc.txt
( time timeout 200 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "bench_ref rc=$?" >> gpurun_out/f_rc.txt
( time timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/f_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/f_rc.txt
cat gpurun_out/f_rc.txt; tail -4 gpurun_out/f_pytest.log; tail -2 gpurun_out/f_smoke.log; cut -c1-300 gpurun_out/f_bench.json
