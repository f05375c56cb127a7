set +e
mkdir -p gpurun_out
( timeout 300 python -m cProfile -o gpurun_out/q_prof.pstats bench.py --config 5 --steps 100 --no-cpu-baseline --no-e2e ) > /dev/null 2>&1
python - <<'PY' > gpurun_out/q_callers.txt 2>&1
import pstats
p = pstats.Stats('gpurun_out/q_prof.pstats')
p.print_callers('_cuda_getDeviceCount')
p.print_callers('is_available')
p.sort_stats('cumtime').print_stats(35)
PY
cut -c1-220 gpurun_out/q_callers.txt | head -120
