set +e
mkdir -p gpurun_out
( timeout 300 python -m cProfile -s tottime bench.py --config 5 --steps 150 --no-cpu-baseline --no-e2e ) > gpurun_out/q_cprofile.txt 2>&1
head -45 gpurun_out/q_cprofile.txt | cut -c1-200
