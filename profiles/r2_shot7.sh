set +e
mkdir -p gpurun_out
( timeout 200 python tests/probe_umma_timing.py ) > gpurun_out/g_umma_timing.txt 2>&1; echo "timing rc=$?"
cat gpurun_out/g_umma_timing.txt
