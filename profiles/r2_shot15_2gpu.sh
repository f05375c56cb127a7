set +e
mkdir -p gpurun_out
rm -f gpurun_out/o_*
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/o_gpus.txt
( timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_reference_golden.py -q --timeout 600 ) > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_rc.txt
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 500 --warmup 5 ) > gpurun_out/o_bench_2gpu.json 2> gpurun_out/o_bench_2gpu.err; echo "bench2gpu rc=$?" >> gpurun_out/o_rc.txt
( timeout 300 python bench.py --gpus 1 --steps 500 --warmup 5 --no-cpu-baseline ) > gpurun_out/o_bench_1gpu.json 2> gpurun_out/o_bench_1gpu.err; echo "bench1gpu rc=$?" >> gpurun_out/o_rc.txt
cat gpurun_out/o_rc.txt; tail -15 gpurun_out/o_pytest.log; cut -c1-300 gpurun_out/o_bench_2gpu.json gpurun_out/o_bench_1gpu.json
