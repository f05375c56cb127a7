set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/s8_gpus.txt
( time timeout 500 python -m pytest tests/test_gpu_dist.py -x -q ) > gpurun_out/s8_pytest_dist.log 2>&1; echo "pytest_dist rc=$?" >> gpurun_out/s8_rc.txt
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 500 --warmup 5 ) > gpurun_out/s8_bench_2gpu.json 2> gpurun_out/s8_bench_2gpu.err; echo "bench2 rc=$?" >> gpurun_out/s8_rc.txt
cat gpurun_out/s8_rc.txt; tail -5 gpurun_out/s8_pytest_dist.log; cut -c1-400 gpurun_out/s8_bench_2gpu.json
