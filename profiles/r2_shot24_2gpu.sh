set +e
mkdir -p gpurun_out
rm -f gpurun_out/x_* gpurun_out/tcf_check.jsonl
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/x_gpus.txt
( timeout 300 python -m pytest tests/test_gpu_tcf.py -q -x --timeout 120 ) > gpurun_out/x_tcf.log 2>&1; echo "tcf rc=$?" >> gpurun_out/x_rc.txt
( PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py fields ) > gpurun_out/x_fields.log 2>&1; echo "fields rc=$?" >> gpurun_out/x_rc.txt
( timeout 600 python -m pytest tests/test_gpu_dist.py -q -x --timeout 300 ) > gpurun_out/x_dist.log 2>&1; echo "dist rc=$?" >> gpurun_out/x_rc.txt
( timeout 300 python bench.py --steps 600 --no-cpu-baseline ) > gpurun_out/x_bench2_1gpu.json 2> gpurun_out/x_bench2_1gpu.err; echo "bench 1gpu rc=$?" >> gpurun_out/x_rc.txt
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 600 --no-cpu-baseline ) > gpurun_out/x_bench2_2gpu.json 2> gpurun_out/x_bench2_2gpu.err; echo "bench 2gpu rc=$?" >> gpurun_out/x_rc.txt
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --config 4 --steps 300 --no-cpu-baseline ) > gpurun_out/x_bench4_2gpu.json 2> gpurun_out/x_bench4_2gpu.err; echo "bench4 2gpu rc=$?" >> gpurun_out/x_rc.txt
( PE_PDL=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 600 --no-cpu-baseline --no-e2e ) > gpurun_out/x_bench2_2gpu_nopdl.json 2> gpurun_out/x_bench2_2gpu_nopdl.err; echo "bench 2gpu nopdl rc=$?" >> gpurun_out/x_rc.txt
cat gpurun_out/x_rc.txt; tail -n 3 gpurun_out/x_tcf.log; tail -n 3 gpurun_out/x_dist.log; grep fields gpurun_out/x_fields.log | cut -c1-400; for f in gpurun_out/x_bench*.json; do cut -c1-200 $f; done
