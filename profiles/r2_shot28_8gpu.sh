set +e
mkdir -p gpurun_out
rm -f gpurun_out/b8_*
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/b8_gpus.txt
run() { # name, nproc, extra args
  name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then ( timeout 300 python bench.py "$@" ) > gpurun_out/b8_$name.json 2> gpurun_out/b8_$name.err
  else ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$((RANDOM%10)) bench.py --gpus $n "$@" ) > gpurun_out/b8_$name.json 2> gpurun_out/b8_$name.err; fi
  echo "$name rc=$?" >> gpurun_out/b8_rc.txt
}
run c2_n8 8 --steps 600 --no-cpu-baseline
run c2_n1 1 --steps 600 --no-cpu-baseline
run c2_n4 4 --steps 600 --no-cpu-baseline --no-e2e
run c2_n2 2 --steps 600 --no-cpu-baseline --no-e2e
run c4_n8 8 --config 4 --steps 300 --no-cpu-baseline
PE_PEER_ALLREDUCE=0 run c2_n8_nccl 8 --steps 600 --no-cpu-baseline --no-e2e
( timeout 900 python -m pytest tests/test_gpu_dist.py -q -x --timeout 300 -k "tcf or peer" ) > gpurun_out/b8_dist.log 2>&1; echo "dist rc=$?" >> gpurun_out/b8_rc.txt
cat gpurun_out/b8_rc.txt; tail -n 3 gpurun_out/b8_dist.log; for f in gpurun_out/b8_c*.json; do cut -c1-190 $f; done
