set +e
mkdir -p gpurun_out
( time timeout 120 python tests/tcp_gpu_check.py profs ) > gpurun_out/s3_profs.log 2>&1; echo "profs rc=$?" >> gpurun_out/s3_rc.txt
cat gpurun_out/s3_rc.txt
