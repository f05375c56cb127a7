set +e
mkdir -p gpurun_out
( time timeout 300 ncu --set full --clock-control none --import-source on -k regex:resid_tcs -s 3 -c 1 -f -o gpurun_out/r1_tcs_full python tests/ncu_target.py tc3s 5 ) > gpurun_out/s5_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/s5_ncu.log
ls -la gpurun_out/ >> gpurun_out/s5_ncu.log
tail -15 gpurun_out/s5_ncu.log
