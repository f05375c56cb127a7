set +e
mkdir -p gpurun_out
rm -f gpurun_out/n_* gpurun_out/tcf_check.jsonl
( PE_PROF_CTA=100 PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py prof ) > gpurun_out/n_check100.log 2>&1; echo "check100 rc=$?" >> gpurun_out/n_rc.txt
( PE_PROF_CTA=140 PE_CHECK_ENGINES=tcf timeout 200 python tests/tcf_gpu_check.py prof ) > gpurun_out/n_check140.log 2>&1; echo "check140 rc=$?" >> gpurun_out/n_rc.txt
cat gpurun_out/n_rc.txt
