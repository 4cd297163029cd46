set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
tail -5 gpurun_out/r2_t1.log
timeout 300 python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops vox_fwd,vox_fwd_planned,vox_bwd,devox_fwd,devox_fwd_planned,devox_fwd_eval_planned,devox_bwd,devox_bwd_planned,plan_mean > gpurun_out/r2_p1.log 2>&1
timeout 300 python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops devox_fwd --opt devox_fwd=1 > gpurun_out/r2_p1_old.log 2>&1
for v in vox768 vox512; do PCFM_B200_LIB=point-cloud-flow-matching_b200/lib/var/libpcfm_$v.so timeout 300 python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops vox_fwd,vox_fwd_planned > gpurun_out/r2_p1_$v.log 2>&1; done
