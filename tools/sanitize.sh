#!/bin/bash
# compute-sanitizer over a small SQP step of the bench workload (B2G whole_body_rnea, N = 20, 8 instances): memcheck and
# racecheck of the node kernel, the factor kernel and both instantiations of the ADMM kernel (lock-free panel pipeline).
# Logs -> gpurun_out/sanitizer_*.log (copied to profiles/ by hand).
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for lat in 0 1000000; do
    name=$([ $lat = 0 ] && echo throughput || echo latency)
    log=gpurun_out/sanitizer_${tool}_${name}.log
    PLM_ADMM_LATENCY_MAX_BATCH=$lat timeout 900 compute-sanitizer --tool $tool --print-limit 20 \
      python tools/prof_sqp.py --batch 8 --steps 1 > $log 2>&1
    echo "rc=$?" >> $log
    tail -5 $log
  done
done
