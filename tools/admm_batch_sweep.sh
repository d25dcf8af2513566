#!/bin/bash
# ADMM kernel variants (latency: 512 threads, 1 CTA/SM; throughput: 256 threads, 4 CTAs/SM) over the batch size.
for b in 74 148 296 444 512 592 1184; do
  for lat in 0 1000000; do
    echo "batch $b latency_max_batch $lat: $(PLM_ADMM_LATENCY_MAX_BATCH=$lat python tools/prof_sqp.py --batch $b --steps 3 | tail -1)"
  done
done
