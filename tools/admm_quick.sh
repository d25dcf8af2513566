#!/bin/bash
# quick A/B of the QP kernels: parity tests of the QP path, then phase times at 592 / 148 / 8192 instances
python -m pytest tests/test_gpu_qp.py tests/test_gpu_sqp.py -m gpu -q -x 2>&1 | tail -3
for b in 148 592 2368; do
  echo "batch $b: $(python tools/prof_sqp.py --batch $b --steps 3 | tail -1)"
done
echo "batch 8192: $(python tools/prof_sqp.py --batch 8192 --steps 4 | tail -1)"
