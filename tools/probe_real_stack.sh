#!/bin/bash
# Probe of the GPU box for the reference's real third-party stack: casadi / pinocchio / osqp.
# Writes gpurun_out/probe_real_stack.log; copied to profiles/ by hand.
out=gpurun_out/probe_real_stack.log
mkdir -p gpurun_out
{
  echo "== date"; date -u
  echo "== python"; python -V; which python
  for m in casadi pinocchio osqp qdldl fatrop eigenpy hppfcl scipy; do
    python -c "import $m, sys; print('$m', 'OK', getattr($m, '__version__', '?'), $m.__file__)" 2>&1 | tail -1
  done
  echo "== pip download (index)"; timeout 60 python -m pip download --no-deps -d /tmp/pd casadi osqp pin 2>&1 | tail -5
  echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "casadi|osqp|pin|qdldl|fatrop" || echo "no casadi/osqp/pin wheels in /opt/wheelhouse"
  echo "== find"; find / -xdev \( -iname "*casadi*" -o -iname "*pinocchio*" -o -iname "*osqp*" -o -iname "*qdldl*" -o -iname "*fatrop*" \) -not -path "/proc/*" -not -path "*/repo/*" -not -path "/tmp/*" 2>/dev/null | head -20
  echo "== conda"; which conda mamba micromamba 2>&1 | head -3
  echo "== cpu"; nproc; lscpu | grep -E "Model name|Socket|Core|Thread" 
  echo "== gpu"; nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv
} > $out 2>&1
cat $out
