#!/bin/bash
# First gpurun probe: environment facts + pipe microbenchmark (round 1).
mkdir -p gpurun_out
{
  echo "== nvidia-smi"; nvidia-smi -L; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv
  echo "== host"; nproc; grep -m1 "model name" /proc/cpuinfo; free -g | head -2
  echo "== python"; python -c "import jax, jaxlib; print('jax', jax.__version__)" 2>&1 | tail -1
  ls baseline/_ref 2>&1 | head -3
  echo "== pipes"; ./profiles/microbench/pipes
  ./profiles/microbench/pipes
} > gpurun_out/probe_r01.log 2>&1
cat gpurun_out/probe_r01.log
