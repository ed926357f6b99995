#!/bin/bash
# segment-count target / longest segment / quad-reduce threshold of the bucket reduction, per batch size (scratch/r2_batch_cost.py)
for cfg in "256 64 16384" "64 128 16384" "32 256 16384" "128 128 16384" "64 128 32768" "96 128 16384"; do
  set -- $cfg
  echo "SEG_DIV=$1 MAX_SEGLEN=$2 QUAD_REDUCE_MAX=$3"
  B200ZK_MSM_SEG_DIV=$1 B200ZK_MSM_MAX_SEGLEN=$2 B200ZK_MSM_QUAD_REDUCE_MAX=$3 python scratch/r2_batch_cost.py 2>&1 | tail -10 | sed "s/hist.*accumulate/acc/"
done
