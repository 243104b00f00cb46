#!/bin/bash
# A/B: resident CTAs per SM of the streamed exchange kernel (0 = one CTA per tile, the previous behaviour)
N=$1; shift
for c in "$@"; do
  MSS_EXCHANGE_CTAS_PER_SM=$c ./tools/gpu_sweep.sh $N cap$c --exchange stream 2>&1 | grep -v "^\*\*\*\|OMP_NUM"
done
