#!/bin/bash
# A/B: priority of the side stream the streamed exchange runs on (-1 = above the scoring stream, 0 = same)
N=$1
MSS_STREAM_PRIORITY=-1 ./tools/gpu_sweep.sh $N prio_high --exchange stream
MSS_STREAM_PRIORITY=0 ./tools/gpu_sweep.sh $N prio_same --exchange stream
