#!/bin/bash
N=$1
./tools/gpu_sweep.sh $N stream_counted --exchange stream
MSS_STREAM_TILE_ATOMICS=1 ./tools/gpu_sweep.sh $N stream_tileatomics --exchange stream
