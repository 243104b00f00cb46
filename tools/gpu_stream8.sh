#!/bin/bash
N=$1
./tools/gpu_sweep.sh $N auto
./tools/gpu_sweep.sh $N stream --exchange stream
