#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_tiles.py "1x8,1x8:gamma_unroll=2,1x8:gamma_unroll=4,1x8:gamma_unroll=2+gamma_chunk_reduce=0" > gpurun_out/p_tiles.log 2>&1; cat gpurun_out/p_tiles.log
