#!/bin/bash
# round-2 call J (8 GPUs, sweep only): the multi-stream NVLS pipeline at W=8 by piece size, next to the rounds kernel, zero-copy and NCCL
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701"
V="st128:nvls_streams;st64:nvls_streams:nvls_streams_piece_bytes=67108864;st32:nvls_streams:nvls_streams_piece_bytes=33554432;st256:nvls_streams:staging_bytes=536870912,nvls_streams_piece_bytes=268435456;st128b64:nvls_streams:nvls_blocks=64;s32:nvls_sym"
timeout 400 $TR8 tools/sweep.py --algos nvls_pipe --sizes 67108864,134217728,268435456,536870912,1073741824 --variants "$V" > gpurun_out/j_sweep8_streams.log 2>&1
grep "^#" gpurun_out/j_sweep8_streams.log | cut -c1-1200; tail -3 gpurun_out/j_sweep8_streams.log | cut -c1-300
