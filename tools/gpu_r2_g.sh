#!/bin/bash
# round-2 call G (8 GPUs, sweep only): does the staged NVLS path starve the switch?  multimem vectors in flight per thread (4 vs 8),
# stage order of the rounds kernel, lane kernel with larger bursts
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29681"
V="p_u8:nvls_pipe:nvls_unroll=8;p_u8g64:nvls_pipe:nvls_unroll=8,granule_bytes=65536;p_u8g128:nvls_pipe:nvls_unroll=8,granule_bytes=131072;p_o1:nvls_pipe:rounds_order=1;p_o1u8g64:nvls_pipe:rounds_order=1,nvls_unroll=8,granule_bytes=65536;n_u8g64:nvls:nvls_unroll=8,granule_bytes=65536;ln_l32g128u8:nvls_lanes:nvls_lanes=32,lane_granule_bytes=131072,nvls_unroll=8;ln_l48g256u8:nvls_lanes:nvls_lanes=48,lane_granule_bytes=262144,nvls_unroll=8;ln_l64g128u8:nvls_lanes:nvls_lanes=64,lane_granule_bytes=131072,nvls_unroll=8;s32u8:nvls_sym:nvls_unroll=8;s32:nvls_sym"
timeout 400 $TR8 tools/sweep.py --algos nvls_pipe --sizes 67108864,268435456,1073741824 --variants "$V" > gpurun_out/g_sweep8_unroll.log 2>&1
grep "^#" gpurun_out/g_sweep8_unroll.log | cut -c1-1800; tail -3 gpurun_out/g_sweep8_unroll.log | cut -c1-300
