#!/bin/bash
# round-2 call F (8 GPUs): lane kernel at W=8 (correctness + knobs), bench.py at N=8 (parity / p2p / comm-bound / sweeps), W=4 lanes
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multiproc.py -q --maxfail 4 --timeout 150 -k "nvls_allreduce or nvls_pipelined or pool or broadcast_all" > gpurun_out/f_pytest_mp8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest_mp8.log
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29671"
V="lanes:nvls_lanes;ln_g32:nvls_lanes:lane_granule_bytes=32768;ln_g128:nvls_lanes:lane_granule_bytes=131072;ln_l32:nvls_lanes:nvls_lanes=32;ln_l64:nvls_lanes:nvls_lanes=64;ln_l96:nvls_lanes:nvls_lanes=96;ln_l32g128:nvls_lanes:nvls_lanes=32,lane_granule_bytes=131072;ln_l64g32:nvls_lanes:nvls_lanes=64,lane_granule_bytes=32768;s32:nvls_sym"
timeout 300 $TR8 tools/sweep.py --algos nvls_pipe,nvls --sizes 33554432,67108864,268435456,1073741824 --variants "$V" > gpurun_out/f_sweep8_lanes.log 2>&1
timeout 600 $TR8 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/f_bench8.json 2> gpurun_out/f_bench8.err; echo "rc=$?" >> gpurun_out/f_bench8.err
B200COLL_HOOK_MAX_BLOCKS=64 timeout 300 $TR8 bench.py --gpus 8 --steps 10 --warmup 3 --no-sweep --no-parity --no-p2p --no-nccl-ddp > gpurun_out/f_bench8_hook64.json 2> gpurun_out/f_bench8_hook64.err
export CUDA_VISIBLE_DEVICES=0,1,2,3
timeout 200 $TR4 tools/sweep.py --algos twoshot --sizes 67108864,268435456,1073741824 --variants "lanes:nvls_lanes;ln_l32:nvls_lanes:nvls_lanes=32" > gpurun_out/f_sweep4_lanes.log 2>&1
tail -3 gpurun_out/f_pytest_mp8.log; grep "^#" gpurun_out/f_sweep8_lanes.log gpurun_out/f_sweep4_lanes.log | cut -c1-1500; tail -2 gpurun_out/f_bench8.err; tail -c 3000 gpurun_out/f_bench8.json; tail -c 1200 gpurun_out/f_bench8_hook64.json
