#!/bin/bash
# round-2 call C (8 GPUs): the large-message question (NVLS core rate, staging overlap), small-message floor, other collectives,
# correctness of the multicast paths at W=8; then the same sweep at W=4.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c_gpus.txt 2>&1
timeout 250 tools/probe 8 exp > gpurun_out/c_probe8_exp.log 2>&1; echo "rc=$?" >> gpurun_out/c_probe8_exp.log
timeout 420 python -m pytest tests/test_gpu_multiproc.py -q --maxfail 6 --timeout 150 -k "all_gpus or nvls or full_size" > gpurun_out/c_pytest_mp8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest_mp8.log
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631"
V_SYM="s32:nvls_sym:nvls_blocks=32;s64:nvls_sym:nvls_blocks=64;s96:nvls_sym:nvls_blocks=96;s148:nvls_sym:nvls_blocks=148;s148g16:nvls_sym:nvls_blocks=148,granule_bytes=16384;s148g64:nvls_sym:nvls_blocks=148,granule_bytes=65536;s296g16:nvls_sym:granule_bytes=16384;s296g64:nvls_sym:granule_bytes=65536;s296g128:nvls_sym:granule_bytes=131072"
V_STG="p_g64:nvls_pipe:granule_bytes=65536;p_g128:nvls_pipe:granule_bytes=131072;p_g256:nvls_pipe:granule_bytes=262144;p_g128b148:nvls_pipe:granule_bytes=131072,nvls_blocks=148;n_b148:nvls:nvls_blocks=148;n_g64:nvls:granule_bytes=65536;t_g64:twoshot:granule_bytes=65536"
timeout 420 $TR8 tools/sweep.py --algos twoshot,nvls,nvls_pipe,nvls_sym --sizes 16777216,67108864,268435456,1073741824 --variants "$V_SYM;$V_STG" > gpurun_out/c_sweep8_large.log 2>&1
timeout 200 $TR8 tools/sweep.py --algos ll,oneshot,twoshot,nvls --sizes 1024,8192,32768,65536,131072,262144,1048576,4194304 --variants "ll256:ll:ll_max_bytes=262144" > gpurun_out/c_sweep8_small.log 2>&1
timeout 200 $TR8 tools/sweep.py --ops broadcast,allgather,reducescatter,sendrecv --sizes 1048576,16777216,67108864 > gpurun_out/c_sweep8_ops.log 2>&1
B200COLL_SEND_BATCH=2 timeout 120 $TR8 tools/sweep.py --ops sendrecv --sizes 16777216,67108864 --no-nccl > gpurun_out/c_sweep8_send_b2.log 2>&1
export CUDA_VISIBLE_DEVICES=0,1,2,3
timeout 300 $TR4 tools/sweep.py --algos twoshot,nvls,nvls_pipe,nvls_sym --sizes 16777216,67108864,268435456,1073741824 --variants "s148:nvls_sym:nvls_blocks=148;s64:nvls_sym:nvls_blocks=64;p_g128:nvls_pipe:granule_bytes=131072;t_g64:twoshot:granule_bytes=65536" > gpurun_out/c_sweep4_large.log 2>&1
tail -3 gpurun_out/c_pytest_mp8.log; grep "EXP" gpurun_out/c_probe8_exp.log; grep "^#" gpurun_out/c_sweep8_large.log gpurun_out/c_sweep8_small.log gpurun_out/c_sweep8_ops.log gpurun_out/c_sweep8_send_b2.log gpurun_out/c_sweep4_large.log | cut -c1-1800
