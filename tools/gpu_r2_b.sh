#!/bin/bash
# round-2 call B (2 GPUs): probe experiments, multi-process parity tests (NVLS, rounds), W=2 sweeps
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/b_gpus.txt 2>&1
timeout 200 tools/probe 2 exp > gpurun_out/b_probe2_exp.log 2>&1; echo "rc=$?" >> gpurun_out/b_probe2_exp.log
timeout 700 python -m pytest tests/test_gpu_multiproc.py -q --maxfail 8 --timeout 150 -k "nvls or broadcast or parity or full_size or known_answer or sendrecv or reducescatter" > gpurun_out/b_pytest_mp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest_mp.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tools/sweep.py --algos ll,oneshot,twoshot --sizes 1024,4096,16384,32768,65536,262144,1048576,4194304 > gpurun_out/b_sweep2_small.log 2>&1
timeout 400 $TR tools/sweep.py --algos twoshot,nvls,nvls_pipe,nvls_sym --sizes 16777216,67108864,268435456,1073741824 \
  --variants "pipe_g16:nvls_pipe:granule_bytes=16384;pipe_g64:nvls_pipe:granule_bytes=65536;pipe_b148:nvls_pipe:nvls_blocks=148;pipe_b64:nvls_pipe:nvls_blocks=64;sym_b148:nvls_sym:nvls_blocks=148;sym_b64:nvls_sym:nvls_blocks=64;two_g64:twoshot:granule_bytes=65536;two_g16:twoshot:granule_bytes=16384" \
  > gpurun_out/b_sweep2_large.log 2>&1
timeout 300 $TR tools/sweep.py --ops sendrecv,broadcast,allgather,reducescatter --sizes 100000,1048576,67108864 > gpurun_out/b_sweep2_ops.log 2>&1
tail -4 gpurun_out/b_pytest_mp.log; grep "EXP" gpurun_out/b_probe2_exp.log; grep "^#" gpurun_out/b_sweep2_small.log gpurun_out/b_sweep2_large.log gpurun_out/b_sweep2_ops.log | cut -c1-900
