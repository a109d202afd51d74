#!/bin/bash
# round-2 call I (2 GPUs): correctness of the multi-stream NVLS pipeline (tests) + W=2 timing sanity
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multiproc.py -q --maxfail 4 --timeout 150 -k "nvls" > gpurun_out/i_pytest_nvls.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest_nvls.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29691"
timeout 300 $TR tools/sweep.py --algos nvls_pipe,nvls_streams,nvls_sym --sizes 268435456,1073741824 --no-nccl --variants "st64:nvls_streams:nvls_streams_piece_bytes=67108864" > gpurun_out/i_sweep2_streams.log 2>&1
tail -6 gpurun_out/i_pytest_nvls.log; grep "^#" gpurun_out/i_sweep2_streams.log | cut -c1-700; tail -3 gpurun_out/i_sweep2_streams.log | cut -c1-300
