#!/bin/bash
# round-2 call D (2 GPUs): full GPU test suite incl. the new single-GPU R2/RDT/pool tests and the 2-process
# communicator/channel/RDT/pool tests; bench.py at N=2 (validates parity / p2p / comm-bound blocks); ncu of the multi-GPU kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail 10 --timeout 180 > gpurun_out/d_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest_all.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/d_bench2.json 2> gpurun_out/d_bench2.err; echo "rc=$?" >> gpurun_out/d_bench2.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench1.json 2> gpurun_out/d_bench1.err; echo "rc=$?" >> gpurun_out/d_bench1.err
timeout 400 ncu --target-processes application-only --set full --clock-control none --import-source on -k regex:"k_allreduce|k_broadcast|k_send" -c 10 -o gpurun_out/d_prof_multi python tools/profile_multi.py 2 > gpurun_out/d_prof_multi.log 2>&1; echo "rc=$?" >> gpurun_out/d_prof_multi.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_local_scale" -c 4 -o gpurun_out/d_prof_tma python tools/profile_kernel.py > gpurun_out/d_prof_tma.log 2>&1; echo "rc=$?" >> gpurun_out/d_prof_tma.log
tail -15 gpurun_out/d_pytest_all.log; tail -3 gpurun_out/d_bench2.err; tail -c 1500 gpurun_out/d_bench2.json; tail -c 700 gpurun_out/d_bench1.json; tail -12 gpurun_out/d_prof_multi.log; tail -5 gpurun_out/d_prof_tma.log
