#!/bin/bash
# round-2 call E (2 GPUs): fixes from call D re-tested, lane kernel correctness + W=2 timing, p2p with the larger ring,
# bench N=2 parity/p2p blocks, bench N=1, ncu of the multi-GPU kernels (application replay)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_communicator.py tests/test_gpu_rdt.py -q --maxfail 6 --timeout 150 > gpurun_out/e_pytest_comm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest_comm.log
timeout 400 python -m pytest tests/test_gpu_multiproc.py -q --maxfail 6 --timeout 150 -k "nvls or pool" > gpurun_out/e_pytest_nvls.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest_nvls.log
timeout 300 python -m pytest tests/test_gpu_loopback_r2.py tests/test_gpu_loopback.py -q --maxfail 6 --timeout 150 -k "wrong_shape or world_size_one or send_recv or ll_ or pool or multi_reader" > gpurun_out/e_pytest_lb.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest_lb.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651"
timeout 300 $TR tools/sweep.py --algos twoshot,nvls,nvls_pipe --sizes 67108864,268435456,1073741824 --no-nccl \
  --variants "lanes:nvls_lanes;lanes_g32:nvls_lanes:lane_granule_bytes=32768;lanes_l32:nvls_lanes:nvls_lanes=32;lanes_l96:nvls_lanes:nvls_lanes=96" > gpurun_out/e_sweep2_lanes.log 2>&1
timeout 200 $TR tools/sweep.py --ops sendrecv --sizes 100000,1048576,16777216,67108864,268435456 > gpurun_out/e_sweep2_p2p.log 2>&1
timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-sweep --no-comm-bound --no-nccl-ddp > gpurun_out/e_bench2.json 2> gpurun_out/e_bench2.err; echo "rc=$?" >> gpurun_out/e_bench2.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench1.json 2> gpurun_out/e_bench1.err; echo "rc=$?" >> gpurun_out/e_bench1.err
timeout 500 ncu --target-processes application-only --replay-mode application --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --metrics nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_allreduce_nvls|k_broadcast_rounds|k_send" -c 6 -o gpurun_out/e_prof_multi python tools/profile_multi.py 2 > gpurun_out/e_prof_multi.log 2>&1; echo "rc=$?" >> gpurun_out/e_prof_multi.log
tail -4 gpurun_out/e_pytest_comm.log; tail -4 gpurun_out/e_pytest_nvls.log; tail -4 gpurun_out/e_pytest_lb.log; grep "^#" gpurun_out/e_sweep2_lanes.log gpurun_out/e_sweep2_p2p.log | cut -c1-900; tail -2 gpurun_out/e_bench2.err; python -c "
import json
d=json.loads(open('gpurun_out/e_bench2.json').read().strip().splitlines()[-1]); print({k:v for k,v in d['parity'].items() if 'ddp' in k or k=='all_ok'}); print(d['p2p']); print(d.get('optional_section_errors'))
d=json.loads(open('gpurun_out/e_bench1.json').read().strip().splitlines()[-1]); print(d['roofline']); print(d.get('optional_section_errors'))"
tail -8 gpurun_out/e_prof_multi.log
