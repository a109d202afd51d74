#!/bin/bash
# round-2 call M (2 GPUs): per-world-size reduce kernels + packed-byte reduce validated on the multi-process suite,
# A/B of the converting-copy unroll (libb200coll_cu2.so = -DB200C_CONVERT_UNROLL=2) and of a 1 GiB staging area, bench N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multiproc.py tests/test_gpu_communicator.py -x -q -m gpu --timeout 180 > gpurun_out/m_pytest_mp2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest_mp2.log
tail -3 gpurun_out/m_pytest_mp2.log
# mixed (f32 bucket, bf16 wire) two-shot: hook grid (64 CTAs) and default grid, unroll 1 vs 2
for lib in libb200coll.so libb200coll_cu2.so; do
  for mb in 64 296; do
    B200COLL_LIBRARY=$PWD/ant-ray_b200/$lib timeout 200 $TR --master-port 29731 tools/sweep.py --wire bf16 --algos twoshot --sizes 31457280,268435456 --max-blocks $mb --no-nccl > gpurun_out/m_mixed_${lib%.so}_mb$mb.log 2>&1
    echo "$lib mb=$mb: $(grep -o '"bytes": [0-9]*, "twoshot_us": [0-9.]*' gpurun_out/m_mixed_${lib%.so}_mb$mb.log | tr '\n' ' ')"
  done
done
# staging 256 MiB (default) vs 1 GiB: two-shot pieces at 256 MiB / 1 GiB
timeout 300 $TR --master-port 29732 tools/sweep.py --algos twoshot --sizes 268435456,1073741824 --variants "st1g:twoshot:staging_bytes=1073741824" > gpurun_out/m_staging_ab.log 2>&1
grep -o '"bytes": [0-9]*, "st1g_us": [0-9.]*, "st1g_busbw": [0-9.]*, "twoshot_us": [0-9.]*, "twoshot_busbw": [0-9.]*, "nccl_us": [0-9.]*' gpurun_out/m_staging_ab.log
timeout 600 $TR --master-port 29733 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/m_bench2.json 2> gpurun_out/m_bench2.err; echo "rc=$?" >> gpurun_out/m_bench2.err
tail -2 gpurun_out/m_bench2.err; python -c "
import json
d=json.loads(open('gpurun_out/m_bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['baselines'], d['hook_ms_per_step'], d.get('optional_section_errors'))
print(d['parity']['all_ok'], {k:v for k,v in d['parity'].items() if isinstance(v,dict) and not v.get('ok')})
print(d['comm_bound']); print(d['rllib_ppo_shape']); print(d['roofline'])
for r in d['allreduce_sweep']: print(r)"
