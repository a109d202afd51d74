#!/bin/bash
# round-2 call K (8 GPUs): streams pipeline with ramped pieces (sweep), then the full bench.py at N=8 with the final defaults
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711"
V="st128:nvls_streams;st96:nvls_streams:nvls_streams_piece_bytes=100663296;st128u8:nvls_streams:nvls_unroll=8;s32:nvls_sym"
timeout 300 $TR8 tools/sweep.py --algos auto,nvls_pipe --sizes 268435456,402653184,536870912,1073741824 --variants "$V" > gpurun_out/k_sweep8_streams2.log 2>&1
timeout 700 $TR8 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/k_bench8.json 2> gpurun_out/k_bench8.err; echo "rc=$?" >> gpurun_out/k_bench8.err
grep "^#" gpurun_out/k_sweep8_streams2.log | cut -c1-900; tail -2 gpurun_out/k_bench8.err; python -c "
import json
d=json.loads(open('gpurun_out/k_bench8.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['baselines'], d['hook_ms_per_step'], d.get('optional_section_errors'))
print(d['parity']['all_ok'], {k:v for k,v in d['parity'].items() if isinstance(v,dict) and not v.get('ok')})
print(d['comm_bound']); print(d['rllib_ppo_shape']); print(d['p2p'])
for r in d['allreduce_sweep']: print(r)
for r in d['collectives']: print(r)
print(d['roofline'], d['roofline_in_step'])"
