#!/bin/bash
# round-2 call L (4 GPUs): bench.py at N=4 with the final defaults
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/l_bench4.json 2> gpurun_out/l_bench4.err; echo "rc=$?" >> gpurun_out/l_bench4.err
tail -2 gpurun_out/l_bench4.err; python -c "
import json
d=json.loads(open('gpurun_out/l_bench4.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['baselines'], d['hook_ms_per_step'], d.get('optional_section_errors'))
print(d['parity']['all_ok'], {k:v for k,v in d['parity'].items() if isinstance(v,dict) and not v.get('ok')})
print(d['comm_bound']); print(d['rllib_ppo_shape']); print(d['p2p'])
for r in d['allreduce_sweep']: print(r)
for r in d['collectives']: print(r)"
