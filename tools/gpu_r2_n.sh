#!/bin/bash
# round-2 call N (1 GPU, after the per-world-size kernels): what the driver runs on its one-GPU box — the whole GPU suite, smoke (plain and under ncu), bench N=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 180 > gpurun_out/n_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/n_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/n_smoke_ncu.log
timeout 500 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/n_bench1.json 2> gpurun_out/n_bench1.err; echo "rc=$?" >> gpurun_out/n_bench1.err
tail -4 gpurun_out/n_pytest_gpu.log; tail -2 gpurun_out/n_smoke.log; tail -2 gpurun_out/n_smoke_ncu.log; tail -2 gpurun_out/n_bench1.err; python -c "
import json
d=json.loads(open('gpurun_out/n_bench1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline'], d['cpu_baseline'], d.get('optional_section_errors'), d['gpu_launches'])"
