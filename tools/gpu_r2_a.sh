#!/bin/bash
# round-2 call A (1 GPU): loopback parity suite, smoke (plain + under ncu), N=1 bench, ncu --set full of the 1-GPU kernels
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpus.txt 2>&1
nproc >> gpurun_out/a_gpus.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 1 python -c "import os; print(sorted(k for k in os.environ if 'NV' in k or 'INJECT' in k or 'NSIGHT' in k))" > gpurun_out/a_ncu_env.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail 12 --timeout 180 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/a_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/a_smoke_ncu.log
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench1.json 2> gpurun_out/a_bench1.err; echo "rc=$?" >> gpurun_out/a_bench1.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_local_scale|k_allreduce|k_reducescatter" -c 12 -o gpurun_out/a_prof python tools/profile_kernel.py > gpurun_out/a_prof.log 2>&1; echo "rc=$?" >> gpurun_out/a_prof.log
tail -5 gpurun_out/a_pytest.log; cat gpurun_out/a_smoke.log | tail -3; tail -3 gpurun_out/a_smoke_ncu.log; tail -c 600 gpurun_out/a_bench1.json; tail -8 gpurun_out/a_prof.log
