#!/bin/bash
# One GPU-box visit: parity tests, bench (ours), launch list + full ncu capture of the dominant kernel.
# Usage (under gpurun): bash tools/gpu_check.sh [tests|bench|ncu ...]   default: all
set -u
mkdir -p gpurun_out
what="${*:-tests bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
if [[ "$what" == *tests* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
fi
if [[ "$what" == *smoke* ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
fi
if [[ "$what" == *bench* ]]; then
  timeout 900 python bench.py --steps ${BENCH_STEPS:-30} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
fi
if [[ "$what" == *ncu* ]]; then
  CMD="python bench.py --steps 2 --warmup 3"
  timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:3dfb -s ${NCU_SKIP:-300} -c ${NCU_COUNT:-1200} \
      --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-sdf_hg_kernel} -s 30 -c 2 \
      -f -o gpurun_out/prof_${NCU_KERNEL:-sdf_hg_kernel} $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
