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
  # launch list of EXACTLY the timed region of the resident pass (cudaProfilerStart/Stop around it), every kernel that runs
  # there (ours, ATen, memsets are not kernels); then one full capture of $NCU_KERNEL
  CMD="python bench.py --steps ${NCU_STEPS:-20} --warmup 3"
  BENCH_PROFILE_REGION=1 timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 &&
  BENCH_PROFILE_REGION=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 \
      --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1; head -12 gpurun_out/launches.txt
  for K in ${NCU_KERNEL:-gn_eval_kernel}; do
    BENCH_PROFILE_REGION=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 \
        -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_full_$K.log 2>&1
    echo "ncu full $K rc=$?"; tail -2 gpurun_out/ncu_full_$K.log
  done
fi
