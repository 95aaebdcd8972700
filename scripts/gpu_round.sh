#!/bin/bash
# one GPU-box session: parity tests, smoke, NTT sweep, bench arms, ncu launch list + full captures
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
echo "=== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "=== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== ntt sweep" ; timeout 600 python scripts/ntt_bench.py 2>&1 | tee gpurun_out/ntt_bench.log | tail -22
echo "=== bench ours" ; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_ours.log
echo "=== bench reference (cpu port)" ; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref_cpu.log
echo "=== bench reference_gpu" ; timeout 900 python bench.py --impl reference_gpu --steps 10 --warmup 2 2>&1 | tail -2 | tee gpurun_out/bench_ref_gpu.log
if [ "$1" == "ncu" ]; then
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fast_fwd_blockpass -c 1 -o gpurun_out/prof_fast_fwd_block python scripts/ntt_bench.py --quick > gpurun_out/ncu_f1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fast_fwd_colpass -c 1 -o gpurun_out/prof_fast_fwd_col python scripts/ntt_bench.py --quick > gpurun_out/ncu_f2.log 2>&1
fi
ls -la gpurun_out | head -40
