#!/bin/bash
# one GPU-box session: parity tests, smoke, NTT sweep, bench arms, ncu launch list + traffic + full captures
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
echo "=== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "=== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "=== bench ours" ; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours.log | cut -c1-300
echo "=== bench reference (cpu port)" ; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_cpu.log | cut -c1-200
echo "=== bench reference_gpu" ; timeout 900 python bench.py --impl reference_gpu --steps 10 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_ref_gpu.log | cut -c1-200
echo "=== op bench" ; timeout 600 python scripts/op_bench.py 2>&1 | grep '^{' | tee gpurun_out/op_bench.log | cut -c1-400
echo "=== ntt sweep" ; timeout 600 python scripts/ntt_bench.py 2>&1 | tee gpurun_out/ntt_bench.log | tail -4 | cut -c1-300
if [ "$1" == "ncu" ]; then
echo "=== ncu launch list (timed steps)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --profile-range > gpurun_out/ncu_launch.log 2>&1
echo "=== ncu roofline traffic (warm caches)"
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/roofline_traffic.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --profile-roofline > gpurun_out/ncu_roofline.log 2>&1
for k in fast_fwd_blockpass_w fast_fwd_colpass k_ksk_inner_fast k_extend_fast; do
  timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:${k}" -s 1 -c 1 -f -o gpurun_out/prof8_${k} python bench.py --steps 1 --warmup 2 --no-cpu-baseline --profile-range > gpurun_out/ncu8_${k}.log 2>&1
  tail -1 gpurun_out/ncu8_${k}.log | cut -c1-120
done
fi
ls -la gpurun_out | head -40
