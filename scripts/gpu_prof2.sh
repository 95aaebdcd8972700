#!/bin/bash
# launch list of the timed steps + full ncu captures of the hot kernels inside the real mult (bench.py --profile-range)
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --profile-range > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
for k in fast_fwd_blockpass fast_fwd_colpass fast_inv_colpass k_extend_fast k_ksk_inner_fast; do
  timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:${k}" -s 1 -c 1 -f -o gpurun_out/prof5_${k} python bench.py --steps 1 --warmup 2 --no-cpu-baseline --profile-range > gpurun_out/ncu5_${k}.log 2>&1
  tail -1 gpurun_out/ncu5_${k}.log | cut -c1-200
done
python - <<'PY'
import torch, time
x = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty_like(x, device="cuda")
for name, a, b in (("h2d", d, x), ("d2h", x, d)):
    a.copy_(b, non_blocking=True); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); print(name, 5 * 64 / 1024 / (time.perf_counter() - t), "GiB/s")
PY
