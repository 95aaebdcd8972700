"""bench.py -- HE mults/sec (ct x ct + relinearize) at logN=16 (gold preset), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1: one rank per GPU, NCCL)

A "step" is one engine.mult(ct_a, ct_b, evk) = rescale x2 + 4 NTT + tensor product + 3 iNTT + hybrid key switch
(ModUp -> beta*E NTTs -> evk inner product -> 2 iNTT -> ModDown) on level-0 gold ciphertexts.
  value : mults/s with ciphertexts and keys resident in HBM (CUDA events, max over ranks); the step is the engine
          call replayed as a CUDA graph (engine.capture; --graph off launches it eagerly: same kernels, +4 %)
  e2e   : mults/s through the same engine call with the two input ciphertexts in pinned HOST memory and the
          result read back to host inside the timed region (keys stay resident: they are long-lived operands)
  roofline : the batched forward NTT of the key switch ([E', N] limbs per call), timed alone with CUDA events;
          algorithmic bytes = 16 * E' * N per transform against the measured HBM copy peak
  cpu_baseline / --impl reference : the oracle port (oracle/engine_oracle.OracleEngine, C + OpenMP) doing the same
          mult on the host cores -- the reference has no CPU implementation of its own.
N > 1 shards the RNS limbs of ONE multiplication over the ranks (the reference's own partitioning) with one
all_gather (ModUp digits) and one packed broadcast (rescale limbs) per step: total work is fixed -> "strong" scaling.
Working set (2 ciphertexts + evk ~ 507 MB at gold) exceeds the 126 MB L2, so no explicit L2 flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "liberate-fhe_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

PRESET = "gold"
METRIC = "he_mults_per_sec_logN16"
UNIT = "mult/s"


def measured_peak():
    """HBM copy bandwidth in GB/s: the driver-written MEASURED_PEAKS.json ("hbm_gbs"; any numeric entry whose key
    mentions hbm is accepted, TB/s values are converted), else the profiling guide's fallback of 6650 GB/s"""
    def find(obj):
        if isinstance(obj, dict):
            if isinstance(obj.get("hbm_gbs"), (int, float)):
                return float(obj["hbm_gbs"])
            for k, v in obj.items():
                if "hbm" in str(k).lower() and isinstance(v, (int, float)) and v > 0:
                    return float(v) * (1000.0 if v < 100 else 1.0)
            for v in obj.values():
                r = find(v)
                if r:
                    return r
        return None
    try:
        v = find(json.loads((ROOT / "MEASURED_PEAKS.json").read_text()))
        if v:
            return v, "measured"
    except Exception:
        pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed regions.  In-process NVML (nvidia-ml-py) every 10 ms;
    falls back to an `nvidia-smi -lms` child process.  (The child process was the default at first: its polling
    stalled host<->device copies for tens of ms at a time and made the e2e leg jump between 50 and 600 mult/s.)"""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []          # (sm_mhz, max_mhz, reasons bitmask or set)
        self.active = False
        self.stop = False
        self.thread = None
        self.proc = None
        self.nvml = None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                              "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
                self.thread = threading.Thread(target=self._read_smi, daemon=True)
                self.thread.start()
            except Exception:
                self.proc = None
        return self

    def resume(self):
        self.active = True

    def pause(self):
        self.active = False

    def _poll_nvml(self):
        nv = self.nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
        except Exception:
            mx = 0
        while not self.stop:
            if self.active:
                try:
                    sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                    mask = reasons_fn(self.handle)
                    self.samples.append((float(sm), float(mx), {n for n, b in names.items() if mask & b}))
                except Exception:
                    pass
            time.sleep(0.01)

    def _read_smi(self):
        for line in self.proc.stdout:
            if not self.active:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            rs = {n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9])
                  if v.lower().startswith("active")}
            self.samples.append((sm, mx, rs))

    def __exit__(self, *a):
        self.stop = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self):
        sm = sorted(x[0] for x in self.samples)
        mx = max([x[1] for x in self.samples], default=0)
        reasons = set().union(*[x[2] for x in self.samples]) if self.samples else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def roofline_traffic():
    """DRAM bytes per launch of the roofline kernel pair from the committed ncu capture (profiles/), or None"""
    try:
        return float(json.loads((ROOT / "profiles" / "r01_roofline_traffic.json").read_text())["per_launch_traffic_bytes"])
    except Exception:
        return None


def gold_params():
    from liberate_b200.fhe.presets import params
    return {k: v for k, v in params[PRESET].items() if k != "devices"}


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
    """the oracle port of the same mult on the host cores (bounded sample: `steps` whole multiplications)"""
    import numpy as np
    from liberate_b200.fhe.context import ckks_context
    from oracle import oracle as O
    from oracle.engine_oracle import OracleEngine
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ctx = ckks_context(**gold_params())
    K = ctx.num_special_primes
    eng = OracleEngine(ctx.q, ctx.logN, K)
    rng = np.random.default_rng(0)
    L0 = eng.L0
    q = np.array(ctx.q, dtype=np.int64)[:, None]
    mk = lambda rows: rng.integers(0, q[rows], (len(rows), eng.N), dtype=np.int64)
    a = (mk(list(range(L0))), mk(list(range(L0))))
    b = (mk(list(range(L0))), mk(list(range(L0))))
    allrows = list(range(L0 + K))
    evk = [(mk(allrows), mk(allrows)) for _ in eng.partitions]   # random key material: same arithmetic, same bytes
    # bounded sample: whole multiplications, as many of the requested steps as fit in ~150 s of CPU time
    t0 = time.perf_counter()
    for _ in range(min(args.warmup, 1)):
        eng.mult(a, b, evk, 0)
    one = max(time.perf_counter() - t0, 1e-3) if args.warmup else 5.0
    steps = max(1, min(args.steps, int(150.0 / one)))
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.mult(a, b, evk, 0)
    dt = time.perf_counter() - t0
    v = steps / dt
    args.steps = steps
    cores = O.C.num_threads()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "gold preset (logN=16, 35 ordinary + 4 special limbs) ct*ct mult + relinearize, level 0"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} whole multiplications by the C/OpenMP oracle port (the reference has no CPU path)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_reference_gpu(args):
    """EXTRA arm (not part of the driver contract): the reference's own engine + CUDA kernels, installed under
    oracle/_ref/site by oracle/build_ref.py --engine, timed on the same box for the same workload."""
    import numpy as np
    import torch
    from oracle import ref_engine
    if not ref_engine.available():
        print(json.dumps({"impl": "reference_gpu", "unavailable": "oracle/_ref/site not installed"}))
        return
    ref_fhe, cache = ref_engine.load()
    eng = ref_fhe.ckks_engine(devices=[0], cache_folder=cache, **gold_params())
    sk = eng.create_secret_key()
    pk = eng.create_public_key(sk)
    evk = eng.create_evk(sk)
    m = eng.example(-1, 1)
    a, b = eng.encorypt(m, pk), eng.encorypt(m, pk)
    for _ in range(args.warmup):
        eng.mult(a, b, evk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = eng.mult(a, b, evk)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    err = float(np.abs(eng.decrode(out, sk) - m * m).max())
    print(json.dumps({"impl": "reference_gpu", "metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "dtype": "int64", "data": "synthetic", "decrypt_error": err,
                      "config": {"workload": "gold preset ct*ct mult + relinearize, level-0 inputs; the reference's own "
                                             "ckks_engine + ntt_cuda kernels compiled for sm_100 (unmodified sources)"}}))


# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    import liberate_b200
    from liberate_b200 import fhe
    from liberate_b200._lib import lib, check
    from liberate_b200.csprng import Csprng

    params = gold_params()
    if world > 1:
        eng = fhe.ckks_engine(devices=[f"cuda:{local_rank}"] * world, distributed=True, **params)
    else:
        eng = fhe.ckks_engine(devices=[local_rank], **params)
    # identical sampler state on every rank so that the replicated channels agree
    eng.rng = Csprng(eng.ctx.N, [len(d) for d in eng.ntt.p.d], max(eng.ntt.num_special_primes, 2),
                     devices=eng.ntt.devices, local_ids=eng.local_ids, seed=20260925)
    sk = eng.create_secret_key()
    pk = eng.create_public_key(sk)
    evk = eng.create_evk(sk)
    rs = np.random.default_rng(1)
    ma = rs.uniform(-1, 1, eng.num_slots) + 1j * rs.uniform(-1, 1, eng.num_slots)
    mb = rs.uniform(-1, 1, eng.num_slots) + 1j * rs.uniform(-1, 1, eng.num_slots)
    ct_a, ct_b = eng.encorypt(ma, pk), eng.encorypt(mb, pk)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup, finish=None):
        for _ in range(warmup):
            fn()
        if finish:
            finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.launches
        e0.record()
        for _ in range(steps):
            out = fn()
        if finish:
            finish()                # side streams join the timed stream before the closing event
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1)) / steps, (lib.launches - n0) // steps, out

    clk = ClockSampler(local_rank).__enter__()      # samples while the device-timed legs run (value, roofline)
    clk.resume()
    time.sleep(0.1)
    if args.profile_range:
        torch.cuda.profiler.start()
    use_graph = args.graph in ("on", "auto")
    graphed = None
    if use_graph:
        try:
            graphed = eng.capture(eng.mult, ct_a, ct_b, evk)
        except Exception as e:      # capture is an optimisation: fall back to eager launches, and say so
            if world > 1:
                raise               # (ranks must agree on the collectives they issue: no per-rank fallback)
            print(f"[bench] CUDA graph capture failed ({e!r}); running eagerly", file=sys.stderr)
            use_graph = False
    if graphed is not None:        # same kernels and collectives, replayed as ONE graph launch per step
        n_kernels = lib.launches
        eng.mult(ct_a, ct_b, evk)
        n_kernels = lib.launches - n_kernels

        def step():
            graphed.replay()
            return graphed.result
    else:
        def step():
            return eng.mult(ct_a, ct_b, evk)
    ms, launches, prod = timed(step, args.steps, args.warmup)
    if graphed is not None:
        launches = n_kernels        # kernels inside one replay (the library's counter only sees eager launches)
    if args.profile_range:
        torch.cuda.profiler.stop()

    # correctness guard inside the bench: the product decrypts to ma*mb
    if rank == 0 and world == 1:
        err = float(np.abs(eng.decrode(prod, sk) - ma * mb).max())
        assert err < 1e-6, f"bench product does not decrypt: {err}"

    clk.pause()
    # ---- e2e: host-resident operands, result back to host ----------------------------------------------
    def pinned(ct):
        return [[t.cpu().pin_memory() if t is not None else None for t in poly] for poly in ct.data]

    ha, hb = pinned(ct_a), pinned(ct_b)
    out_host = [[torch.empty_like(t, device="cpu").pin_memory() if t is not None else None for t in poly]
                for poly in prod.data]
    h2d = sum(t.numel() * 8 for h in (ha, hb) for poly in h for t in poly if t is not None)
    d2h = sum(t.numel() * 8 for poly in out_host for t in poly if t is not None)
    dev = torch.device(f"cuda:{local_rank}")

    def e2e_serial_step():
        da = [[t.to(dev, non_blocking=True) if t is not None else None for t in poly] for poly in ha]
        db = [[t.to(dev, non_blocking=True) if t is not None else None for t in poly] for poly in hb]
        r = eng.mult(ct_a._replace(data=da), ct_b._replace(data=db), evk)
        for poly, hp in zip(r.data, out_host):
            for t, h in zip(poly, hp):
                if t is not None:
                    h.copy_(t, non_blocking=True)
        return r

    ms_e2e_serial, _, _ = timed(e2e_serial_step, max(5, args.steps // 2), 3)

    # The same calls as a serving loop issues them: every step still copies its own operands host->device and its
    # own product device->host, but on copy streams, two steps deep, so PCIe (both directions) overlaps the kernels.
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    DEPTH = 2
    dev_in = [[[[torch.empty_like(t, device=dev) if t is not None else None for t in poly] for poly in h]
               for h in (ha, hb)] for _ in range(DEPTH)]
    out_hosts = [out_host] + [[[torch.empty_like(t).pin_memory() if t is not None else None for t in poly]
                               for poly in out_host] for _ in range(DEPTH - 1)]
    res_buf = [[[torch.empty_like(t) if t is not None else None for t in poly] for poly in prod.data]
               for _ in range(DEPTH)] if graphed is not None else None
    ev_in = [torch.cuda.Event() for _ in range(DEPTH)]
    ev_used = [torch.cuda.Event() for _ in range(DEPTH)]
    ev_out = [torch.cuda.Event() for _ in range(DEPTH)]
    step_no = [0]

    def e2e_step():
        k = step_no[0] % DEPTH
        step_no[0] += 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_used[k])             # the mult that last read this operand buffer has finished
            for h, d in zip((ha, hb), dev_in[k]):
                for hp, dp in zip(h, d):
                    for th, td in zip(hp, dp):
                        if th is not None:
                            td.copy_(th, non_blocking=True)
            ev_in[k].record(s_in)
        main.wait_event(ev_in[k])
        if graphed is not None:     # the graph reads the captured operand tensors: refresh them in place, then replay
            for src, dst in ((dev_in[k][0], ct_a.data), (dev_in[k][1], ct_b.data)):
                for sp, dp in zip(src, dst):
                    for ts, td in zip(sp, dp):
                        if ts is not None:
                            td.copy_(ts, non_blocking=True)
            graphed.replay()
            # the graph rewrites its result tensors on every replay: hand a per-slot copy to the D2H stream
            main.wait_event(ev_out[k])
            for poly, bp in zip(graphed.result.data, res_buf[k]):
                for t, bt in zip(poly, bp):
                    if t is not None:
                        bt.copy_(t, non_blocking=True)
            r = graphed.result._replace(data=res_buf[k])
        else:
            r = eng.mult(ct_a._replace(data=dev_in[k][0]), ct_b._replace(data=dev_in[k][1]), evk)
        ev_used[k].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_used[k])
            s_out.wait_event(ev_out[k])             # (host buffer k is free: the caller consumed step i-DEPTH)
            for poly, hp in zip(r.data, out_hosts[k]):
                for t, h in zip(poly, hp):
                    if t is not None:
                        t.record_stream(s_out)
                        h.copy_(t, non_blocking=True)
            ev_out[k].record(s_out)
        return r

    def e2e_finish():
        main.wait_stream(s_in)
        main.wait_stream(s_out)

    ms_e2e, _, _ = timed(e2e_step, max(20, args.steps), 4, finish=e2e_finish)
    if world == 1:   # the pipelined products reach the host intact
        torch.cuda.synchronize()
        for k in range(DEPTH):
            for poly, hp in zip(prod.data, out_hosts[k]):
                for t, h in zip(poly, hp):
                    assert t is None or torch.equal(t.cpu(), h), "pipelined e2e product differs from the resident one"

    # ---- roofline of the dominant kernel pair: the key switch's batched forward NTT ----------------------
    # (fast_fwd_colpass + fast_fwd_blockpass over all partitions' extended limbs: [parts*E, N] rows per launch)
    clk.resume()
    peak, peak_kind = measured_peak()
    d0 = eng.local_ids[0]
    level = 1
    plan = eng._plan(level, d0)
    E, N, logN, parts = plan.E, eng.ctx.N, eng.ctx.logN, len(plan.sids)
    rows = parts * E
    buf = torch.randint(0, 1 << 40, (rows, N), dtype=torch.int64, device=dev)   # 200 MB at gold: > L2
    st = torch.cuda.current_stream().cuda_stream
    dsc = plan.desc

    def ntt_call():
        check(lib.ckks_ntt_fast(buf.data_ptr(), N, rows, E, logN, dsc.twf_u64, dsc.twf_f64, dsc.twpf_u64, dsc.twpf_f64,
                                dsc.q, dsc.qinv, None, None, 0, st), "ntt_fast")

    if args.profile_roofline:        # ncu --profile-from-start off: exactly two launches of the measured kernel pair
        ntt_call()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        ntt_call()
        ntt_call()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    ms_ntt, _, _ = timed(ntt_call, 20, 3)
    achieved = 16.0 * rows * N / (ms_ntt * 1e-3) / 1e9
    clk.__exit__()
    clocks = clk.summary()

    line = {
        "metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": "gold preset (logN=16, 35 ordinary + 4 special limbs) ct*ct mult + relinearize, level-0 inputs",
                   "parallelism": f"rns-limb-shard{world}", "cuda_graph": bool(use_graph), "l2": "working set 507 MB > 126 MB L2, no flush needed",
                   "arithmetic": "results bit-identical to the reference; FP64 error-free + Shoup butterflies inside the fused path"},
        "clocks": clocks,
        "e2e": {"value": 1e3 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "mode": "every step copies its operands H2D and its product D2H (pinned host memory) on copy streams, "
                        "2 steps in flight", "serial_value": 1e3 / ms_e2e_serial},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "fast_fwd_colpass + fast_fwd_blockpass_w (ckks_ntt_fast: the key switch's batched forward NTT, 380 limbs)",
                     "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": roofline_traffic(), "algorithmic_bytes": 16.0 * rows * N,
                     "limbs_per_launch": rows, "ms_per_launch": ms_ntt,
                     "note": "algorithmic bytes = 16 B per coefficient per transform (SURVEY 8d); traffic = dram read+write "
                             "per launch from profiles/r01_roofline_traffic.json (ncu, warm caches); FP64-pipe bound: the "
                             "register-only butterfly ceiling is 56 % of the HBM roofline, see DESIGN.md 6"},
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(eng, ct_a, ct_b, evk, prod)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        torch.cuda.synchronize()
        dist.barrier()
        if graphed is not None:
            # a captured graph keeps NCCL work alive; tearing the process group down under it can block for minutes.
            # Every rank has printed / synchronised: leave without the NCCL teardown.
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


def cpu_baseline(eng, ct_a, ct_b, evk, prod):
    """one whole multiplication by the oracle port on the host cores, on the SAME operands (also checks the GPU
    result bit for bit)"""
    import numpy as np
    from oracle import oracle as O
    from oracle.engine_oracle import OracleEngine
    n = lambda t: t.cpu().numpy()
    orc = OracleEngine(eng.ctx.q, eng.ctx.logN, eng.ctx.num_special_primes)
    a = (n(ct_a.data[0][0]), n(ct_a.data[1][0]))
    b = (n(ct_b.data[0][0]), n(ct_b.data[1][0]))
    keys = [(n(p.data[0][0]), n(p.data[1][0])) for p in evk.data]
    t0 = time.perf_counter()
    out, _ = orc.mult(a, b, keys, 0)
    dt = time.perf_counter() - t0
    exact = bool((out[0] == n(prod.data[0][0])).all() and (out[1] == n(prod.data[1][0])).all())
    return {"value": 1.0 / dt, "unit": UNIT, "cores": O.C.num_threads(), "kind": "port",
            "sample": "1 whole gold multiplication by the C/OpenMP oracle port on the same operands",
            "gpu_result_bit_exact": exact}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step (engine.capture) as a CUDA graph; auto = on")
    ap.add_argument("--profile-roofline", action="store_true", help="cudaProfilerStart/Stop around two launches of the roofline kernel pair")
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference_gpu":
        run_reference_gpu(args)
    else:
        run_ours(args)
