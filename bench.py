"""bench.py -- HE mults/sec (ct x ct + relinearize) at logN=16 (gold preset), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference_gpu]
                    [--workload gold-mult|platinum-depth10|ntt-sweep]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1: one rank per GPU, NCCL)

Default workload (the driver's): a "step" is one engine.mult(ct_a, ct_b, evk) = rescale x2 + 4 NTT + tensor product +
3 iNTT + hybrid key switch (ModUp -> beta*E NTTs -> evk inner product -> 2 iNTT -> ModDown) on level-0 gold ciphertexts.
  value      : mults/s with ciphertexts and keys resident in HBM (CUDA events, max over ranks); the step is the engine
               call replayed as a CUDA graph (engine.capture; --graph off launches it eagerly: same kernels)
  e2e        : mults/s through the same engine call with the two input ciphertexts in pinned HOST memory and the
               result read back to host inside the timed region (keys stay resident: they are long-lived operands)
  sustained  : the same step replayed back to back for >= 2 s, with SM clock and board power sampled alongside
  roofline   : the batched forward NTT of the key switch ([beta*E', N] limbs per call, exactly as the executor runs it),
               timed alone with CUDA events; algorithmic bytes = 16 * rows * N per transform against the
               measured HBM copy peak
  reference_cuda : the reference's OWN engine + CUDA kernels (oracle/_ref/site, built unmodified from /root/reference)
               doing the same mult on the same box with the same steps / warm-up; at N = 1 on the very same keys and
               ciphertexts, with the output compared bit for bit against ours
  n_gt1_bit_exact (N > 1): every rank's output rows compared with a single-process run of the same engine on rank 0's GPU
  pipeline   : BASELINE config 4 -- per-op latency of mult and rotate_galois(delta = 1) at this N
  ntt_sweep  : BASELINE config 3 (N = 1 only) -- batched NTT GB/s for N in 2^14..2^17 x L in {1,2,4,8,16,32,60}
  platinum_depth10 (N = 8, or --platinum / --workload platinum-depth10): BASELINE config 5
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
# one string for both arms: the driver compares the two config dicts
WORKLOAD = "gold preset (logN=16, 35 ordinary + 4 special limbs) ct*ct mult + relinearize, level-0 inputs"


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
    """SM clock, board power and throttle reasons sampled DURING the timed regions.  In-process NVML (nvidia-ml-py) every
    10 ms; falls back to an `nvidia-smi -lms` child process.  (The child process was the default at first: its polling
    stalled host<->device copies for tens of ms at a time and made the e2e leg jump between 50 and 600 mult/s.)"""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []          # (sm_mhz, max_mhz, reasons, power_w)
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

    def mark(self):
        return len(self.samples)

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
                    try:
                        pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                    except Exception:
                        pw = None
                    self.samples.append((float(sm), float(mx), {n for n, b in names.items() if mask & b}, pw))
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
            try:
                pw = float(f[3])
            except ValueError:
                pw = None
            rs = {n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9])
                  if v.lower().startswith("active")}
            self.samples.append((sm, mx, rs, pw))

    def __exit__(self, *a):
        self.stop = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self, start=0, end=None):
        part = self.samples[start:end]
        sm = sorted(x[0] for x in part)
        mx = max([x[1] for x in part], default=0)
        pw = [x[3] for x in part if x[3] is not None]
        reasons = set().union(*[x[2] for x in part]) if part else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def roofline_traffic(rows):
    """DRAM bytes per launch of the roofline kernel pair from the committed ncu capture (profiles/), or None when the
    capture was taken for another number of rows (N > 1: every rank transforms fewer limbs)"""
    for name in ("r02_roofline_traffic.json", "r01_roofline_traffic.json"):
        try:
            d = json.loads((ROOT / "profiles" / name).read_text())
            if int(d.get("limbs_per_launch", 380)) == int(rows):
                return float(d["per_launch_traffic_bytes"])
        except Exception:
            pass
    return None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the oracle port of the same mult on the host cores (bounded sample: `steps` whole
    multiplications).  Loads NO product code: the prime chain comes from oracle/params.py (the shipped data file), the
    thread count is set here (torchrun exports OMP_NUM_THREADS=1 to its ranks)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as O
    from oracle.engine_oracle import OracleEngine
    from oracle.params import preset_chain
    O.C.set_num_threads(host_cores())
    q, logN, K = preset_chain(PRESET)
    eng = OracleEngine(q, logN, K)
    rng = np.random.default_rng(0)
    L0 = eng.L0
    qa = np.array(q, dtype=np.int64)[:, None]
    mk = lambda rows: rng.integers(0, qa[rows], (len(rows), eng.N), dtype=np.int64)
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
    cores = O.C.num_threads()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} whole multiplications by the C/OpenMP oracle port on {cores} threads "
                                   "(the reference has no CPU path)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def preset_params(name):
    from liberate_b200.fhe.presets import params
    return {k: v for k, v in params[name].items() if k != "devices"}


def reference_cuda_leg(torch, np, devices, steps, warmup, shared=None, preset=PRESET):
    """the reference's own ckks_engine + ntt_cuda kernels (oracle/_ref/site) for the same mult.
    shared = (sk, evk, ct_a, ct_b, ma*mb, our product): run on OUR tensors and compare the output bit for bit."""
    from oracle import ref_engine
    if not ref_engine.available():
        return {"unavailable": "oracle/_ref/site not installed (python oracle/build_ref.py --engine in the build container)"}
    ref_fhe, cache = ref_engine.load()
    eng = ref_fhe.ckks_engine(devices=devices, cache_folder=cache, **preset_params(preset))
    if shared is not None:
        sk, evk, a, b, want, ours = shared
    else:
        sk = eng.create_secret_key()
        pk = eng.create_public_key(sk)
        evk = eng.create_evk(sk)
        m = eng.example(-1, 1)
        a, b, want, ours = eng.encorypt(m, pk), eng.encorypt(m, pk), m * m, None
    # cc_mult is what mult dispatches to for two ciphertexts of one level (engine.py:2252 -> :2236 -> :1072); calling it
    # directly lets the reference take OUR data_struct tuples (its dispatch table is keyed by its own class object)
    for _ in range(warmup):
        eng.cc_mult(a, b, evk)
    for d in devices:
        torch.cuda.synchronize(d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        out = eng.cc_mult(a, b, evk)
    e1.record()
    for d in devices:
        torch.cuda.synchronize(d)
    wall = (time.perf_counter() - t0) / steps * 1e3
    # one device: CUDA events on its stream; several devices in one process: the host clock around a full sync of all of them
    ms = e0.elapsed_time(e1) / steps if len(devices) == 1 else wall
    res = {"value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup, "devices": len(devices),
           "decrypt_error": float(np.abs(eng.decrode(out, sk) - want).max()),
           "timing": "cuda events" if len(devices) == 1 else "host clock around torch.cuda.synchronize of every device",
           "engine": "the reference's ckks_engine.mult -> cc_mult -> relinearize (src/liberate/fhe/ckks_engine.py:2252, "
                     ":1072-1151) on its own ntt_cuda extension compiled for sm_100 (sources unmodified)"}
    if ours is not None:
        res["bit_exact_vs_ours"] = all(bool(torch.equal(x, y)) for pa, pb in zip(out.data, ours.data) for x, y in zip(pa, pb))
        res["inputs"] = "the same secret key, evaluation key and ciphertexts as our arm"
    else:
        res["inputs"] = "keys and ciphertexts generated by the reference engine itself (its sampler cannot be seeded)"
    del eng
    return res


def run_reference_gpu(args):
    """EXTRA arm: only the reference's CUDA path (what reference_cuda in the default line holds)"""
    import numpy as np
    import torch
    res = reference_cuda_leg(torch, np, list(range(args.gpus)), args.steps, args.warmup)
    print(json.dumps({"impl": "reference_gpu", "metric": METRIC, "n_gpus": args.gpus, "higher_is_better": True,
                      "dtype": "int64", "data": "synthetic", "config": {"workload": WORKLOAD}, **res}))


# ---------------------------------------------------------------------------------------------------------
class Harness:
    """process-group plumbing + the timing primitive shared by all workloads"""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device(f"cuda:{self.local_rank}")
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        from liberate_b200._lib import lib
        self.lib = lib

    def engine(self, preset, seed=20260925):
        from liberate_b200 import fhe
        from liberate_b200.csprng import Csprng
        params = preset_params(preset)
        if self.world > 1:
            eng = fhe.ckks_engine(devices=[f"cuda:{self.local_rank}"] * self.world, distributed=True, **params)
        else:
            eng = fhe.ckks_engine(devices=[self.local_rank], **params)
        # identical sampler state (key AND nonce) on every rank so that the replicated channels agree
        eng.rng = Csprng(eng.ctx.N, [len(d) for d in eng.ntt.p.d], max(eng.ntt.num_special_primes, 2),
                         devices=eng.ntt.devices, local_ids=eng.local_ids, seed=seed, nonce=seed ^ 0x5DEECE66D)
        return eng

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce_max(self, ms):
        if self.dist is None:
            return ms
        t = self.torch.tensor([ms], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup, finish=None):
        torch = self.torch
        for _ in range(warmup):
            fn()
        if finish:
            finish()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = self.lib.launches
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        if finish:
            finish()                # side streams join the timed stream before the closing event
        e1.record()
        self.barrier()
        return self.reduce_max(e0.elapsed_time(e1)) / steps, (self.lib.launches - n0) // steps, out

    def graphed(self, eng, fn, *a):
        """(step callable, kernels per step, graph)"""
        graph = eng.capture(fn, *a)
        n = self.lib.launches
        fn(*a)
        n = self.lib.launches - n

        def step():
            graph.replay()
            return graph.result
        return step, n, graph

    def finish(self, hard_exit):
        if self.dist is not None:
            self.torch.cuda.synchronize()
            self.dist.barrier()
            if hard_exit:
                # a captured graph keeps NCCL work alive; tearing the process group down under it can block for minutes.
                # Every rank has printed / synchronised: leave without the NCCL teardown.
                sys.stdout.flush()
                sys.stderr.flush()
                os._exit(0)
            self.dist.destroy_process_group()


def gather_rows(H, per_dev):
    """list over logical devices (None for remote ones) -> on rank 0 the full list of tensors (on rank 0's GPU)"""
    torch, dist = H.torch, H.dist
    out = []
    for owner, t in enumerate(per_dev):        # logical device d lives on rank d
        shape = torch.zeros(2, dtype=torch.int64, device="cuda")
        if H.rank == owner and t is not None:
            shape[0], shape[1] = t.size(0), t.size(1)
        dist.broadcast(shape, src=owner)
        r, c = int(shape[0].item()), int(shape[1].item())
        if r == 0:
            out.append(None)
        elif owner == 0:
            out.append(t if H.rank == 0 else None)
        elif H.rank == owner:
            dist.send(t.contiguous(), dst=0)
            out.append(None)
        elif H.rank == 0:
            buf = torch.empty((r, c), dtype=torch.int64, device="cuda")
            dist.recv(buf, src=owner)
            out.append(buf)
        else:
            out.append(None)
    return out


def gather_struct(H, x):
    """a data_struct whose per-device lists are complete on rank 0 (other ranks: lists of None)"""
    from liberate_b200.fhe.data_struct import data_struct
    if isinstance(x.data[0], data_struct):
        return x._replace(data=[gather_struct(H, p) for p in x.data])
    return x._replace(data=[gather_rows(H, poly) for poly in x.data])


def check_against_single_process(H, preset, ct_a, ct_b, evk, prod, rotk=None, rot=None):
    """N > 1: rank 0 collects every rank's operands and result rows and repeats the call with a single-process engine
    (LocalComm, all logical devices on its own GPU): the distributed result must be the same bits, device by device."""
    from liberate_b200 import fhe
    torch = H.torch
    A, B, K, P = (gather_struct(H, x) for x in (ct_a, ct_b, evk, prod))
    R = gather_struct(H, rotk) if rotk is not None else None
    RO = gather_struct(H, rot) if rot is not None else None
    ok = True
    if H.rank == 0:
        solo = fhe.ckks_engine(devices=[H.local_rank] * H.world, **preset_params(preset))

        def same(x, y):
            for pa, pb in zip(x.data, y.data):
                for d in range(max(len(pa), len(pb))):
                    u = pa[d] if d < len(pa) else None
                    v = pb[d] if d < len(pb) else None
                    if (u is None) != (v is None) or (u is not None and not torch.equal(u, v)):
                        return False
            return True
        want = solo.mult(A, B, K)
        ok = same(want, P)
        if R is not None:
            ok = ok and same(solo.rotate_single(want, R), RO)
        del solo
    flag = torch.tensor([1 if ok else 0], device="cuda")
    H.dist.broadcast(flag, src=0)
    return bool(flag.item())


# ---------------------------------------------------------------------------------------------------------
def sweep_primes(logN, count):
    """scale primes, then the 60-bit base / special primes of the N = 2^logN tables"""
    t = json.loads((ROOT / "liberate-fhe_b200" / "liberate_b200" / "fhe" / "cache" / "primes.json").read_text())
    N = 1 << logN
    q = list(t["scale_primes"][f"40,{N}"]) + list(t["message_special_primes"]["60"][str(N)])
    return q[:count]


def ntt_sweep(H, peak):
    """BASELINE config 3: batched NTT throughput for N in 2^14..2^17 x L in {1,2,4,8,16,32,60} limbs (scale primes first,
    then 60-bit primes), forward and inverse timed separately, GB/s of algorithmic traffic (16 B per coefficient).
    Buffers are rotated so that every call reads rows that are not in L2 (>= 256 MB in flight)."""
    torch = H.torch
    from liberate_b200._lib import check, lib
    from liberate_b200.ntt import fused
    out = []
    st = torch.cuda.current_stream().cuda_stream
    P = lambda t: t.data_ptr()
    for logN in (14, 15, 16, 17):
        N = 1 << logN
        q_all = sweep_primes(logN, 60)
        for L in (1, 2, 4, 8, 16, 32, 60):
            q = (q_all * 60)[:L]
            qd = torch.tensor(q, dtype=torch.int64, device="cuda")
            g = torch.Generator(device="cuda").manual_seed(0)
            tw = (torch.randint(0, 1 << 62, (L, N), dtype=torch.int64, device="cuda", generator=g) % qd[:, None]).contiguous()
            tf = fused.fast_tables(tw, qd)
            qinv = fused.reciprocals(qd)
            sc = (torch.randint(1, 1 << 62, (L,), dtype=torch.int64, device="cuda", generator=g) % qd).contiguous()
            sh = [(int(s) << 64) // int(m) for s, m in zip(sc.tolist(), q)]
            sc_sh = torch.tensor([w - (1 << 64) if w >= (1 << 63) else w for w in sh], dtype=torch.int64, device="cuda")
            nbuf = max(2, min(64, (256 << 20) // (L * N * 8) + 1))
            bufs = [(torch.randint(0, 1 << 62, (L, N), dtype=torch.int64, device="cuda", generator=g) % qd[:, None]).contiguous()
                    for _ in range(nbuf)]

            def fwd(b):
                check(lib.ckks_ntt_fast(P(b), N, L, L, logN, P(tf.sh), P(tf.dbl), P(tf.psh), P(tf.pdbl), P(qd), P(qinv),
                                        None, None, 0, 0, st), "ntt_fast")

            def inv(b):
                check(lib.ckks_intt_fast(P(b), N, L, L, logN, P(tf.sh), P(tf.dbl), P(tf.psh), P(tf.pdbl), P(qd), P(qinv),
                                         P(sc), P(sc_sh), 0, 0, 0, st), "intt_fast")
            row = {"logN": logN, "L": L}
            for name, fn in (("fwd", fwd), ("inv", inv)):
                iters = max(20, min(400, int(2e9 / (L * N * 16))))
                for i in range(3):
                    fn(bufs[i % nbuf])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(iters):
                    fn(bufs[i % nbuf])
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / iters * 1e3
                gbs = 16.0 * L * N / us / 1e3
                row[name] = {"us": round(us, 2), "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
            out.append(row)
            del bufs, tf, tw
    return out


def run_ours(args):
    import numpy as np
    import torch
    H = Harness()
    world, rank, local_rank, dist, lib, dev = H.world, H.rank, H.local_rank, H.dist, H.lib, H.dev
    from liberate_b200._lib import check

    eng = H.engine(PRESET)
    sk = eng.create_secret_key()
    pk = eng.create_public_key(sk)
    evk = eng.create_evk(sk)
    rs = np.random.default_rng(1)
    ma = rs.uniform(-1, 1, eng.num_slots) + 1j * rs.uniform(-1, 1, eng.num_slots)
    mb = rs.uniform(-1, 1, eng.num_slots) + 1j * rs.uniform(-1, 1, eng.num_slots)
    ct_a, ct_b = eng.encorypt(ma, pk), eng.encorypt(mb, pk)
    timed = H.timed

    clk = ClockSampler(local_rank).__enter__()      # samples while the device-timed legs run (value, sustained, roofline)
    clk.resume()
    time.sleep(0.1)
    if args.profile_range:
        torch.cuda.profiler.start()
    use_graph = args.graph in ("on", "auto")
    graph = None
    if use_graph:
        try:
            step, n_kernels, graph = H.graphed(eng, eng.mult, ct_a, ct_b, evk)
        except Exception as e:      # capture is an optimisation: fall back to eager launches, and say so
            if world > 1:
                raise               # (ranks must agree on the collectives they issue: no per-rank fallback)
            print(f"[bench] CUDA graph capture failed ({e!r}); running eagerly", file=sys.stderr)
            use_graph = False
    if graph is None:
        def step():
            return eng.mult(ct_a, ct_b, evk)
    ms, launches, prod = timed(step, args.steps, args.warmup)
    if graph is not None:
        launches = n_kernels        # kernels inside one replay (the library's counter only sees eager launches)
    if args.profile_range:
        torch.cuda.profiler.stop()

    # ---- sustained: the same step back to back for >= 2 s (clocks and power sampled over exactly this stretch) ----
    sustained = None
    if not (args.profile_range or args.profile_roofline):
        m0 = clk.mark()
        n_sus = max(args.steps, int(2200.0 / ms) + 1)
        ms_sus, _, _ = timed(step, n_sus, 0)
        sus_clk = clk.summary(m0)
        sustained = {"seconds": ms_sus * n_sus / 1e3, "steps": n_sus, "value": 1e3 / ms_sus, "unit": UNIT,
                     "sm_mhz_median": sus_clk["sm_mhz"], "power_w_max": sus_clk["power_w_max"], "reasons": sus_clk["reasons"]}

    # correctness guard inside the bench: the product decrypts to ma*mb
    if rank == 0:
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
               for _ in range(DEPTH)] if graph is not None else None
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
        if graph is not None:     # the graph reads the captured operand tensors: refresh them in place, then replay
            for src, dst in ((dev_in[k][0], ct_a.data), (dev_in[k][1], ct_b.data)):
                for sp, dp in zip(src, dst):
                    for ts, td in zip(sp, dp):
                        if ts is not None:
                            td.copy_(ts, non_blocking=True)
            graph.replay()
            # the graph rewrites its result tensors on every replay: hand a per-slot copy to the D2H stream
            main.wait_event(ev_out[k])
            for poly, bp in zip(graph.result.data, res_buf[k]):
                for t, bt in zip(poly, bp):
                    if t is not None:
                        bt.copy_(t, non_blocking=True)
            r = graph.result._replace(data=res_buf[k])
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
    torch.cuda.synchronize()
    for k in range(DEPTH):          # the pipelined products reach the host intact (every rank checks its own rows)
        for poly, hp in zip(prod.data, out_hosts[k]):
            for t, h in zip(poly, hp):
                assert t is None or torch.equal(t.cpu(), h), "pipelined e2e product differs from the resident one"

    # ---- roofline of the dominant kernel pair: the key switch's batched forward NTT ----------------------
    # (fast_fwd_colpass + fast_fwd_blockpass over all partitions' extended limbs: [parts*E, N] rows per launch, NTT-domain
    #  output in the executor's warp-interleaved order -- the call ckks_exec_keyswitch_stage makes, as one launch pair)
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
    perm = 1 if lib.ckks_get_option(18) else 0

    def ntt_call():
        check(lib.ckks_ntt_fast(buf.data_ptr(), N, rows, E, logN, dsc.twf_u64, dsc.twf_f64, dsc.twpf_u64, dsc.twpf_f64,
                                dsc.q, dsc.qinv, None, None, 0, perm, st), "ntt_fast")

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
    del buf

    # ---- BASELINE config 4: mult -> rotate_galois(delta = 1), per-op latency at this N -----------------------
    rotk = eng.create_rotation_key(sk, 1)
    gk1 = eng._ct([rotk], 0, "galk", include_special=True, ntt_state=True, montgomery_state=True)   # the delta = 1 member of a Galois key
    rot_fn = lambda c: eng.rotate_galois(c, gk1, 1)
    if use_graph:
        rstep, _, rgraph = H.graphed(eng, rot_fn, prod)
    else:
        rstep, rgraph = (lambda: rot_fn(prod)), None
    ms_rot, _, rot = timed(rstep, args.steps, args.warmup)
    if rank == 0:
        err_rot = float(np.abs(eng.decrode(rot, sk) - np.roll(ma * mb, 1)).max())
        assert err_rot < 1e-6, f"bench rotation does not decrypt: {err_rot}"
    pipeline = {"workload": "gold mult (rescale inside) -> rotate_galois(delta=1), level-0 inputs", "mult_ms": ms,
                "rotate_ms": ms_rot, "ops_per_s": 2e3 / (ms + ms_rot)}
    # hoisted rotations: 8 rotations of one ciphertext sharing one ModUp, against 8 single rotations
    keys8 = [rotk] + [eng.create_rotation_key(sk, 1 << i) for i in range(1, 8)]
    hoist_fn = lambda c: eng.rotate_hoisted(c, keys8)
    single_fn = lambda c: [eng.rotate_single(c, k) for k in keys8]
    if use_graph:       # both as CUDA graphs, like the other legs (eager, the comparison measures the host's launch rate)
        hstep, _, _hg = H.graphed(eng, hoist_fn, prod)
        sstep8, _, _sg = H.graphed(eng, single_fn, prod)
    else:
        hstep, sstep8 = (lambda: hoist_fn(prod)), (lambda: single_fn(prod))
    ms_h8, _, hoisted = timed(hstep, max(5, args.steps // 2), 3)
    ms_s8, _, _ = timed(sstep8, max(5, args.steps // 2), 3)
    if rank == 0:
        err_h = max(float(np.abs(eng.decrode(o, sk) - np.roll(ma * mb, 1 << i)).max()) for i, o in enumerate(hoisted))
        assert err_h < 1e-6, f"hoisted rotations do not decrypt: {err_h}"
    pipeline["hoisted_8_rotations_ms"] = ms_h8
    pipeline["single_8_rotations_ms"] = ms_s8
    del keys8, hoisted
    eng.release_key_cache()
    clk.__exit__()
    clocks = clk.summary()

    line = {
        "metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "setup": {"parallelism": f"rns-limb-shard{world}", "cuda_graph": bool(use_graph),
                  "l2": "working set 507 MB > 126 MB L2, no flush needed",
                  "arithmetic": "results bit-identical to the reference; FP64 error-free + Shoup butterflies inside the fused path"},
        "clocks": clocks,
        "e2e": {"value": 1e3 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "mode": "every step copies its operands H2D and its product D2H (pinned host memory) on copy streams, "
                        "2 steps in flight", "serial_value": 1e3 / ms_e2e_serial},
        "sustained": sustained,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "fast_fwd_colpass + fast_fwd_blockpass (ckks_ntt_fast as the key switch calls it: "
                                               f"batched forward NTT of {rows} limbs, warp-interleaved output)",
                     "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": roofline_traffic(rows), "algorithmic_bytes": 16.0 * rows * N,
                     "limbs_per_launch": rows, "ms_per_launch": ms_ntt,
                     "note": "algorithmic bytes = 16 B per coefficient per transform (SURVEY 8d); traffic = dram read+write "
                             "per launch from the committed ncu capture (null when this run transforms another number of "
                             "limbs); the kernel is FP64-pipe bound: the register-only butterfly ceiling is 56 % of the "
                             "HBM roofline (DESIGN.md 6)"},
        "pipeline": pipeline,
    }

    # ---- N > 1: the distributed result is the single-process result, bit for bit (mult and rotate) -----------
    if world > 1:
        line["n_gt1_bit_exact"] = check_against_single_process(H, PRESET, ct_a, ct_b, evk, prod, rotk, rot)
        # what the two collectives cost inside the step: the same graph captured with one / both left out (results are then
        # wrong by construction -- timing only; engine.debug_skip_collectives)
        line["deep_chain"] = deep_chain_check(H, eng, sk, pk, evk, np)
        comm = {"step_ms": ms}
        for name, skip in (("no_digit_all_gather_ms", {"gather"}), ("no_rescale_broadcast_ms", {"bcast"}),
                           ("no_collectives_ms", {"gather", "bcast"})):
            eng.debug_skip_collectives = skip
            if use_graph:
                sstep, _, _g = H.graphed(eng, eng.mult, ct_a, ct_b, evk)
            else:
                sstep = lambda: eng.mult(ct_a, ct_b, evk)
            comm[name], _, _ = timed(sstep, args.steps, args.warmup)
        eng.debug_skip_collectives = set()
        line["comm_breakdown"] = comm

    if rank == 0 and world == 1 and not args.no_sweep:
        line["ntt_sweep"] = {"unit": "GB/s of algorithmic traffic (16 B/coefficient)", "peak": peak,
                             "rows": ntt_sweep(H, peak)}
    if world == 8 or args.platinum:
        try:
            line["platinum_depth10"] = platinum_depth10(H, np, use_graph)
        except Exception as e:
            if world > 1:
                raise
            line["platinum_depth10"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(eng, ct_a, ct_b, evk, prod)
    # ---- the reference's own CUDA path on the same box (LAST: whatever it does to the process cannot cost the line) -------
    # N = 1: in this process, on OUR keys and ciphertexts, output compared bit for bit.  N > 1: rank 0 runs the reference's own
    # multi-GPU mode (one process, devices=[0..N-1], engine.py:56-60) in a CHILD process -- its CUDA context is isolated from
    # ours (an illegal access inside the reference's kernels at 8 devices took rank 0 down with it when it ran in-process) --
    # while the other ranks leave their GPUs alone: they wait on the rendezvous store (host side), not in an NCCL barrier.
    if not args.no_reference_cuda:
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        if rank == 0:
            try:
                if world == 1:
                    line["reference_cuda"] = reference_cuda_leg(torch, np, [local_rank], args.steps, args.warmup,
                                                                shared=(sk, evk, ct_a, ct_b, ma * mb, prod))
                else:
                    env = {k: v for k, v in os.environ.items()
                           if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS",
                                        "TORCHELASTIC_RUN_ID", "GROUP_RANK", "ROLE_RANK", "LOCAL_WORLD_SIZE", "ROLE_WORLD_SIZE")}
                    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference_gpu", "--gpus", str(world),
                                          "--steps", str(args.steps), "--warmup", str(args.warmup)],
                                         capture_output=True, text=True, timeout=420, env=env, cwd=str(ROOT))
                    rows = [x for x in out.stdout.splitlines() if x.startswith("{")]
                    if out.returncode == 0 and rows:
                        got = json.loads(rows[-1])
                        line["reference_cuda"] = {k: v for k, v in got.items() if k not in ("impl", "metric", "config", "data", "dtype")}
                    else:
                        tail = (out.stderr or out.stdout).strip().splitlines()
                        why = next((x for x in reversed(tail) if "Error" in x or "error" in x), tail[-1] if tail else "no output")
                        line["reference_cuda"] = {"unavailable": f"the reference's own {world}-device mode failed on this box "
                                                                 f"(exit {out.returncode}): {why.strip()[:240]}"}
            except Exception as e:
                line["reference_cuda"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        if dist is not None:
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                store.set("bench_reference_cuda_done", "1")
            else:
                import datetime
                store.wait(["bench_reference_cuda_done"], datetime.timedelta(seconds=600))

    if rank == 0:
        print(json.dumps(line), flush=True)
    H.finish(hard_exit=graph is not None)


def deep_chain_check(H, eng, sk, pk, evk, np):
    """N > 1: x <- (x + x) * c down the levels until every rank's device has been the source of a rescale broadcast
    (K * N + 1 levels), then decrypt on rank 0.  The bit-for-bit comparison above cannot see ranks that disagree on shared
    randomness (decryption reads device 0's limbs only, and so would the first levels of such a ciphertext); this does."""
    rs = np.random.default_rng(11)
    m = rs.uniform(-1, 1, eng.num_slots) + 1j * rs.uniform(-1, 1, eng.num_slots)
    m /= np.abs(m).max() * 1.5
    w = 0.5 * np.exp(0.3j)
    x = eng.encorypt(m, pk)
    cw = eng.encorypt(np.full(eng.num_slots, w), pk)
    depth = min(eng.num_levels - 1, eng.ntt.num_special_primes * H.world + 1)
    sources = set()
    v = m
    for _ in range(depth):
        sources.add(eng.ntt.p.rescaler_loc[x.level])
        x = eng.mult(eng.add(x, x), cw, evk)
        v = 2 * v * w
    out = eng.decrode(x, sk)
    res = {"depth": depth, "rescale_sources": sorted(sources)}
    if H.rank == 0:
        res["decrypt_error"] = float(np.abs(out - v).max())
        res["decrypt_ok"] = bool(res["decrypt_error"] < 1e-5)
    return res


def platinum_depth10(H, np, use_graph=True):
    """BASELINE config 5: platinum preset (logN=17, 73 + 6 limbs), depth-10 chain of (add, mult, rotate) from level 0,
    end-to-end HE ops/s over all ranks; the first (add, mult, rotate) is checked against the single-process engine when
    N > 1 and the final ciphertext must decrypt to the plaintext circuit."""
    torch = H.torch
    t0 = time.perf_counter()
    eng = H.engine("platinum", seed=777)
    sk = eng.create_secret_key()
    pk = eng.create_public_key(sk)
    evk = eng.create_evk(sk)
    rotk = eng.create_rotation_key(sk, 1)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    rs = np.random.default_rng(5)
    m = rs.uniform(-1, 1, eng.num_slots) + 1j * rs.uniform(-1, 1, eng.num_slots)
    m /= np.abs(m).max() * 1.5
    w = 0.5 * np.exp(0.3j)                          # (x + x) * w keeps |x|: the chain neither decays nor blows up
    ct = eng.encorypt(m, pk)
    cw = eng.encorypt(np.full(eng.num_slots, w), pk)
    depth = 10

    def circuit(x):
        for _ in range(depth):
            y = eng.add(x, x)                       # 2x
            z = eng.mult(y, cw, evk)                # 2x * w  (cw is a level-0 ciphertext: auto-levelled up to x, engine.py:2225-2240)
            x = eng.rotate_single(z, rotk)          # roll by 1
        return x

    def plain(v):
        for _ in range(depth):
            v = np.roll(2 * v * w, 1)
        return v

    out = circuit(ct)                               # warm-up: plans, workspaces, NCCL channels
    # the whole circuit (30 operations, ~600 kernels, the collectives of every rank) as ONE CUDA graph when it captures on
    # every rank -- eager, the chain is bound by the host's launch rate, which differs from box to box
    step, graphed, why = (lambda: circuit(ct)), False, None
    gstep = _graph = None
    if use_graph:
        ok = 1
        try:
            gstep, _n, _graph = H.graphed(eng, circuit, ct)
        except Exception as e:      # (reported below; every rank falls back together)
            ok, why = 0, f"{type(e).__name__}: {e}"[:200]
        if H.dist is not None:
            flag = torch.tensor([ok], device="cuda")
            H.dist.all_reduce(flag, op=H.dist.ReduceOp.MIN)
            ok = int(flag.item())
        if ok:
            step, graphed = gstep, True
    out = step()
    H.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10 if graphed else 3
    e0.record()
    for _ in range(reps):
        out = step()
    e1.record()
    H.barrier()
    ms = H.reduce_max(e0.elapsed_time(e1)) / reps
    res = {"workload": "platinum preset (logN=17, 73 ordinary + 6 special limbs): depth-10 chain x <- rotate((x + x) * c, 1) of "
                       "(add, ct*ct mult+relin with auto-levelling of c, rotate) from level-0 ciphertexts", "cuda_graph": graphed,
           "n_gpus": H.world, "ops": 3 * depth, "ms_per_circuit": ms,
           "he_ops_per_s": 3 * depth * 1e3 / ms, "setup_s": setup_s, "final_level": out.level}
    if H.rank == 0:
        want = plain(m)
        res["decrypt_error"] = float(np.abs(eng.decrode(out, sk) - want).max())
        res["plain_absmax"] = float(np.abs(want).max())
        res["decrypt_ok"] = bool(res["decrypt_error"] < 1e-5)      # (reported, not asserted: the ranks stay in lockstep)
    if why:
        res["graph_fallback"] = why
    if H.world > 1:
        y = eng.add(ct, ct)
        first = eng.mult(y, ct, evk)
        frot = eng.rotate_single(first, rotk)
        res["first_mult_rotate_bit_exact_vs_single_process"] = check_against_single_process(
            H, "platinum", y, ct, evk, first, rotk, frot)
    step = gstep = _graph = out = None      # (the graph's private pool goes with it)
    del eng
    torch.cuda.empty_cache()
    return res


def run_workload(args):
    """builder-run workloads (the driver runs the default): one JSON line each"""
    import numpy as np
    H = Harness()
    peak, peak_kind = measured_peak()
    if args.workload == "ntt-sweep":
        line = {"workload": "ntt-sweep", "peak": peak, "peak_kind": peak_kind, "rows": ntt_sweep(H, peak)}
    elif args.workload == "platinum-depth10":
        line = {"workload": "platinum-depth10", **platinum_depth10(H, np)}
    else:
        raise SystemExit(f"unknown workload {args.workload}")
    if H.rank == 0:
        print(json.dumps(line), flush=True)
    H.finish(hard_exit=False)


def cpu_baseline(eng, ct_a, ct_b, evk, prod):
    """one whole multiplication by the oracle port on the host cores, on the SAME operands (also checks the GPU
    result bit for bit)"""
    from oracle import oracle as O
    from oracle.engine_oracle import OracleEngine
    O.C.set_num_threads(host_cores())
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
    ap.add_argument("--workload", default="gold-mult", choices=["gold-mult", "ntt-sweep", "platinum-depth10"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--platinum", action="store_true", help="also run the platinum depth-10 circuit (default: only at N = 8)")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step (engine.capture) as a CUDA graph; auto = on")
    ap.add_argument("--profile-roofline", action="store_true", help="cudaProfilerStart/Stop around two launches of the roofline kernel pair")
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.profile_range or args.profile_roofline:      # under ncu: only the kernels of interest
        args.no_reference_cuda = args.no_sweep = args.no_cpu_baseline = True
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference_gpu":
        run_reference_gpu(args)
    elif args.workload != "gold-mult":
        run_workload(args)
    else:
        run_ours(args)
