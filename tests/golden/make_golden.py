"""
make_golden.py -- generates tests/golden/*.json|npz by running the UNMODIFIED reference Python
(/root/reference/src/liberate) in the build container.  See ref_shims.py for what is stubbed
(only the CUDA extensions: the 15 ntt_cuda ops run on the CPU oracle, the csprng is the seeded
sampler of tests/seeded_rng.py).  Re-run with:   python tests/golden/make_golden.py

Outputs
  partition.json      rns_partition attributes for a sweep of (num_ordinary, K, num_devices)   [part.py]
  context.json        ckks_context scalars for several (logN, scale_bits, K, num_scales) + preset shapes
  tables_logN12.npz   Montgomery constants, compact Montgomery twiddles, index tables (cctx.py / nctx.py)
  ntt_consts_*.json   ntt_context Garner constants (Y_scalar / L_scalar / L_enter) and engine scalars
  engine_D{1,2,3}.json + engine_D*_full.npz   digests of every object of tests/flows.hot_path_flow
"""
import hashlib
import json
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))
import ref_shims  # noqa: E402

CACHE = ref_shims.install()
warnings.simplefilter("ignore")

from liberate import fhe  # noqa: E402
from liberate.fhe.context.ckks_context import ckks_context  # noqa: E402
from liberate.fhe.context.security_parameters import maximum_qbits  # noqa: E402
from liberate.ntt.rns_partition import rns_partition  # noqa: E402
from liberate.fhe.data_struct import data_struct  # noqa: E402
import flows  # noqa: E402

ENGINE_PARAMS = dict(logN=12, num_scales=6, num_special_primes=2, scale_bits=40, is_secured=False)
# the reference's own tests sweep scale_bits 20..45 (src/liberate/fhe/tests/test_generate_engine.py:40-46): one set below the
# FP64 limit of the fused path (q < 2^42) and one above it (every limb on the 64-bit integer path)
EXTRA_SETS = {"_sb30": dict(logN=12, num_scales=5, num_special_primes=2, scale_bits=30, is_secured=False),
              # scale_bits = 42: the cached primes alternate around 2^42 -- FP64 limbs and wide limbs in one partition
              "_sb42": dict(logN=12, num_scales=5, num_special_primes=2, scale_bits=42, is_secured=False),
              "_sb45": dict(logN=12, num_scales=5, num_special_primes=3, scale_bits=45, is_secured=False)}


from golden_utils import Recorder, describe, sha  # noqa: E402,F401


def jsonable(x):
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    return x


def gen_partition():
    out = []
    for L, K, D in [(17, 2, 1), (17, 2, 2), (7, 2, 1), (7, 2, 2), (7, 2, 3), (35, 4, 1), (35, 4, 2), (35, 4, 4),
                    (35, 4, 8), (73, 6, 8), (8, 1, 1), (8, 1, 2), (5, 3, 2), (9, 4, 3)]:
        try:
            p = rns_partition(L, K, D)
        except IndexError:
            # the reference itself cannot partition this combination (a device ends up empty)
            out.append(dict(args=[L, K, D], error="IndexError"))
            continue
        item = dict(args=[L, K, D])
        for name in ("num_partitions", "partitions", "part_allocations", "prime_allocations",
                     "flat_prime_allocations", "base_prime_idx", "destination_arrays_with_special",
                     "destination_arrays", "rescaler_loc", "part_cumsums", "part_counts", "parts",
                     "destination_parts", "destination_parts_with_special", "p", "p_special", "diff", "d",
                     "d_special"):
            item[name] = jsonable(getattr(p, name))
        out.append(item)
    (HERE / "partition.json").write_text(json.dumps(out))


def gen_context():
    out = dict(max_qbits={}, contexts=[])
    for logN in range(12, 18):
        out["max_qbits"][str(logN)] = int(maximum_qbits(2 ** logN, 128, "post_quantum", "uniform"))
    cases = [dict(logN=14, num_special_primes=1, scale_bits=40), dict(logN=15, num_special_primes=2, scale_bits=40),
             dict(logN=16, num_special_primes=4, scale_bits=40), dict(logN=17, num_special_primes=6, scale_bits=40),
             dict(logN=12, num_special_primes=2, scale_bits=40, num_scales=6, is_secured=False),
             dict(logN=13, num_special_primes=3, scale_bits=30, num_scales=5, is_secured=False),
             dict(logN=15, num_special_primes=2, scale_bits=35)]
    for kw in cases:
        # generate_paints (python-int psi series) is slow for large N and irrelevant to the scalars
        paints = ckks_context.generate_paints
        if kw["logN"] > 13:
            ckks_context.generate_paints = lambda self: None
        try:
            c = ckks_context(cache_folder=CACHE, read_cache=False, save_cache=False, **kw)
        finally:
            ckks_context.generate_paints = paints
        out["contexts"].append(dict(args=kw, q=[int(x) for x in c.q], num_scales=c.num_scales,
                                    max_qbits=c.max_qbits, total_qbits=c.total_qbits,
                                    generation_string=c.generation_string,
                                    R_square=[int(x) for x in c.R_square], k=[int(x) for x in c.k]))
    (HERE / "context.json").write_text(json.dumps(out))


def compact(painted, forward):
    return ref_shims._cc.get(torch.from_numpy(np.ascontiguousarray(painted)), forward)


def gen_tables(eng):
    ctx, ntt = eng.ctx, eng.ntt
    N = ctx.N
    psi_m = ntt.psi[0].numpy()      # [C, logN, N/2] Montgomery form (after psi_enter), device order == q order for D=1
    ipsi_m = ntt.ipsi[0].numpy()
    fpsi = compact(psi_m, True)
    ipsi = compact(ipsi_m, False)
    # check that the painted layout is exactly "twiddle m+i for the t butterflies of block i"
    logN = ctx.logN
    for lvl in range(logN):
        m, t = 1 << lvl, N >> (lvl + 1)
        assert (psi_m[:, lvl, :] == np.repeat(fpsi[:, m:2 * m], t, axis=1)).all()
        h, t2 = N >> (lvl + 1), 1 << lvl
        assert (ipsi_m[:, lvl, :] == np.repeat(ipsi[:, h:2 * h], t2, axis=1)).all()
    np.savez_compressed(
        HERE / "tables_logN12.npz",
        q=np.array(ctx.q, dtype=np.int64), order=np.array(ntt.p.d_special[0]),
        Rs=ntt.Rs[0].numpy(), ql=ntt.ql[0].numpy(), qh=ntt.qh[0].numpy(), kl=ntt.kl[0].numpy(), kh=ntt.kh[0].numpy(),
        _2q=ntt._2q[0].numpy(), Ninv=ntt.Ninv[0].numpy(), Rs_scale=ntt.Rs_scale[0].numpy(),
        psi=fpsi, ipsi=ipsi,
        even=ntt.even[0].numpy(), odd=ntt.odd[0].numpy(), ieven=ntt.ieven[0].numpy(), iodd=ntt.iodd[0].numpy())


def gen_ntt_consts(eng, D):
    ntt = eng.ntt
    out = dict(starts=jsonable(ntt.starts), stops=jsonable(ntt.stops), parts={})
    for dev in range(D):
        for key, item in ntt.parts_pack[dev].items():
            if "Y_scalar" in item:
                ent = {}
                ent["Y_scalar"] = None if item["Y_scalar"] is None else item["Y_scalar"].tolist()
                ent["L_scalar"] = None if item["L_scalar"] is None else [x.tolist() for x in item["L_scalar"]]
                ent["L_enter"] = [None if le is None else [x.tolist() for x in le] for le in item["L_enter"]]
                out["parts"][f"{dev}:{','.join(map(str, key))}"] = ent
    out["rescale_scales"] = [[t.tolist() for t in lvl] for lvl in eng.rescale_scales]
    out["PiRs"] = [[[t.tolist() for t in pind] for pind in lvl] for lvl in eng.PiRs]
    out["mont_PR"] = [t.tolist() for t in eng.mont_PR]
    out["final_scalar"] = [t.tolist() for t in eng.final_scalar]
    out["deviations"] = [float(x) for x in eng.deviations]
    out["corrections"] = [float(x) for x in eng.corrections]
    out["parts_alloc"] = jsonable(eng.parts_alloc)
    out["stor_ids"] = jsonable(eng.stor_ids)
    out["hash"] = eng.hash
    (HERE / f"ntt_consts_D{D}.json").write_text(json.dumps(out))


def gen_engine(D, tag="", params=ENGINE_PARAMS):
    eng = fhe.ckks_engine(devices=["cpu"] * D, cache_folder=CACHE, read_cache=False, save_cache=False,
                          **params)
    if D == 1 and not tag:
        gen_tables(eng)
    if not tag:
        gen_ntt_consts(eng, D)
    rec = Recorder()
    objs = flows.hot_path_flow(eng, rec)
    flows.extra_flow(eng, rec, objs)
    # float results: kept in full, compared with a tolerance in the tests
    rec.full["decode_a"] = eng.decrode(objs["ct_a"], objs["sk"])
    rec.full["decode_ab"] = eng.decrode(objs["ct_ab"], objs["sk"])
    rec.full["ma"] = objs["ma"]
    rec.full["mb"] = objs["mb"]
    err = np.abs(rec.full["decode_ab"] - objs["ma"] * objs["mb"]).max()
    print(f"D={D}{tag}: {len(rec.digests)} objects, mult error {err:.2e}")
    assert err < (1e-6 if params["scale_bits"] >= 40 else 1e-3)
    (HERE / f"engine_D{D}{tag}.json").write_text(json.dumps(dict(params=params, q=[int(x) for x in eng.ctx.q],
                                                                 digests=rec.digests)))
    np.savez_compressed(HERE / f"engine_D{D}{tag}_full.npz", **rec.full)
    if D == 1 and not tag:
        # full INPUT tensors of the hot path for the CPU tests of the oracle's own engine restatement
        # (tests/test_oracle_engine.py); outputs stay digests.
        T = lambda t: t.numpy()
        inputs = dict(ct_a0=T(objs["ct_a"].data[0][0]), ct_a1=T(objs["ct_a"].data[1][0]),
                      ct_b0=T(objs["ct_b"].data[0][0]), ct_b1=T(objs["ct_b"].data[1][0]),
                      ct_ab0=T(objs["ct_ab"].data[0][0]), ct_ab1=T(objs["ct_ab"].data[1][0]))
        for name in ("evk", "rotk1"):
            for i, part in enumerate(objs[name].data):
                inputs[f"{name}/{i}/0"] = T(part.data[0][0])
                inputs[f"{name}/{i}/1"] = T(part.data[1][0])
        np.savez_compressed(HERE / "engine_D1_inputs.npz", **inputs)


if __name__ == "__main__":
    gen_partition()
    gen_context()
    for D in (1, 2, 3):
        gen_engine(D)
    for tag, params in EXTRA_SETS.items():
        for D in (1, 2):
            gen_engine(D, tag, params)
    print("golden vectors written to", HERE)
