"""Micro-benchmark of the HE kernels on one GPU (CUDA events, warm-up, L2 flush between
iterations). Writes gpurun_out/he_micro_v2.json (persistent TMA/cluster kernels, the default) or,
with HEFL_HE_V1=1, gpurun_out/he_micro_v1.json (the first-generation kernels) for an A/B."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hefl_b200 import _ext  # noqa: E402
from hefl_b200.config import HE_PRESETS  # noqa: E402
from hefl_b200.he.context import CKKSContext  # noqa: E402

ops = _ext.ops()


def timeit(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    peaks = {}
    try:
        peaks = json.load(open("MEASURED_PEAKS.json"))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    out = {"hbm_gbs_ref": hbm, "results": []}
    for preset, nvals in [("n4096_l3", 222722), ("n8192_l4", 11_689_512), ("n16384_l4", 25_557_032)]:
        p = HE_PRESETS[preset]
        ctx = CKKSContext(p["n"], prime_bits=p["prime_bits"], scale_bits=p["scale_bits"], device="cuda")
        sk, pk = ctx.keygen(seed=1)
        vals = torch.randn(nvals, device="cuda") * 0.05
        C = ctx.num_ct(nvals)
        ct = ctx.encrypt(vals, pk, seed=3)
        ct_bytes = ct.nbytes()
        rows = C * 2 * ctx.L
        t_enc = timeit(lambda: ctx.encode(vals), flush=flush)
        msg = ctx.encode(vals)
        t_encrypt = timeit(lambda: ops.encrypt_out(msg, pk, C, ctx.L, ctx.logn, ctx.tables, ctx.consts, None, 5, 0, ct.data), flush=flush)
        t_ntt = timeit(lambda: ops.ntt_(ct.data, ctx.tables, ctx.consts, ctx.L, ctx.logn, False), flush=flush)
        t_intt = timeit(lambda: ops.ntt_(ct.data, ctx.tables, ctx.consts, ctx.L, ctx.logn, True), flush=flush)
        ct = ctx.encrypt(vals, pk, seed=3)
        t_dec = timeit(lambda: ctx.decrypt_residues(ct, sk), flush=flush)
        res = ctx.decrypt_residues(ct, sk)
        t_decode = timeit(lambda: ops.ckks_decode_residues(res, ctx.consts_cpu, ctx.q0_inv_q1, ctx.logn, 1.0 / ctx.scale, ctx.rot, ctx.ksi, nvals), flush=flush)
        t_mod = timeit(lambda: ops.reduce_mod_(ct.data, ctx.L, ctx.consts), flush=flush)
        r = dict(preset=preset, nvals=nvals, C=C, ct_mbytes=ct_bytes / 1e6,
                 encode_ms=t_enc, encrypt_ms=t_encrypt, ntt_ms=t_ntt, intt_ms=t_intt,
                 decrypt_ms=t_dec, decode_ms=t_decode, reduce_mod_ms=t_mod,
                 ntt_gbs=2 * ct_bytes / t_ntt / 1e6, ntt_frac_hbm=2 * ct_bytes / t_ntt / 1e6 / hbm,
                 encrypt_out_gbs=ct_bytes / t_encrypt / 1e6,
                 reduce_mod_gbs=2 * ct_bytes / t_mod / 1e6)
        print(json.dumps(r))
        out["results"].append(r)
        del ct, res, msg, vals
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    tag = "v1" if os.environ.get("HEFL_HE_V1") == "1" else "v2"
    out["kernels"] = tag
    json.dump(out, open(f"gpurun_out/he_micro_{tag}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
