"""Tiny driver for ncu captures of the HE kernels: python bench/he_prof.py <preset> [n_ct]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hefl_b200 import _ext  # noqa: E402
from hefl_b200.config import HE_PRESETS  # noqa: E402
from hefl_b200.he.context import CKKSContext  # noqa: E402

ops = _ext.ops()
preset = sys.argv[1] if len(sys.argv) > 1 else "n8192_l4"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 600
p = HE_PRESETS[preset]
ctx = CKKSContext(p["n"], prime_bits=p["prime_bits"], scale_bits=p["scale_bits"], device="cuda")
sk, pk = ctx.keygen(seed=1)
vals = torch.randn(C * ctx.slots, device="cuda") * 0.05
for _ in range(3):
    ct = ctx.encrypt(vals, pk, seed=3)
    res = ctx.decrypt_residues(ct, sk)
    ops.ntt_(ct.data, ctx.tables, ctx.consts, ctx.L, ctx.logn, True)
    ops.ntt_(ct.data, ctx.tables, ctx.consts, ctx.L, ctx.logn, False)
torch.cuda.synchronize()
print("ok", preset, C)
