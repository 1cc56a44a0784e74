import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hefl_b200 import _ext
from hefl_b200.he.context import CKKSContext
ops = _ext.ops()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
bits = (54, 54, 54, 55) if n >= 8192 else (36, 36, 37)
ctx = CKKSContext(n, prime_bits=bits, scale_bits=40, device="cuda")
C = int(sys.argv[2]) if len(sys.argv) > 2 else 512
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.stack([torch.randint(0, q, (C, 2, n), generator=g, device="cuda", dtype=torch.int64) for q in ctx.primes], dim=2).contiguous()
for _ in range(3):
    ops.ntt_(x, ctx.tables, ctx.consts, ctx.L, ctx.logn, False)
    ops.ntt_(x, ctx.tables, ctx.consts, ctx.L, ctx.logn, True)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); ops.ntt_(x, ctx.tables, ctx.consts, ctx.L, ctx.logn, False); e.record(); torch.cuda.synchronize()
print("ntt fwd ms", s.elapsed_time(e), "GB/s", 2 * x.numel() * 8 / s.elapsed_time(e) / 1e6)
