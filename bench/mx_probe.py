"""Probe of the block-scaled e4m3 GEMM (kind::mxf8f6f4.block_scale): python bench/mx_probe.py <idesc_variant> <layout>
Quantises random matrices to e4m3 with one UE8M0 scale per row and 32 elements of K, tiles the scales under a layout
hypothesis, runs the kernel and compares with the de-quantised fp32 product (exact up to the bf16 output rounding).
layout 0: tile[l][i][kb] (row 32*i + l), 1: tile[l][kb][i], 2: tile[row = 0..127][kb] (plain row-major)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hefl_b200 import _ext

ops = _ext.ops()
variant, layout = int(sys.argv[1]), int(sys.argv[2])
M, N, K = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (256, 256, 256)
torch.manual_seed(0)
dev = "cuda"


def quant(x):
    R, Kk = x.shape
    xb = x.view(R, Kk // 32, 32)
    amax = xb.abs().amax(-1).clamp_min(1e-30)
    e = torch.ceil(torch.log2(amax / 448.0)).clamp(-127, 127)
    q = (xb / torch.exp2(e).unsqueeze(-1)).to(torch.float8_e4m3fn)
    deq = (q.float() * torch.exp2(e).unsqueeze(-1)).view(R, Kk)
    return q.view(R, Kk).view(torch.uint8).contiguous(), (e + 127).to(torch.uint8), deq


def tile(sf, layout):
    R, KB = sf.shape
    Rp = (R + 127) // 128 * 128
    full = torch.full((Rp, KB), 127, dtype=torch.uint8, device=sf.device)
    full[:R] = sf
    t = full.view(Rp // 128, 4, 32, KB // 4, 4)            # [mb][i][l][kg][kb]
    if layout == 0:
        t = t.permute(0, 3, 2, 1, 4)                       # [mb][kg][l][i][kb]
    elif layout == 1:
        t = t.permute(0, 3, 2, 4, 1)                       # [mb][kg][l][kb][i]
    else:
        t = t.permute(0, 3, 1, 2, 4)                       # [mb][kg][i][l][kb] = row-major rows
    return t.contiguous().view(-1)


a = torch.randn(M, K, device=dev) * torch.exp2((torch.arange(M, device=dev) % 7).float()).unsqueeze(1)
a = a * torch.exp2(((torch.arange(K, device=dev) // 32) % 3).float()).unsqueeze(0)
b = torch.randn(N, K, device=dev) * torch.exp2((torch.arange(N, device=dev) % 5).float()).unsqueeze(1)
qa, sa, da = quant(a)
qb, sb, db = quant(b)
ref = da @ db.t()
out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
ops.gemm_mxfp8(qa, qb, tile(sa, layout), tile(sb, layout), out, variant)
torch.cuda.synchronize()
o = out.float()
rel = float((o - ref).abs().max() / ref.abs().max())
ratio = (o.abs().mean(1) / ref.abs().mean(1)).log2()
print(f"variant {variant} layout {layout} M{M} N{N} K{K}: rel err {rel:.4g}; log2(row scale out/ref) first 12 rows:",
      [round(float(v), 2) for v in ratio[:12]], " rows 32..36:", [round(float(v), 2) for v in ratio[32:36]])
