"""Layer-1 forward (conv + bias + ReLU + 2x2 max-pool + arg-max) on the s-packed input: tap-GEMM kernel vs pair-row
kernel, batch 32 at 256x256, CUDA events, L2 flushed between iterations."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hefl_b200 import _ext

ops = _ext.ops()
B, H, Co = 32, 256, 32
Hp = (H - 2) // 2
P = B * H * H
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randint(0, 256, (B, H, H, 3), dtype=torch.uint8, device="cuda", generator=g)
X = torch.zeros(P + 8, 16, dtype=torch.bfloat16, device="cuda")
ops.preprocess_u8(x, None, X[:P], 0, None, True)
Wp = (torch.randn(3, Co, 16, device="cuda", generator=g) * 0.2).to(torch.bfloat16)
bias = torch.randn(Co, device="cuda", generator=g) * 0.1
out = torch.zeros(B * Hp * Hp, Co, dtype=torch.bfloat16, device="cuda")
am = torch.zeros(B * Hp * Hp, Co, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for name in ("tap_gemm", "pair_row"):
    ts = []
    for it in range(25):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if name == "tap_gemm":
            ops.conv_fwd_pool(X[:P], Wp.view(-1), bias, out, am, B, H, H, 16, Co, True)
        else:
            ops.conv_fwd_pool_pair(X, Wp.view(-1), bias, out, am, B, H, H, 16, Co, True)
        b.record()
        torch.cuda.synchronize()
        if it >= 5:
            ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    res[name] = {"us_median": ts[len(ts) // 2], "us_min": ts[0]}
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/fwd0_micro.json", "w"))
