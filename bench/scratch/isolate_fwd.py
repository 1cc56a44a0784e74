"""Bottleneck isolation for the layer-1/2 forward kernel: skip TMA / MMA / epilogue in turn."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hefl_b200 import _ext
ops = _ext.ops()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[n // 2]
for (B, H, CK, CO) in [(32, 256, 16, 32), (32, 127, 32, 32)]:
    X = torch.randn(B * H * H, CK, device="cuda").to(torch.bfloat16)
    Wf = torch.randn(9, CO, CK, device="cuda").to(torch.bfloat16)
    bias = torch.zeros(CO, device="cuda")
    Hp = (H - 2) // 2
    out = torch.zeros(B * Hp * Hp, CO, dtype=torch.bfloat16, device="cuda")
    am = torch.zeros(B * Hp * Hp, CO, dtype=torch.uint8, device="cuda")
    for mask, name in [(0, "full"), (1, "no TMA"), (2, "no MMA"), (4, "no epilogue"), (3, "no TMA+MMA"), (5, "no TMA+epi"), (6, "no MMA+epi"), (7, "nothing")]:
        ops.conv_set_debug(mask)
        print(f"H={H} CK={CK}: {name:12s} {t(lambda: ops.conv_fwd_pool(X, Wf, bias, out, am, B, H, H, CK, CO)):7.1f} us")
    ops.conv_set_debug(0)
