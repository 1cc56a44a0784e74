"""Layer-1 weight gradient from the pooled gradient: FP32-pipe gather (wgrad_gather.cu) vs masked GEMMs on the
tensor cores (wgrad0_mma.cu), batch 32 at 256x256, CUDA events, L2 flushed between iterations."""
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
gp = (torch.randn(B * Hp * Hp, Co, device="cuda", generator=g)).to(torch.bfloat16)
amax = torch.randint(0, 8, (B * Hp * Hp, Co), dtype=torch.uint8, device="cuda", generator=g)
dW = torch.zeros(145 * 32, dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for name, sp in (("gather", False), ("masked_gemm", True)):
    ts = []
    for it in range(25):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.wgrad0_gather(X, gp, amax, dW, B, H, H, sp)
        b.record()
        torch.cuda.synchronize()
        if it >= 5:
            ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    res[name] = {"us_median": ts[len(ts) // 2], "us_min": ts[0]}
res["bytes_MB"] = (P * 32 + B * Hp * Hp * Co * 3) / 1e6
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/wgrad0_micro.json", "w"))
