"""Fused BN(+residual)+ReLU kernels vs ATen/cuDNN BN + add + ReLU (fwd+bwd), and a ResNet-18
training step with the fused path on/off. CUDA events, warm-up, L2 flushed between iterations.
Writes gpurun_out/resnet_micro.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torch.nn.functional as F
from hefl_b200.models import create_model
from hefl_b200.ops import resnet_ops

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    res = {}
    for (B, C, H) in [(32, 64, 64), (32, 128, 32), (32, 256, 16), (32, 512, 8)]:
        x = torch.randn(B, C, H, H, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        r = torch.randn_like(x)
        g = torch.randn_like(x)
        bn = nn.BatchNorm2d(C).cuda()
        def run():
            xa = x.detach().requires_grad_(True); ra = r.detach().requires_grad_(True)
            y = resnet_ops.bn_act(bn, xa, ra, True)
            y.backward(g)
        resnet_ops.ENABLE = True
        t_f = timeit(run)
        resnet_ops.ENABLE = False
        t_a = timeit(run)
        resnet_ops.ENABLE = True
        nbytes = x.numel() * 2
        res[f"bn_add_relu_fwd_bwd_B{B}_C{C}_H{H}"] = {"fused_us": t_f, "aten_us": t_a, "speedup": t_a / t_f,
                                                     "fused_GBps": 9 * nbytes / t_f / 1e3}
    m = create_model("resnet18", num_classes=10).cuda().train()
    x = torch.randn(32, 224, 224, 3, device="cuda").permute(0, 3, 1, 2)
    y = torch.randint(0, 10, (32,), device="cuda")
    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(m(x).float(), y)
        loss.backward()
    def graphed(enable):
        resnet_ops.ENABLE = enable
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        t = timeit(lambda: g.replay(), iters=8)
        resnet_ops.ENABLE = True
        return t
    resnet_ops.ENABLE = True
    t_f = timeit(step, iters=5)
    resnet_ops.ENABLE = False
    t_a = timeit(step, iters=5)
    resnet_ops.ENABLE = True
    res["resnet18_b32_224_fwd_bwd_eager"] = {"fused_us": t_f, "aten_us": t_a, "speedup": t_a / t_f}
    t_f, t_a = graphed(True), graphed(False)
    res["resnet18_b32_224_fwd_bwd_cudagraph"] = {"fused_us": t_f, "aten_us": t_a, "speedup": t_a / t_f}
    for k, v in res.items():
        print(k, json.dumps(v))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/resnet_micro.json", "w"), indent=1)


if __name__ == "__main__":
    main()
