"""tcgen05 GEMM convolutions vs cuDNN / cuBLAS on the ResNet shapes (batch 32 at 224 x 224): per-layer forward,
dgrad, wgrad times with achieved TFLOP/s, and whole ResNet-18 / ResNet-50 training steps (forward + backward under a
CUDA graph) with the hand-written convolutions on and off. CUDA events, warm-up, L2 flushed between iterations.
Writes gpurun_out/tcconv_micro.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from hefl_b200.models import create_model
from hefl_b200.ops import fp8, tc_conv

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
torch.backends.cudnn.benchmark = True


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


def layer(B, Cin, Cout, H, k):
    x = torch.randn(B, Cin, H, H, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
    wb = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(B, Cout, H, H, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    flops = 2.0 * B * H * H * Cin * Cout * k * k
    out = {"B": B, "Cin": Cin, "Cout": Cout, "H": H, "k": k, "gflop": flops / 1e9}

    def tc_fb():
        xa = x.detach().requires_grad_(True); wa = w.detach().requires_grad_(True)
        y = tc_conv.conv1x1(xa, wa) if k == 1 else tc_conv.conv3x3(xa, wa)
        y.backward(gy)

    def lib_fb():
        xa = x.detach().requires_grad_(True); wa = wb.detach().requires_grad_(True)
        y = F.conv2d(xa, wa, padding=k // 2)
        y.backward(gy)

    def tc_f():
        with torch.no_grad():
            return tc_conv.conv1x1(x, w) if k == 1 else tc_conv.conv3x3(x, w)

    def lib_f():
        with torch.no_grad():
            return F.conv2d(x, wb, padding=k // 2)

    for name, fn, mult in (("tc_fwd", tc_f, 1), ("lib_fwd", lib_f, 1), ("tc_fwd_bwd", tc_fb, 3), ("lib_fwd_bwd", lib_fb, 3)):
        t = timeit(fn)
        out[name + "_us"] = t
        out[name + "_tflops"] = mult * flops / t / 1e6
    if k == 1 and Cin % 128 == 0 and Cout % 128 == 0:
        def tc_mx():
            with torch.no_grad():
                return tc_conv.conv1x1(x, w, 1, "mx")
        out["tc_fwd_mxfp8_us"] = timeit(tc_mx)
    return out


def model_step(name):
    m = create_model(name, num_classes=1000).cuda().train()
    x = torch.randn(32, 224, 224, 3, device="cuda").permute(0, 3, 1, 2)
    y = torch.randint(0, 1000, (32,), device="cuda")

    def step():
        for p in m.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(m(x).float(), y)
        loss.backward()

    res = {}
    for on in (True, False):
        tc_conv.set_model_tc(m, on)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        res["tcgen05_us" if on else "cudnn_us"] = timeit(lambda: g.replay(), iters=8)
        del g
    res["ratio_cudnn_over_tcgen05"] = res["cudnn_us"] / res["tcgen05_us"]
    if name == "resnet50":                              # 1x1 convolutions forward + dgrad in block-scaled e4m3
        tc_conv.set_model_tc(m, True)
        fp8.set_model_fp8(m, True)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        res["tcgen05_mxfp8_us"] = timeit(lambda: g.replay(), iters=8)
        fp8.set_model_fp8(m, False)
    return res


def main():
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
    out = {"bf16_tflops_ref": peaks.get("bf16_tflops"), "layers": [], "models": {}}
    for cfg in [(32, 64, 64, 56, 3), (32, 128, 128, 28, 3), (32, 256, 256, 14, 3), (32, 512, 512, 7, 3),
                (32, 64, 256, 56, 1), (32, 256, 64, 56, 1), (32, 512, 128, 28, 1), (32, 1024, 256, 14, 1),
                (32, 512, 2048, 7, 1), (32, 2048, 512, 7, 1)]:
        r = layer(*cfg)
        print(json.dumps(r), flush=True)
        out["layers"].append(r)
    for name in ("resnet18", "resnet50"):
        out["models"][name] = model_step(name)
        print(name, json.dumps(out["models"][name]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/tcconv_micro.json", "w"), indent=1)


if __name__ == "__main__":
    main()
