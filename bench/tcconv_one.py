"""One 3x3 and one 1x1 convolution, forward + backward, on the tcgen05 path (for ncu launch lists / captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hefl_b200.ops import tc_conv

def run(B, Cin, Cout, H, k):
    x = torch.randn(B, Cin, H, H, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
    gy = torch.randn(B, Cout, H, H, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        xa = x.detach().requires_grad_(True); wa = w.detach().requires_grad_(True)
        y = tc_conv.conv1x1(xa, wa) if k == 1 else tc_conv.conv3x3(xa, wa)
        y.backward(gy)
    torch.cuda.synchronize()

run(32, 128, 128, 28, 3)
run(32, 64, 64, 56, 3)
run(32, 512, 128, 28, 1)
# block-scaled e4m3 1x1 (forward + dgrad)
x = torch.randn(32, 512, 28, 28, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = torch.randn(128, 512, 1, 1, device="cuda") * 0.05
gy = torch.randn(32, 128, 28, 28, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
for _ in range(3):
    xa = x.detach().requires_grad_(True); wa = w.detach().requires_grad_(True)
    tc_conv.conv1x1(xa, wa, 1, "mx").backward(gy)
torch.cuda.synchronize()
print("ok")
