import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hefl_b200 import _ext
ops = _ext.ops()
B, H, Ci, CK, Co = [int(v) for v in sys.argv[1:6]]
g = torch.Generator(device="cuda").manual_seed(3)
x = torch.zeros(B, H, H, CK, device="cuda"); x[..., :Ci] = torch.randn(B, H, H, Ci, device="cuda", generator=g)
x = x.to(torch.bfloat16)
Ho = H - 2
dyv = torch.randn(B, Co, Ho, Ho, device="cuda", generator=g).to(torch.bfloat16)
dY = torch.zeros(B, H, H, Co, dtype=torch.bfloat16, device="cuda"); dY[:, :Ho, :Ho, :] = dyv.permute(0, 2, 3, 1)
P = B * H * H
dW32 = torch.zeros((9 * CK + 1) * Co, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
ops.conv_wgrad(x.view(-1, CK), dY.view(-1, Co), dW32, B, H, H, CK, Co)
torch.cuda.synchronize()
wref = torch.nn.grad.conv2d_weight(x[..., :Ci].float().permute(0, 3, 1, 2), (Co, Ci, 3, 3), dyv.float())
got = dW32[: 9 * CK * Co].view(9, CK, Co)[:, :Ci, :].permute(2, 1, 0).reshape(Co, Ci, 3, 3)
bref = dyv.float().sum((0, 2, 3)); print("bias err", float((dW32[9*CK*Co:] - bref).abs().max()), float(bref.abs().max()))
print("OK max err", float((got - wref).abs().max()), "ref max", float(wref.abs().max()))
