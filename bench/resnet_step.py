"""One ResNet training step (forward + backward, batch 32 at 224 x 224, bf16 autocast) on the tcgen05 convolutions or
the cuDNN arm -- for ncu launch lists:  python bench/resnet_step.py resnet18 tc|cudnn [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from hefl_b200.models import create_model
from hefl_b200.ops import tc_conv

name = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
on = (sys.argv[2] if len(sys.argv) > 2 else "tc") == "tc"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.backends.cudnn.benchmark = False      # no autotuning launches in the list: every step has the same launches
m = create_model(name, num_classes=1000).cuda().train()
tc_conv.set_model_tc(m, on)
x = torch.randn(32, 224, 224, 3, device="cuda").permute(0, 3, 1, 2)
y = torch.randint(0, 1000, (32,), device="cuda")
for _ in range(steps):
    for p in m.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = F.cross_entropy(m(x).float(), y)
    loss.backward()
torch.cuda.synchronize()
print("ok", name, on, float(loss))
