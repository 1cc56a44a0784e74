import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from hefl_b200.config import FLConfig
from hefl_b200.models import ParamPack, create_model
from hefl_b200.ops.conv_engine import MedCNNEngine
torch.manual_seed(0)
B = 8
cfg = FLConfig(model="medcnn", batch_size=B, image_size=256)
dev = torch.device("cuda")
model = create_model("medcnn").to(dev)
pack = ParamPack(model)
pack.flat.copy_(pack.flat.to(torch.bfloat16).float())
eng = MedCNNEngine(model, pack, cfg, dev)
g = torch.Generator(device="cuda").manual_seed(7)
x = torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, device="cuda", generator=g)
y = torch.randint(0, 2, (B,), device="cuda", generator=g)
out = torch.zeros(2, device="cuda")
eng.train_step(x, y, out, augment=False)
torch.cuda.synchronize()
g_eng = pack.grad.clone(); pack.grad.zero_()
class RoundBF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t): return t.to(torch.bfloat16).float()
    @staticmethod
    def backward(ctx, gr): return gr.to(torch.bfloat16).float()
xr = (x.float() / 255.0).to(torch.bfloat16).float().permute(0, 3, 1, 2)
h = xr
acts = []
for conv in model.convs:
    h = RoundBF.apply(F.max_pool2d(F.relu(conv(h)), 2)); acts.append(h)
h = h.permute(0, 2, 3, 1).flatten(1)
for fc in model.fcs[:-1]: h = F.relu(fc(h))
logits = model.fcs[-1](h)
loss = F.cross_entropy(logits, y); loss.backward()
g_ref = pack.grad.clone()
print("loss eng", float(out[0]), "ref", float(loss))
for l, a in enumerate(acts):
    e = eng.X[l + 1].view(B, a.shape[2], a.shape[3], a.shape[1]).permute(0, 3, 1, 2).float()
    print("act", l, "max abs diff", float((e - a).abs().max()), "ref max", float(a.abs().max()))
for key, shape, off, n in pack.entries:
    a, b = g_eng[off:off + n], g_ref[off:off + n]
    rel = float((a - b).abs().max() / (b.abs().max() + 1e-12))
    cos = float(F.cosine_similarity(a, b, dim=0))
    print(key, tuple(shape), "rel", round(rel, 4), "cos", round(cos, 5), "|ref|", float(b.norm()))
