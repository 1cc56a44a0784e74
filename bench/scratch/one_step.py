"""Runs a few eager training steps of the medical CNN on the tcgen05 engine (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hefl_b200.config import FLConfig
from hefl_b200.models import ParamPack, create_model
from hefl_b200.fl.trainer import LocalTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = FLConfig(model="medcnn", batch_size=32, nn_backend="tcgen05")
dev = torch.device("cuda")
model = create_model("medcnn").to(dev)
pack = ParamPack(model)
tr = LocalTrainer(model, pack, cfg, dev, backend="tcgen05", use_graph=False)
x = torch.randint(0, 256, (32, 256, 256, 3), dtype=torch.uint8, device="cuda")
y = torch.randint(0, 2, (32,), device="cuda")
for _ in range(steps):
    tr.train_step(x, y)
torch.cuda.synchronize()
print("done", tr.out_train.tolist())
