"""Kernel timeline (torch.profiler / CUPTI, analysis only) of the pipelined epoch runner: shows how the
pre-processing of batch i+1 on the side stream overlaps the graph of step i."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from hefl_b200.config import FLConfig
from hefl_b200.models import ParamPack, create_model
from hefl_b200.fl.trainer import LocalTrainer
from hefl_b200.fl.data import SyntheticImageDataset, ResidentFeeder

cfg = FLConfig(model="medcnn", batch_size=32, nn_backend="tcgen05")
dev = torch.device("cuda")
model = create_model("medcnn").to(dev)
pack = ParamPack(model)
tr = LocalTrainer(model, pack, cfg, dev, backend="tcgen05", use_graph=True)
ds = SyntheticImageDataset(256, 256, 3, 2, seed=0)
train = ResidentFeeder(ds, range(0, 256), 32, dev, seed=0)
stats = torch.zeros(train.steps, 2).pin_memory()
tr._run_epoch(train, True, stats)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr._run_epoch(train, True, stats)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
# print steps 3..4 (skip pipeline fill)
relay = [e for e in evs if "relayout" in e.name]
lo = relay[2].time_range.end if len(relay) > 4 else t0
hi = relay[4].time_range.end if len(relay) > 4 else evs[-1].time_range.end
for e in evs:
    if lo - 30 <= e.time_range.start <= hi:
        print(f"{e.time_range.start - lo:8.1f} {e.time_range.end - e.time_range.start:7.1f}  {e.name[:60]}")
print("total span per step:", (relay[-1].time_range.end - relay[0].time_range.end) / (len(relay) - 1))
