"""Where does a round's time go beyond (steps x step time)? Times fit() of the medical CNN for E epochs
with CUDA events and host timestamps around the epoch boundaries."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hefl_b200.config import FLConfig
from hefl_b200.models import ParamPack, create_model
from hefl_b200.fl.trainer import LocalTrainer
from hefl_b200.fl.data import SyntheticImageDataset, ResidentFeeder

cfg = FLConfig(model="medcnn", batch_size=32, nn_backend="tcgen05")
dev = torch.device("cuda")
model = create_model("medcnn").to(dev)
pack = ParamPack(model)
tr = LocalTrainer(model, pack, cfg, dev, backend="tcgen05", use_graph=True)
ds = SyntheticImageDataset(816, 256, 3, 2, seed=0)
train = ResidentFeeder(ds, range(80, 816), 32, dev, seed=0)
val = ResidentFeeder(ds, range(0, 80), 32, dev, seed=0)
tr.fit(train, val, 2, early_stopping=None)
torch.cuda.synchronize()
for E in (1, 10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record(); tr.fit(train, val, E, early_stopping=None); e.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"epochs={E}: device {s.elapsed_time(e):.2f} ms, host {1e3*(t1-t0):.2f} ms, per epoch {s.elapsed_time(e)/E:.3f} ms "
          f"(train steps {train.steps}, val steps {val.steps})")
# pure back-to-back steps without epoch boundaries
x, y = next(iter(train.epoch()))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import torch as _t
stats = _t.zeros(train.steps, 2).pin_memory()
s.record(); tr._run_epoch(train, True, stats); e.record(); torch.cuda.synchronize()
print(f"one training pass ({train.steps} steps): {s.elapsed_time(e):.3f} ms = {1e3*s.elapsed_time(e)/train.steps:.1f} us/step")
stats = _t.zeros(val.steps, 2).pin_memory()
s.record(); tr._run_epoch(val, False, stats); e.record(); torch.cuda.synchronize()
print(f"one validation pass ({val.steps} steps): {s.elapsed_time(e):.3f} ms = {1e3*s.elapsed_time(e)/val.steps:.1f} us/step")
# host cost of one pass with the GPU idle at the end
t0 = time.perf_counter(); tr._run_epoch(train, True, _t.zeros(train.steps, 2).pin_memory()); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue time of a training pass {1e3*(t1-t0):.2f} ms, then {1e3*(t2-t1):.2f} ms until the GPU drains")
