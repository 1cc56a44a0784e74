"""In-graph kernel timeline of one training step (torch.profiler / CUPTI; analysis only — numbers
reported anywhere else come from CUDA events without a profiler attached)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from hefl_b200.config import FLConfig
from hefl_b200.models import ParamPack, create_model
from hefl_b200.fl.trainer import LocalTrainer

cfg = FLConfig(model="medcnn", batch_size=32, nn_backend="tcgen05")
dev = torch.device("cuda")
model = create_model("medcnn").to(dev)
pack = ParamPack(model)
tr = LocalTrainer(model, pack, cfg, dev, backend="tcgen05", use_graph=True)
x = torch.randint(0, 256, (32, 256, 256, 3), dtype=torch.uint8, device="cuda")
y = torch.randint(0, 2, (32,), device="cuda")
for _ in range(5):
    tr.train_step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.train_step(x, y)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# keep the last replay: split on the H2D copies of static_x (Memcpy) — simply take the last third
n = len(evs) // 3
last = evs[-n:]
t0 = last[0].time_range.start
rows = []
for e in last:
    rows.append((e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:70]))
end = max(r[0] + r[1] for r in rows)
busy = sum(r[1] for r in rows)
print(f"{len(rows)} device activities, span {end:.1f} us, sum of durations {busy:.1f} us")
prev_end = 0.0
for st, du, nm in rows:
    gap = st - prev_end
    print(f"{st:8.1f} {du:7.1f}  gap {gap:6.1f}  {nm}")
    prev_end = max(prev_end, st + du)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/timeline.json", "w"))
