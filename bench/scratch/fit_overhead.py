"""Where does a round's time go outside the CUDA graph? (device time per phase)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hefl_b200.config import FLConfig
from hefl_b200.fl import FederatedRunner
from hefl_b200.fl.data import ResidentFeeder

cfg = FLConfig(model="medcnn", nn_backend="tcgen05", transport="fused", device="cuda")
run = FederatedRunner(cfg, device=torch.device("cuda"))
tr = run.trainer
per = len(run.dataset); nval = cfg.val_steps * cfg.batch_size
feed = ResidentFeeder(run.dataset, range(nval, per), cfg.batch_size, run.device, seed=0)
vfeed = ResidentFeeder(run.dataset, range(0, nval), cfg.batch_size, run.device, seed=0)
tr.fit(feed, vfeed, 1, early_stopping=None)
torch.cuda.synchronize()

def dev_time(fn, n=1):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n, (time.perf_counter() - t0) * 1e3 / n

x, y = next(iter(feed.epoch()))
print("train graph replay only     : %.3f ms dev, %.3f ms wall" % dev_time(lambda: tr._graph_train.replay(), 230))
print("eval graph replay only      : %.3f ms dev, %.3f ms wall" % dev_time(lambda: tr._graph_eval.replay(), 30))
print("train_step(x,y) incl copies : %.3f ms dev, %.3f ms wall" % dev_time(lambda: tr.train_step(x, y), 230))
def feed_only():
    for _ in feed.epoch(): pass
print("feeder epoch only (23 steps): %.3f ms dev, %.3f ms wall" % dev_time(feed_only, 5))
print("fit 1 epoch (23+3 steps)    : %.3f ms dev, %.3f ms wall" % dev_time(lambda: tr.fit(feed, vfeed, 1, early_stopping=None), 5))
print("fit 10 epochs               : %.3f ms dev, %.3f ms wall" % dev_time(lambda: tr.fit(feed, vfeed, 10, early_stopping=None), 1))
