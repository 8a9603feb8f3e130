"""Experiment: do two C3 forwards in flight on two streams (two engines with their own workspaces, same weights) finish
sooner than two forwards back to back?   python tools/experiments/two_in_flight.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import isr_b200
from isr_b200.pipeline import FusionEngine
from oracle import fusion_oracle as O

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = isr_b200.CompleteEnhancedFusionSR(None).eval().to(dev)
m.precision = "bf16"
sets = []
for s in range(2):
    lr, imgs, fts, _ = O.synthetic_inputs(1, 339, 510)
    sets.append((lr.to(dev), [imgs[k].to(dev) for k in O.EXPERT_ORDER], {k: v.to(dev) for k, v in fts.items()}))
engs = [FusionEngine(m), FusionEngine(m)]
for e in engs:
    e.overlap_routing = False
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]

def run(n, two):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        k = i & 1 if two else 0
        st = streams[k] if two else torch.cuda.current_stream(dev)
        if two:
            st.wait_event(e0)
        with torch.cuda.stream(st):
            lr, il, ft = sets[k]
            engs[k].forward(lr, il, ft, 1356, 2040, False)
    if two:
        for st in streams:
            torch.cuda.current_stream(dev).wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for two in (False, True):
    run(4, two)
print("sequential  : %.3f ms per image" % run(20, False))
print("two streams : %.3f ms per image" % run(20, True))
print("sequential  : %.3f ms per image" % run(20, False))
