"""Developer probe: joint5 with the Gauss law on the RF target vs the exponential law (per-kernel ms)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayhunter_b200 as bh
import bench

dev = torch.device("cuda:0")
for g in (False, True):
    print(json.dumps(bench.time_config(bh, torch, dev, "joint5", 8192, 5, 10, 3, gauss_rf=g)))
