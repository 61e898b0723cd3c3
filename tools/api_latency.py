"""Host-side cost of one small absorptionCoefficient_Voigt call (config 1: 20 000 lines, 50 001 points, one state):
wall time per call and a cProfile breakdown of the Python / ctypes / torch path.   python tools/api_latency.py [calls]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radtxfr_b200 import hapi_compat, synthetic as syn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
cols = syn.config1_lines()
hapi_compat.storeTable("cfg1", cols)
grid = np.linspace(800.0, 1300.0, 50001)
env = {"T": 296.0, "p": 1.0}
f = lambda: hapi_compat.absorptionCoefficient_Voigt(SourceTables="cfg1", Environment=env, OmegaGrid=grid)
for _ in range(20):
    f()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    f()
dt = (time.perf_counter() - t0) / n
print("absorptionCoefficient_Voigt (numpy in, numpy out): %.3f ms per call" % (1e3 * dt))
g = torch.from_numpy(grid).cuda()
f2 = lambda: hapi_compat.absorptionCoefficient_Voigt(SourceTables="cfg1", Environment=env, OmegaGrid=g)
for _ in range(20):
    f2()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    f2()
torch.cuda.synchronize()
print("same with a CUDA tensor grid in, CUDA tensor out: %.3f ms per call" % (1e3 * (time.perf_counter() - t0) / n))
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    f()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative")
st.print_stats(28)
