"""Kernel timeline (CUPTI through torch.profiler) of the ensemble steady state: per call, every kernel's start and
duration, so that idle gaps between kernels - host-side latency - can be told from kernel time."""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import splitfxm_b200 as sf  # noqa: E402
from splitfxm_b200.ensemble import EnsembleSimulation  # noqa: E402

dev = torch.device("cuda:0")
B = 65536
x, y0 = bench.initial_state()
c, nu = bench.member_params(0, B, B)
d = sf.Domain.from_size(bench.NX, 1, 1, bench.COMPS)
ens = EnsembleSimulation(d, "ad", {"c": c, "nu": nu}, bench.BCS, sf.FDSchemes.CENTRAL, method="FDM", device=dev)
Y = torch.from_numpy(bench.member_states(y0, 0, B)).to(dev)
for _ in range(3):
    ens.steady_state(Y, split=True, split_locs=[1])
torch.cuda.synchronize()
times = []
for _ in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ens.steady_state(Y, split=True, split_locs=[1])
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
print("per-call ms:", [round(t, 3) for t in times])
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        ens.steady_state(Y, split=True, split_locs=[1])
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        ev.append((e.time_range.start, e.time_range.end - e.time_range.start, e.name[:60]))
ev.sort()
t0 = ev[0][0]
out = [{"t_us": round(s - t0, 1), "dur_us": round(dur, 1), "name": n} for s, dur, n in ev]
json.dump({"per_call_ms": times, "kernels": out}, open(os.path.join(ROOT, "gpurun_out", "trace_newton.json"), "w"))
busy = sum(k["dur_us"] for k in out)
print("kernels", len(out), "busy ms", busy / 1e3, "span ms", (out[-1]["t_us"] + out[-1]["dur_us"]) / 1e3)
