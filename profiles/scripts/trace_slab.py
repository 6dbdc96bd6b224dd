"""Kernel timeline (CUPTI through torch.profiler) of the fused-halo explicit Euler loop on slabs, rank 0: start and
duration of every kernel of a few steps, so that the step's critical path (interior kernel, wait / end warps / signal
chain, graph node gaps) can be read off.  Launch: torchrun --nproc-per-node N profiles/scripts/trace_slab.py [nx]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench_slab  # noqa: E402
import splitfxm_b200 as sf  # noqa: E402
from splitfxm_b200.distributed import SlabSimulation  # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
overlap = os.environ.get("SLAB_OVERLAP", "1") != "0"
x = np.linspace(0.0, 1.0, nx)
slab = SlabSimulation(x, bench_slab.COMPS, "euler_cons", {"gamma": 1.4}, bench_slab.BCS, sf.FVSchemes.LAX_FRIEDRICHS, "FVM",
                      rank, world, device=dev)
y = slab.eng.to_device(bench_slab.sod_state(x[slab.lo:slab.hi]).reshape(-1))
dt = 0.4 * (x[1] - x[0]) / 2.2
slab.enable_peer_halos()
y, _ = slab.evolve_euler_fused(y, 20 * dt * 0.999999, dt, overlap=overlap)
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    y, _ = slab.evolve_euler_fused(y, 24 * dt * 0.999999, dt, overlap=overlap)
    torch.cuda.synchronize()
trace_rank = int(os.environ.get("TRACE_RANK", "0"))
if rank == trace_rank:
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            ev.append((e.time_range.start, e.time_range.end - e.time_range.start, e.name[:48]))
    ev.sort()
    t0 = ev[0][0]
    out = [{"t_us": round(s - t0, 1), "dur_us": round(d, 1), "name": n} for s, d, n in ev]
    tag = "overlap" if overlap else "serial"
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"trace_slab_n{world}_r{rank}_{tag}.json"), "w"))
    big = [k for k in out if "residual_warp" in k["name"] and k["dur_us"] > 15]
    print(tag, "kernels", len(out), "interior launches", len(big))
    for k in out[-44:]:
        print(k["t_us"], k["dur_us"], k["name"][:34])
dist.barrier()
