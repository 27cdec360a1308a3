#!/usr/bin/env python
"""DRAM traffic of K consecutive step launches measured as ONE ncu range, so that the write-back of the outputs
(which leaves L2 after the kernel that produced them has ended and is therefore missing from a per-kernel
capture) is inside the measurement.  Every launch of the range has its own input and output buffers.

On the GPU box:
    ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/range_<scenario>.csv python tools/ncu_range_traffic.py <scenario> <batch> <K>
prints the algorithmic bytes of the K launches; the csv holds the measured ones."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casclik_b200 import scenarios  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ur5_track"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
K = int(sys.argv[3]) if len(sys.argv) > 3 else 8
sc = scenarios.get(name)
ctrl = sc.make_controller()
ctrl.setup_solver()
dev = torch.device("cuda", 0)
ins, outs = [], []
for s in range(K + 1):
    inp = sc.sample(B, seed=300 + s)
    ins.append(tuple(None if inp[k] is None else torch.from_numpy(np.ascontiguousarray(inp[k])).to(dev) for k in ("t", "q", "x", "y")))
    outs.append(ctrl.solve_batch(*ins[-1]))            # allocates the outputs of set s (and warms up)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for s in range(K):
    ctrl.solve_batch(*ins[s], out=outs[s])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
meta = ctrl.kernel_meta
per = meta["qp_bytes_per_step"] if sc.controller == "qp" else meta["pinv_bytes_per_step"]
print(json.dumps({"scenario": name, "batch": B, "launches_in_range": K, "algorithmic_bytes": per * B * K,
                  "algorithmic_bytes_per_launch": per * B}))
