#!/usr/bin/env python
"""ncu target: K2 forward exactly as the training step launches it (in-kernel dropout 0.5 + activation bit mask) and K2
backward from that mask, twice each, at the large synthetic shape (BASELINE.json configs[4]).
  ncu --set full --clock-control none --import-source on -k regex:aggregate_ --launch-skip 2 --launch-count 2 \
      -o gpurun_out/k2 python tools/k2_ncu_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import ops, synth  # noqa: E402

B, T, H = 4096, 512, 512
batch = synth.make_batch_torch(7, B, T, device='cuda')
csr = ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], -1)
y = torch.randn(B * T, H, device='cuda')
bias = torch.zeros(H, device='cuda')
rng = torch.tensor([1, 1], dtype=torch.int64, device='cuda')
gout = torch.randn(B, T, H, device='cuda')
for _ in range(2):
    out, act = ops.aggregate_fwd(y, csr, bias, drop_p=0.5, rng_state=rng, want_act=True)
    del out
    dy, db = ops.aggregate_bwd(gout, None, csr, drop_p=0.5, act=act)
    del dy
torch.cuda.synchronize()
print('done')
