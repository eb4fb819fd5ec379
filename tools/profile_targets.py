#!/usr/bin/env python
"""Launch every hot kernel of the path a few times at the large synthetic shape (BASELINE.json configs[4]) so that one
`ncu --set full -k regex:...` capture holds them all:  K1 prune_csr, K3 projection (tcgen05 TF32 and 3xTF32), K2
aggregate forward (dropout + activation mask, as the training step runs it), K2 backward (pre-scaled), K4 pool3."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=4096)
    ap.add_argument('--T', type=int, default=512)
    ap.add_argument('--H', type=int, default=512)
    ap.add_argument('--k', type=int, default=-1)
    ap.add_argument('--reps', type=int, default=2)
    a = ap.parse_args()
    B, T, H = a.B, a.T, a.H
    batch = synth.make_batch_torch(7, B, T, device='cuda')
    rng = torch.tensor([1, 1], dtype=torch.int64, device='cuda')
    x = torch.randn(B * T, 360, device='cuda')
    w = torch.randn(H, 360, device='cuda') * 0.05
    bias = torch.zeros(H, device='cuda')
    for _ in range(a.reps):
        csr = ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], a.k)
        for mode in ('tf32', 'tf32x3'):
            ws = ops.weight_prep_buffer(w, mode)
            ops.weight_prep(w, mode, out=ws)
            y = ops.linear_fwd(x, w, mode, ws)
        out, act = ops.aggregate_fwd(y, csr, bias, drop_p=0.5, rng_state=rng, want_act=True)
        g = torch.randn(B, T, H, device='cuda')
        db = torch.zeros(H, device='cuda')
        dy = ops.aggregate_bwd_pre(g, csr, dbias_out=db)
        pooled, argmax = ops.pool3_fwd(out, csr, 0)
        del y, out, act, g, dy
    torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
