#!/usr/bin/env python
"""The large-shape projection GEMMs, two rounds of [TF32, 3xTF32] forward (+ the tcgen05 weight gradient), for an ncu
capture:  ncu --set full -k regex:'gemm_persistent|wgrad_tf32x3' -s 3 -c 3 ...  (the first round is warm-up)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--M', type=int, default=2097152)
    ap.add_argument('--N', type=int, default=512)
    ap.add_argument('--K', type=int, default=360)
    ap.add_argument('--cg', type=int, default=2)
    a = ap.parse_args()
    ops.gemm_persist_config(a.cg, 65536)
    x = torch.randn(a.M, a.K, device='cuda')
    w = torch.randn(a.N, a.K, device='cuda') * 0.05
    dy = torch.randn(a.M, a.N, device='cuda')
    dw = torch.zeros(a.N, a.K, device='cuda')
    flags = torch.ones(a.M, dtype=torch.uint8, device='cuda')
    for _ in range(2):
        for mode in ('tf32', 'tf32x3'):
            ws = ops.weight_prep(w, mode)
            y = ops.linear_fwd(x, w, mode, ws)
        ops.linear_wgrad(dy, x, 'tf32x3', out=dw, accumulate=True, flags=flags)
    torch.cuda.synchronize()
    print('done', float(y[0, 0]))


if __name__ == '__main__':
    main()
