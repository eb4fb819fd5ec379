#!/usr/bin/env python
"""K6 micro-benchmark: forward-only vs forward+backward, timed as 50 back-to-back launches inside one CUDA graph."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import ops  # noqa: E402


def graph_time(fn, n=50):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 5 / n * 1e3


def main():
    for B, H, C, L in ((50, 200, 42, 2), (20, 200, 42, 2), (148, 200, 42, 2), (50, 200, 42, 1)):
        pooled = torch.randn(B, 3 * H, device='cuda')
        labels = torch.randint(0, C, (B,), device='cuda')
        ws = [torch.randn(H, 3 * H if l == 0 else H, device='cuda') * 0.05 for l in range(L)]
        bs = [torch.zeros(H, device='cuda') for _ in range(L)]
        wc, bc = torch.randn(C, H, device='cuda') * 0.1, torch.zeros(C, device='cuda')
        buf = ops.HeadBuffers(B, H, C, L, 'cuda')
        dws = [torch.empty_like(w) for w in ws]
        dbs = [torch.empty_like(b) for b in bs]
        dwc, dbc = torch.empty_like(wc), torch.empty_like(bc)
        t_f = graph_time(lambda: ops.head_fwd_bwd(pooled, labels, ws, bs, wc, bc, 0.003, buf, train=False))
        t_fb = graph_time(lambda: ops.head_fwd_bwd(pooled, labels, ws, bs, wc, bc, 0.003, buf, train=True))
        t_w = graph_time(lambda: ops.head_wgrad(pooled, buf, dws, dbs, dwc, dbc))
        print('B=%d H=%d C=%d L=%d: fwd %.1f us, fwd+bwd %.1f us, wgrad %.1f us' % (B, H, C, L, t_f, t_fb, t_w))


if __name__ == '__main__':
    main()
