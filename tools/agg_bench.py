#!/usr/bin/env python
"""Micro-benchmark of K2 (aggregation fwd/bwd) and K1/K4 at the large synthetic shape; prints GB/s vs measured peak."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import ops, synth  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=4096)
    ap.add_argument('--T', type=int, default=512)
    ap.add_argument('--H', type=int, default=512)
    ap.add_argument('--k', type=int, default=-1)
    ap.add_argument('--vec', type=int, default=0)
    a = ap.parse_args()
    peak = 6551.0
    if os.path.exists('MEASURED_PEAKS.json'):
        peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
    B, T, H = a.B, a.T, a.H
    batch = synth.make_batch_torch(7, B, T, device='cuda')
    t_k1 = timeit(lambda: ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], a.k))
    csr = ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], a.k)
    nnz = int(csr.rowptr[:, T].sum())
    rows = int((csr.flags != 0).sum())
    k1_bytes = B * T * 33 + 4 * B * (T + 1) + 5 * nnz + 5 * B * T + 8 * B
    print('K1 prune_csr   %.3f ms  %.0f GB/s (%.1f%% of %.0f)  nnz=%d rows=%d' %
          (t_k1, k1_bytes / t_k1 / 1e6, 100 * k1_bytes / t_k1 / 1e6 / peak, peak, nnz, rows))
    y = torch.randn(B * T, H, device='cuda')
    bias = torch.zeros(H, device='cuda')
    rng = torch.tensor([1, 1], dtype=torch.int64, device='cuda')
    fwd_bytes = 2 * B * T * H * 4 + 4 * B * (T + 1) + 4 * nnz + 5 * B * T
    for name, p in (('fwd', 0.0), ('fwd+dropout', 0.5)):
        t = timeit(lambda: ops.aggregate_fwd(y, csr, bias, drop_p=p, rng_state=rng, force_vec=a.vec))
        print('K2 %-12s %.3f ms  %.0f GB/s (%.1f%%)' % (name, t, fwd_bytes / t / 1e6, 100 * fwd_bytes / t / 1e6 / peak))
    t = timeit(lambda: ops.aggregate_fwd(y, csr, bias, drop_p=0.5, rng_state=rng, force_vec=a.vec, want_act=True))
    fb = fwd_bytes + B * T * H // 8
    print('K2 %-12s %.3f ms  %.0f GB/s (%.1f%%)' % ('fwd+drop+act', t, fb / t / 1e6, 100 * fb / t / 1e6 / peak))
    out, act = ops.aggregate_fwd(y, csr, bias, drop_p=0.5, rng_state=rng, want_act=True)
    gout = torch.randn(B, T, H, device='cuda')
    bwd_bytes = 3 * B * T * H * 4 + 4 * B * (T + 1) + 4 * nnz + 4 * B * T
    t = timeit(lambda: ops.aggregate_bwd(gout, out, csr, drop_p=0.5, force_vec=a.vec), reps=6)
    print('K2 %-12s %.3f ms  %.0f GB/s (%.1f%%)' % ('bwd(out)', t, bwd_bytes / t / 1e6, 100 * bwd_bytes / t / 1e6 / peak))
    bwd_bytes = 2 * B * T * H * 4 + B * T * H // 8 + 4 * B * (T + 1) + 4 * nnz + 4 * B * T
    t = timeit(lambda: ops.aggregate_bwd(gout, None, csr, drop_p=0.5, force_vec=a.vec, act=act), reps=6)
    print('K2 %-12s %.3f ms  %.0f GB/s (%.1f%%)' % ('bwd(act)', t, bwd_bytes / t / 1e6, 100 * bwd_bytes / t / 1e6 / peak))
    g = torch.randn(B, T, H, device='cuda')
    bwd_bytes = 2 * B * T * H * 4 + 4 * B * (T + 1) + 4 * nnz + 4 * B * T
    db = torch.zeros(H, device='cuda')
    t = timeit(lambda: ops.aggregate_bwd_pre(g, csr, dbias_out=db), reps=6)
    print('K2 %-12s %.3f ms  %.0f GB/s (%.1f%%)' % ('bwd(pre)', t, bwd_bytes / t / 1e6, 100 * bwd_bytes / t / 1e6 / peak))
    del g
    h = out
    pool_bytes = B * T * H * 4 + B * 3 * H * 8 + B * T
    t = timeit(lambda: ops._Pool3.apply(h, csr, 0))
    print('K4 %-12s %.3f ms  %.0f GB/s (%.1f%%)' % ('pool3 fwd', t, pool_bytes / t / 1e6, 100 * pool_bytes / t / 1e6 / peak))
    # plain device copy of the same size for reference
    z = torch.empty_like(y)
    t = timeit(lambda: z.copy_(y))
    print('copy %-10s %.3f ms  %.0f GB/s' % ('y->z', t, 2 * y.numel() * 4 / t / 1e6))


if __name__ == '__main__':
    main()
