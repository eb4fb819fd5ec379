#!/usr/bin/env python
"""Micro-benchmark of K3: the FFMA fp32 GEMM vs the tcgen05/TMEM TF32 GEMM (fwd / dgrad / wgrad), TFLOP/s."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import ops  # noqa: E402


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
    ap.add_argument('--shapes', default='2750x200x360,2750x200x200,2097152x512x360,2097152x512x512')
    ap.add_argument('--persist', action='store_true', help='only compare the large-M projection kernels + cuBLAS')
    a = ap.parse_args()
    for shp in a.shapes.split(','):
        M, N, K = (int(v) for v in shp.split('x'))
        x = torch.randn(M, K, device='cuda')
        w = torch.randn(N, K, device='cuda')
        dy = torch.randn(M, N, device='cuda')
        fl = 2.0 * M * N * K
        if M >= 65536 and a.persist:
            # the large-M projection kernels side by side: one tile per CTA (0), persistent single-CTA tiles (1), persistent
            # CTA pairs (2, the default); algorithmic HBM bytes = X read once + Y written once
            gb = (M * K + M * N) * 4 / 1e9
            for cg in (0, 1, 2):
                ops.gemm_persist_config(cg, 65536)
                for mode in ('tf32', 'tf32x3'):
                    ws = ops.weight_prep(w, mode)
                    t = timeit(lambda: ops.linear_fwd(x, w, mode, ws))
                    print('%-22s fwd   %-6s persist=%d %9.3f ms %8.1f TFLOP/s %7.0f GB/s' % (shp, mode, cg, t, fl / t / 1e9, gb / t * 1e3))
                    t = timeit(lambda: ops.linear_dgrad(dy, w, mode, ws))
                    print('%-22s dgrad %-6s persist=%d %9.3f ms %8.1f TFLOP/s' % (shp, mode, cg, t, fl / t / 1e9))
            ops.gemm_persist_config(2, 65536)
        for mode in (('fp32', 'tf32', 'tf32x3') if not a.persist else ()):
            t = timeit(lambda: ops.linear_fwd(x, w, mode))
            print('%-22s fwd   %-5s %9.3f ms %8.1f TFLOP/s' % (shp, mode, t, fl / t / 1e9))
            t = timeit(lambda: ops.linear_dgrad(dy, w, mode))
            print('%-22s dgrad %-5s %9.3f ms %8.1f TFLOP/s' % (shp, mode, t, fl / t / 1e9))
        if not a.persist:
            t = timeit(lambda: ops.linear_wgrad(dy, x), reps=3, warm=1)
            print('%-22s wgrad %-5s %9.3f ms %8.1f TFLOP/s' % (shp, 'fp32', t, fl / t / 1e9))
        if ops.wgrad_tc_ok(M, N, K):
            dw = torch.zeros(N, K, device='cuda')
            t = timeit(lambda: ops.linear_wgrad(dy, x, 'tf32x3', out=dw, accumulate=True))
            print('%-22s wgrad %-5s %9.3f ms %8.1f TFLOP/s' % (shp, 'tf32x3', t, fl / t / 1e9))
        torch.backends.cuda.matmul.allow_tf32 = True
        t = timeit(lambda: torch.matmul(x, w.t()))
        print('%-22s fwd   %-5s %9.3f ms %8.1f TFLOP/s' % (shp, 'cublas-tf32', t, fl / t / 1e9))
        torch.backends.cuda.matmul.allow_tf32 = False
        t = timeit(lambda: torch.matmul(x, w.t()))
        print('%-22s fwd   %-5s %9.3f ms %8.1f TFLOP/s' % (shp, 'cublas-fp32', t, fl / t / 1e9))
        del x, w, dy
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
