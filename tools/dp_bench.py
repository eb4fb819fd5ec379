#!/usr/bin/env python
"""K8 exchange kernels (csrc/dp.cu) timed on ONE GPU with W virtual ranks (W regions in one process, as
tests/test_gpu_fused.py does): device time of push / reduce / apply at the TACRED-shaped sizes of bench.py.  The push
here goes to local memory, so this measures the kernels' own latency chains, not the NVLink transfer."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--worlds', default='1,2,4,8')
    ap.add_argument('--n_flat', type=int, default=283520)
    ap.add_argument('--V', type=int, default=50000)
    ap.add_argument('--E', type=int, default=300)
    ap.add_argument('--cap', type=int, default=6400)
    ap.add_argument('--slots', type=int, default=2750)       # B*T token slots per rank
    ap.add_argument('--live', type=float, default=0.37)      # share of slots with flags != 0 (k=1)
    ap.add_argument('--reps', type=int, default=20)
    a = ap.parse_args()
    dev = 'cuda'
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    for W in [int(w) for w in a.worlds.split(',')]:
        gen = torch.Generator().manual_seed(W)
        regions = [ops.ExchangeRegion(W, a.cap, a.E, a.V, a.n_flat) for _ in range(W)]
        ptrs = [r.ptr for r in regions]
        shape = regions[0].shape
        params = [torch.randn(a.n_flat, device=dev) for _ in range(W)]
        embs = [torch.randn(a.V, a.E, device=dev) for _ in range(W)]
        partials = torch.zeros(max(1024, regions[0].n_partials), device=dev)
        counter = torch.tensor([1, 0], dtype=torch.int64, device=dev)
        states = [ops.SparseEmbeddingState(embs[r], a.V) for r in range(W)]
        t = {'push': 0.0, 'reduce': 0.0, 'apply': 0.0}
        ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
        for rep in range(a.reps + 3):
            gs = []
            for r in range(W):
                words = torch.randint(2, a.V, (a.slots,), generator=gen)
                live = torch.rand(a.slots, generator=gen) < a.live
                st = states[r]
                st.words = words.to(dev)
                first = {}
                for i, (w, lv) in enumerate(zip(words.tolist(), live.tolist())):
                    if lv and w not in first:
                        first[w] = i
                idx = torch.tensor(list(first.keys()), device=dev)
                st.owner[idx] = torch.tensor(list(first.values()), dtype=torch.int32, device=dev)
                st.G[idx] = 0.01
                gs.append(torch.randn(a.n_flat, device=dev) * 0.01)
            flush.fill_(0.0)
            e = [ev() for _ in range(4)]
            e[0].record()
            for r in range(W):
                ops.dp_push(ptrs, r, shape, gs[r], states[r])
            for r in range(W):
                ops.dp_signal(ptrs, r, shape)
            e[1].record()
            for r in range(W):
                ops.dp_reduce(ptrs, r, shape, gs[r], partials, signal=False)
            e[2].record()
            for r in range(W):
                ops.dp_apply(ptrs[r], shape, params[r], gs[r], embs[r], partials, 5.0, 0.3, None, counter[1:])
            e[3].record()
            torch.cuda.synchronize()
            if rep >= 3:
                for k, i in (('push', 0), ('reduce', 1), ('apply', 2)):
                    t[k] += e[i].elapsed_time(e[i + 1]) / W
        print('W=%d  per rank: push %.1f us  reduce %.1f us  apply %.1f us   (slots %d, live %.0f%%)' %
              (W, t['push'] / a.reps * 1e3, t['reduce'] / a.reps * 1e3, t['apply'] / a.reps * 1e3, a.slots,
               100 * a.live), flush=True)
        for r in regions:
            r.free()


if __name__ == '__main__':
    main()
