#!/usr/bin/env python
"""K5 backward at the large shape, alone: the scatter kernel (fp reductions into the word table) vs the grouped path
(counting sort by word + one sum per word), CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import ops  # noqa: E402


def main():
    n, V, E, Dp, Dn = 4096 * 512, 50000, 300, 30, 30
    g = torch.Generator(device='cuda').manual_seed(1)
    words = torch.randint(2, V, (n,), device='cuda', generator=g)
    pos = torch.randint(2, 47, (n,), device='cuda', generator=g)
    ner = torch.randint(2, 15, (n,), device='cuda', generator=g)
    flags = torch.ones(n, dtype=torch.uint8, device='cuda')
    dx = torch.randn(n, E + Dp + Dn, device='cuda', generator=g)
    rng = torch.tensor([5, 9], dtype=torch.int64, device='cuda')
    G = torch.zeros(V, E, device='cuda')
    gp, gn = torch.zeros(47, Dp, device='cuda'), torch.zeros(15, Dn, device='cuda')
    owner = torch.full((V,), 0x7fffffff, dtype=torch.int32, device='cuda')
    for name, min_rows, tables in (('scatter', 1 << 30, (gp, gn)), ('grouped', 65536, (gp, gn)),
                                   ('grouped, word table only', 65536, (None, None)),
                                   ('scatter, word table only', 1 << 30, (None, None))):
        ops.EMBED_GROUPED_MIN_ROWS = min_rows
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for i in range(7):
            if i >= 2:
                ev[i - 2][0].record()
            ops.embed_bwd(dx, flags, words, pos if tables[0] is not None else None, ner if tables[1] is not None else None,
                          G, tables[0], tables[1], owner, V, E, V, 0.5, rng, 0xE0)
            if i >= 2:
                ev[i - 2][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
        print('%-28s %7.3f ms   (dX read once = %.2f GB -> %.0f GB/s)' % (name, ms, dx.numel() * 4 / 1e9,
                                                                          dx.numel() * 4 / ms / 1e6))


if __name__ == '__main__':
    main()
