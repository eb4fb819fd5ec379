import sys; sys.path.insert(0, '.')
import torch
from gcn_over_pruned_trees_b200 import ops
for (M, N, K) in [(2750, 200, 400), (2750, 200, 360), (2750, 200, 512), (2750, 512, 200), (300, 200, 400)]:
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(M, K, device='cuda', generator=g); w = torch.randn(N, K, device='cuda', generator=g)
    dy = torch.randn(M, N, device='cuda', generator=g)
    for mode in ('tf32', 'tf32x3'):
        y = ops.linear_fwd(x, w, mode); ref = x.double() @ w.double().t()
        e1 = ((y.double() - ref).abs().max() / ref.abs().max()).item()
        dx = ops.linear_dgrad(dy, w, mode); ref2 = dy.double() @ w.double()
        err = (dx.double() - ref2).abs()
        e2 = (err.max() / ref2.abs().max()).item()
        colerr = err.max(0)[0] / ref2.abs().max()
        bad = (colerr > 1e-2).nonzero().flatten().tolist()
        print(M, N, K, mode, 'fwd %.2e dgrad %.2e' % (e1, e2), 'bad cols', bad[:6], '...', bad[-3:], len(bad))
