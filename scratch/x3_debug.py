import sys; sys.path.insert(0, '.')
import torch
from gcn_over_pruned_trees_b200 import ops
for (M, N, K) in [(128, 200, 360), (256, 200, 360), (2750, 200, 360), (2750, 16, 32), (2750, 200, 32), (2750, 200, 64)]:
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(M, K, device='cuda', generator=g); w = torch.randn(N, K, device='cuda', generator=g)
    ref = (x.double() @ w.double().t())
    for rep in range(2):
        y = ops.linear_fwd(x, w, 'tf32x3')
        err = (y.double() - ref).abs()
        rowerr = err.max(1)[0] / ref.abs().max()
        blocks = rowerr.view(-1)[: (M // 128) * 128].view(-1, 128).max(1)[0] if M >= 128 else rowerr.max()[None]
        print(M, N, K, 'rep', rep, 'max rel %.2e' % (err.max() / ref.abs().max()).item(), 'bad blocks', (blocks > 4e-6).nonzero().flatten().tolist()[:20], 'of', blocks.numel())
