import sys, torch
sys.path.insert(0, '/root/repo')
from gcn_over_pruned_trees_b200 import ops
torch.manual_seed(0)
for (M, N, K) in [(65541, 512, 512), (65536, 512, 512), (65536, 128, 512), (65536, 128, 384), (65536, 128, 288), (65536, 128, 256)]:
    dy = torch.randn(M, N, device='cuda'); x = torch.randn(M, K, device='cuda')
    dw = torch.zeros(N, K, device='cuda')
    ops.linear_wgrad(dy, x, 'tf32x3', out=dw, accumulate=True)
    ref = dy.double().t() @ x.double()
    e = (dw.double() - ref).abs()
    print(M, N, K, 'rel err %.3e' % float(e.max() / ref.abs().max()), 'err by 32-col block:', [ '%.1e' % float(e[:, c:c+32].max() / ref.abs().max()) for c in range(0, K, 32)])
