import sys; sys.path.insert(0, '.')
import torch
from gcn_over_pruned_trees_b200 import ops
M, N, K = 128, 16, 32
# x[0,0] = 1 + 2^-11 + 2^-12 (below the tf32 lsb 2^-10); w = identity-ish so y[0,0] = x[0,0] as seen by the tensor core
x = torch.zeros(M, K, device='cuda'); w = torch.zeros(N, K, device='cuda')
vals = [1 + 2**-11 + 2**-12, 1 + 2**-11, 1 + 2**-12, 1 + 2**-10 + 2**-11, -(1 + 2**-11 + 2**-12), 1 + 2**-13]
for i, v in enumerate(vals):
    x[i, 0] = v
w[0, 0] = 1.0
y = ops.linear_fwd(x, w, 'tf32')
for i, v in enumerate(vals):
    print('in % .10f -> out % .10f   trunc % .10f' % (v, y[i, 0].item(), torch.tensor(v).view(torch.int32).bitwise_and(-8192).view(torch.float32).item()))
# same for the B operand
x.zero_(); w.zero_(); x[0, 0] = 1.0
for i, v in enumerate(vals[:4]):
    w[i, 0] = v
y = ops.linear_fwd(x, w, 'tf32')
print('B operand:', [y[0, i].item() for i in range(4)])
