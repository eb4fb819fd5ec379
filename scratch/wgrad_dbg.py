import sys, torch, ctypes
sys.path.insert(0, '/root/repo')
from gcn_over_pruned_trees_b200 import ops, _lib
torch.manual_seed(0)
M, N, K = 32768, 128, 32
dbg = torch.full((32768,), -7.0, device='cuda')
_lib.lib().gpt_wgrad_set_debug(ctypes.c_void_p(dbg.data_ptr()))
dy = torch.arange(M * N, device='cuda', dtype=torch.float32).view(M, N) % 1000 + 1; x = torch.ones(M, K, device='cuda') * 2
dw = torch.zeros(N, K, device='cuda')
ops.linear_wgrad(dy, x, 'tf32x3', out=dw, accumulate=True)
torch.cuda.synchronize()
d = dbg.cpu()
for j in range(5):
    print('box', j, d[j * 512: j * 512 + 16].tolist())
print('acc', d[20000:20008].tolist())
print('dw', dw[0, :4].tolist(), float(dw.abs().max()))
