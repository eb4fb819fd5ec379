import sys, time, torch
sys.path.insert(0, '/root/repo')
from gcn_over_pruned_trees_b200 import synth
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
torch.manual_seed(0)
opt = synth.tacred_opt(vocab_size=50000, cuda=True, gemm_mode='tf32x3', prune_k=1, rnn=True, rnn_hidden=200, rnn_layers=1)
tr = GCNTrainer(opt); tr.model.train()
bs = [tuple(t.cuda() if torch.is_tensor(t) else t for t in synth.make_batch(2000 + i, batch_size=50, vocab_size=50000)) for i in range(4)]
params = list(tr.model.parameters())
def step(b, T):
    t0 = time.perf_counter(); tr.optimizer.zero_grad(set_to_none=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    loss = tr.update(b); torch.cuda.synchronize(); t2 = time.perf_counter()
    loss.backward(); torch.cuda.synchronize(); t3 = time.perf_counter()
    torch.nn.utils.clip_grad_norm_(params, 5.0); torch.cuda.synchronize(); t4 = time.perf_counter()
    tr.optimizer.step(); torch.cuda.synchronize(); t5 = time.perf_counter()
    for k, v in zip(('zero', 'fwd', 'bwd', 'clip', 'sgd'), (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)): T[k] = T.get(k, 0) + v
for i in range(10): step(bs[i % 4], {})
T = {}
for i in range(40): step(bs[i % 4], T)
print({k: round(v / 40 * 1e3, 3) for k, v in T.items()}, 'ms; total', round(sum(T.values()) / 40 * 1e3, 3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(5): step(bs[i % 4], {})
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=60)[:4500])
