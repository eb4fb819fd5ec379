#!/usr/bin/env python
"""Training throughput of the other BASELINE.json configurations on one B200 (bench.py reports configs[1] at k=1):
configs[1] prune_k sweep (-1, 0, 1, 2), configs[2] C-GCN (BiLSTM 200 + 2-layer GCN), configs[3] SemEval shape
(9-tuple batches, 19 classes, lengths clip(Poisson(19), 5, 97)).  One step = zero_grad + forward + loss + backward +
clip + SGD through `trainer.train_step` (the engine the configuration supports) from device-resident batches; CUDA
events over `--steps` steps, L2 flushed between steps."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import synth  # noqa: E402
from gcn_over_pruned_trees_b200.engine import PackedBatch  # noqa: E402
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer  # noqa: E402


def run(name, over, batch_kw, steps, n_batches=8):
    torch.manual_seed(0)
    opt = synth.tacred_opt(vocab_size=50000, cuda=True, gemm_mode='tf32x3', **over)
    sys.stdout = open(os.devnull, 'w')
    tr = GCNTrainer(opt)
    sys.stdout = sys.__stdout__
    tr.model.train()
    host = [synth.make_batch(2000 + i, batch_size=50, vocab_size=50000, **batch_kw) for i in range(n_batches)]
    dev = [tuple(t.cuda() if torch.is_tensor(t) else t for t in b) for b in host]
    for _ in range(5):
        for b in dev:
            tr.train_step(b)
    engine = type(tr._graphed).__name__ if getattr(tr, '_graphed', None) is not None else 'eager (cuDNN LSTM + autograd)'
    if engine == 'FusedTrainStep':
        dev = [PackedBatch(b, device='cpu').to('cuda') for b in host]
        for b in dev:
            tr.train_step(b)
    flush = torch.empty(64 << 20, dtype=torch.float32, device='cuda')
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(ev):
        flush.fill_(0.0)
        a.record()
        loss = tr.train_step(dev[i % n_batches])
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    kept = 0
    csr = getattr(tr.model.gcn_model, 'last_csr', None) or getattr(tr._graphed, 'last_csr', None)
    if csr is not None:
        kept = int((csr.flags & 1).sum())
    out = {'config': name, 'engine': engine, 'ms_per_step': ms, 'sentences_per_s': 50 / ms * 1e3, 'loss': float(loss),
           'kept_rows_last_batch': kept}
    print(json.dumps(out), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--relation', action='store_true',
                    help='only the relation-aware adjacency modes (K10): full_deprel with 50 relation slots '
                         '(train_cgcn.sh:5) on a 200-wide input (the mode needs in_dim == hidden_dim), diagonal_deprel')
    a = ap.parse_args()
    if a.relation:
        run('gcn_full_deprel_d50_k1', dict(prune_k=1, adj_type='full_deprel', deprel_emb_dim=50, emb_dim=140), {},
            a.steps)
        run('gcn_diagonal_deprel_k1', dict(prune_k=1, adj_type='diagonal_deprel'), {}, a.steps)
        return
    for k in (-1, 0, 1, 2):
        run('gcn_tacred_k%d' % k, dict(prune_k=k), {}, a.steps)
    run('cgcn_tacred_k1', dict(prune_k=1, rnn=True, rnn_hidden=200, rnn_layers=1), {}, a.steps)
    run('gcn_semeval_k1_19cls', dict(prune_k=1, dataset='semeval', num_class=19),
        dict(dataset='semeval', num_class=19, mean_len=19, min_len=5, max_len=97), a.steps)


if __name__ == '__main__':
    main()
