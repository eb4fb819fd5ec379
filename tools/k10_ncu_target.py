#!/usr/bin/env python
"""ncu target: a few eager training steps of the full_deprel mode (D = 50 relation slots on a 200-wide input, B = 50
TACRED-shaped sentences, prune_k = 1) -- every K10 kernel of the step, launched one after the other.

    ncu --set full --clock-control none --import-source on -k regex:'relmix|agg3|tf32_gemm|wgrad_tf32x3|live_rows|gather_rows|scatter_rows|colsum' \
        --launch-skip <2 steps> -o gpurun_out/k10 python tools/k10_ncu_target.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import synth  # noqa: E402
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer  # noqa: E402


def main(steps=3):
    torch.manual_seed(0)
    tr = GCNTrainer(synth.tacred_opt(vocab_size=50000, cuda=True, gemm_mode='tf32x3', prune_k=1, adj_type='full_deprel',
                                     deprel_emb_dim=50, emb_dim=140))
    tr.model.train()
    batch = tuple(t.cuda() if torch.is_tensor(t) else t for t in synth.make_batch(2000, batch_size=50, vocab_size=50000))
    for _ in range(steps):
        tr.optimizer.zero_grad()
        loss = tr.update(batch)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(tr.model.parameters(), 5.0)
        tr.optimizer.step()
    torch.cuda.synchronize()
    print('loss %.4f' % float(loss))


if __name__ == '__main__':
    main()
