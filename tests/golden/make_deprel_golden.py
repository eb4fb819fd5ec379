#!/usr/bin/env python
"""Generate tests/golden/deprel.npz: outputs of the REAL reference (imported from /root/reference, read-only) for the
relation-aware adjacency modes ``full_deprel`` / ``diagonal_deprel``.

    python tests/golden/make_deprel_golden.py        # build container only; the GPU box has no /root/reference

Stored per cases.DEPREL_CASES entry: eval-mode logits, pooled h_out and loss of GCNClassifier / GCNTrainer.update with
weights from weights.make_state(seed); for DEPREL_GRAD_CASES and DEPREL_RANDOM_CASES also the train-mode loss and
gradient digests with every random draw (input/GCN dropout, edge dropout, relation forgetting) taken after
torch.manual_seed(DROPOUT_SEED).  Inputs are rebuilt from seeds (synth.make_batch) or from adjacency.npz.
"""
import contextlib
import io
import os

import numpy as np
import torch

import make_golden as mg          # imports the reference, applies the Tree.head shim
from make_golden import cases, weights, synth, GCNTrainer, HERE


def main():
    out = {}
    golden_adj = np.load(os.path.join(HERE, 'adjacency.npz'))
    todo = dict(cases.DEPREL_CASES)
    todo.update(cases.DEPREL_RANDOM_CASES)
    for name, (over, source, wseed) in todo.items():
        if source[0] == 'split':
            batch = cases.batch_from_npz(golden_adj, source[1])
            over = dict(over, vocab_size=int(golden_adj['vocab_size']))
        else:
            batch = cases.make_case_batch(source, over)
        opt = synth.tacred_opt(**over)
        with contextlib.redirect_stdout(io.StringIO()):
            trainer = GCNTrainer(dict(opt))
        state = {k: torch.from_numpy(v) for k, v in weights.make_state(opt, wseed).items()}
        trainer.model.load_state_dict(state)           # strict: the key layout of these modes is part of the fixture
        trainer.model.eval()
        with torch.no_grad():
            logits, h_out = trainer.model(list(batch[:-2]))
            loss = trainer.update(batch)
        out['%s/logits' % name] = logits.numpy()
        out['%s/h_out' % name] = h_out.numpy()
        out['%s/eval_loss' % name] = np.float32(loss.item())
        if name in cases.DEPREL_GRAD_CASES or name in cases.DEPREL_RANDOM_CASES:
            trainer.model.train()
            trainer.model.zero_grad()
            torch.manual_seed(cases.DROPOUT_SEED)
            loss = trainer.update(batch)
            loss.backward()
            out['%s/train_loss' % name] = np.float32(loss.item())
            seen = set()
            for key, p in trainer.model.named_parameters():
                if p.grad is None or id(p) in seen:
                    continue
                seen.add(id(p))
                sample, norm, total = weights.grad_digest(p.grad.numpy())
                out['%s/grad/%s/sample' % (name, key)] = sample
                out['%s/grad/%s/norm' % (name, key)] = norm
                out['%s/grad/%s/sum' % (name, key)] = total
        print('deprel case', name, 'loss', float(loss.item()), 'max|logit|', float(np.abs(logits.numpy()).max()))
    np.savez_compressed(os.path.join(HERE, 'deprel.npz'), **out)
    print('deprel.npz', os.path.getsize(os.path.join(HERE, 'deprel.npz')), 'bytes')


if __name__ == '__main__':
    main()
