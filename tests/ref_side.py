"""Reference-side half of the checkpoint-exchange tests: runs in its OWN process with baseline/_ref (the unmodified
reference) first on sys.path, on the CPU, and talks to the test through files.  Test infrastructure.

    python tests/ref_side.py save  --workdir W --ckpt C --out O     # reference trainer -> checkpoint + predict() dump
    python tests/ref_side.py load  --workdir W --ckpt C --out O     # checkpoint (from either side) -> predict() dump

predict() dumps: npz with per-split arrays preds [n], probs [n, num_class], loss [n_batches] for the bundled dev/test
splits read through the reference's own DataLoader (evaluation=True, batch 50).
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, 'baseline', '_ref')
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from data.loader import DataLoader  # noqa: E402  (the reference's)
from model.trainer import GCNTrainer  # noqa: E402
from utils import constant, torch_utils  # noqa: E402
from utils.vocab import Vocab  # noqa: E402


def predict_all(trainer, opt, vocab, data_dir):
    out = {}
    for split in ('dev', 'test'):
        loader = DataLoader(os.path.join(data_dir, split + '.json'), 50, opt, vocab, evaluation=True)
        preds, probs, losses = [], [], []
        for b in loader:
            p, q, loss = trainer.predict(b)
            preds += p
            probs += q
            losses.append(loss)
        out[split + '_preds'] = np.array(preds)
        out[split + '_probs'] = np.array(probs, dtype=np.float64)
        out[split + '_loss'] = np.array(losses)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('mode', choices=['save', 'load'])
    ap.add_argument('--workdir', required=True)
    ap.add_argument('--ckpt', required=True)
    ap.add_argument('--out', required=True)
    ap.add_argument('--opt', default=None, help='json file with option overrides (save mode)')
    ap.add_argument('--steps', type=int, default=3)
    args = ap.parse_args()
    data_dir, vocab_dir = os.path.join(args.workdir, 'data'), os.path.join(args.workdir, 'vocab')
    vocab = Vocab(os.path.join(vocab_dir, 'vocab.pkl'), load=True)
    torch.manual_seed(4321)
    np.random.seed(4321)
    if args.mode == 'save':
        opt = json.load(open(args.opt))
        opt.update(vocab_size=vocab.size, num_class=len(constant.LABEL_TO_ID), cuda=False, cpu=True)
        emb = np.load(os.path.join(vocab_dir, 'embedding.npy'))
        trainer = GCNTrainer(opt, emb_matrix=emb)
        train = DataLoader(os.path.join(data_dir, 'train.json'), 50, opt, vocab, evaluation=False)
        trainer.model.train()
        for s in range(args.steps):            # a few real steps so that the weights are not the initial ones
            for b in train:
                trainer.optimizer.zero_grad()
                loss = trainer.update(b)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(trainer.model.parameters(), opt['max_grad_norm'])
                trainer.optimizer.step()
        trainer.save(args.ckpt, 1)
    else:
        opt = torch_utils.load_config(args.ckpt)
        opt['cuda'], opt['cpu'] = False, True
        trainer = GCNTrainer(opt)
        trainer.load(args.ckpt)
        trainer.opt['cuda'], trainer.opt['cpu'] = False, True
    np.savez(args.out, **predict_all(trainer, trainer.opt, vocab, data_dir))
    print('ref_side %s ok' % args.mode)


if __name__ == '__main__':
    main()
