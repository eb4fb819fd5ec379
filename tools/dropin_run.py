#!/usr/bin/env python
"""Run the reference's UNMODIFIED drivers (train.py, eval.py) with `model` (and optionally `data`) resolved to this
package -- the drop-in check of SURVEY.md 4(d) / 8(b).

    python tools/dropin_run.py [--workdir DIR] [--model b200|reference] [--loader b200|reference] [--epochs 2] [--cpu]

What a maintainer of the reference does to switch: put `gcn_over_pruned_trees_b200/` on PYTHONPATH.
The reference's `model/`, `data/` and `utils/` are namespace packages (there is no __init__.py in its tree), so the
regular packages `model` and `data` of this repository win the import even though the script's own directory comes
first on sys.path: `from model.trainer import GCNTrainer` (train.py:22, eval.py:13) and `from data.loader import
DataLoader` (train.py:21) resolve here, `utils` stays the reference's.  --loader reference sets GPT_DATA_LOADER=reference,
which hands `data.*` back to the reference's own modules (host batches; only `model` is replaced).  The drivers run IN
PLACE from baseline/_ref (a byte-identical copy of the reference, tools/install_reference.py).

The bundled sample (dataset/tacred/{train,dev,test}.json, 20 sentences each) is the data; train.py wants
`train_0.1.json`, a vocab pickle and an embedding matrix (prepare_vocab.py needs GloVe, which is not available offline):
`make_fixture` builds them from the sample's own tokens with random vectors.
"""
import argparse
import json
import os
import pickle
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, 'baseline', '_ref')
PKG = os.path.join(REPO, 'gcn_over_pruned_trees_b200')


def make_fixture(workdir, emb_dim=300, seed=1234):
    """workdir/data/{train_0.1,dev,test}.json + workdir/vocab/{vocab.pkl,embedding.npy}."""
    import numpy as np
    data, vocab = os.path.join(workdir, 'data'), os.path.join(workdir, 'vocab')
    os.makedirs(data, exist_ok=True)
    os.makedirs(vocab, exist_ok=True)
    src = os.path.join(REF, 'dataset', 'tacred')
    shutil.copyfile(os.path.join(src, 'train.json'), os.path.join(data, 'train_0.1.json'))
    for split in ('train', 'dev', 'test'):
        shutil.copyfile(os.path.join(src, split + '.json'), os.path.join(data, split + '.json'))
    counts = {}
    for ex in json.load(open(os.path.join(src, 'train.json'))):
        for tok in ex['token']:
            counts[tok] = counts.get(tok, 0) + 1
    # utils/vocab.py:45-62: [<PAD>, <UNK>] + words by descending count (the entity masks the loader substitutes,
    # SUBJ-*/OBJ-*, fall to <UNK> here exactly as they would with a GloVe vocab that lacks them)
    id2word = ['<PAD>', '<UNK>'] + sorted(counts, key=lambda w: (-counts[w], w))
    with open(os.path.join(vocab, 'vocab.pkl'), 'wb') as f:
        pickle.dump(id2word, f)
    rng = np.random.RandomState(seed)
    emb = rng.uniform(-1, 1, (len(id2word), emb_dim))
    emb[0] = 0
    np.save(os.path.join(vocab, 'embedding.npy'), emb)
    return data, vocab


def _env(model, loader):
    env = dict(os.environ)
    env['PYTHONPATH'] = PKG if model == 'b200' else ''
    env['GPT_DATA_LOADER'] = loader
    env.setdefault('MPLBACKEND', 'Agg')
    return env


def run_train(workdir, model='b200', loader='reference', epochs=2, cpu=False, extra=(), model_id='01', timeout=1200):
    if model == 'reference' and loader == 'b200':
        raise ValueError('the reference model takes host batches: use its own loader')
    data, vocab = make_fixture(workdir)
    driver = REF
    save = os.path.join(workdir, 'saved_models_%s_%s' % (model, loader))
    cmd = [sys.executable, os.path.join(driver, 'train.py'), '--data_dir', data, '--vocab_dir', vocab,
           '--model_save_dir', save, '--test_save_dir', os.path.join(workdir, 'test_perf'), '--id', model_id,
           '--no-rnn', '--prune_k', '1', '--num_epoch', str(epochs), '--lr', '0.3', '--pooling_l2', '0.003',
           '--batch_size', '50', '--log_step', '1'] + (['--cpu'] if cpu else []) + list(extra)
    out = subprocess.run(cmd, cwd=workdir, env=_env(model, loader), capture_output=True, text=True, timeout=timeout)
    return out, os.path.join(save, model_id)


def run_eval(workdir, model_dir, model='b200', loader='reference', dataset='test', timeout=600):
    cmd = [sys.executable, os.path.join(REF, 'eval.py'), '--model_dir', model_dir, '--model', 'best_model.pt',
           '--data_dir', os.path.join(workdir, 'data'), '--dataset', dataset]
    return subprocess.run(cmd, cwd=workdir, env=_env(model, loader), capture_output=True, text=True, timeout=timeout)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workdir', default='/tmp/gpt_dropin')
    ap.add_argument('--model', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--loader', default='reference', choices=['b200', 'reference'])
    ap.add_argument('--epochs', type=int, default=2)
    ap.add_argument('--cpu', action='store_true')
    args = ap.parse_args()
    os.makedirs(args.workdir, exist_ok=True)
    out, model_dir = run_train(args.workdir, args.model, args.loader, args.epochs, args.cpu)
    sys.stdout.write(out.stdout[-4000:])
    sys.stderr.write(out.stderr[-4000:])
    if out.returncode != 0:
        sys.exit(out.returncode)
    ev = run_eval(args.workdir, model_dir, args.model, args.loader)
    sys.stdout.write(ev.stdout[-2500:])
    sys.stderr.write(ev.stderr[-2500:])
    sys.exit(ev.returncode)


if __name__ == '__main__':
    main()
