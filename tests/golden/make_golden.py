#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference, read-only).

Run from the repo root, in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

What is stored (inputs + reference outputs, never reference code):
  adjacency.npz  bundled dataset/tacred/{train,dev,test}.json as loader batches (ids only) + the reference's
                 dense adjacency (uint8) for prune_k in PRUNE_KS; sha256/value-sum of synthetic batches;
                 hand-written edge-case trees.  The SURVEY.md §8c whole-split hashes are asserted here.
  model.npz      for every MODEL_CASES entry: eval-mode logits + loss of the reference GCNClassifier /
                 GCNTrainer.update with weights from weights.make_state(seed); for GRAD_CASES also the
                 train-mode loss and gradient digests with dropout drawn after torch.manual_seed(DROPOUT_SEED).

The only shim applied to the reference is ``Tree.head = None`` (prune_k=-1 crashes without it).
"""
import hashlib
import io
import json
import os
import sys
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REF)          # reference's model/ data/ utils/ packages
sys.path.insert(1, REPO)
sys.path.insert(2, HERE)

import cases                     # noqa: E402
import weights                   # noqa: E402
from gcn_over_pruned_trees_b200 import synth   # noqa: E402  (pure numpy/torch host code)

with contextlib.redirect_stdout(io.StringIO()):
    from model import tree as ref_tree          # noqa: E402
    from model.trainer import GCNTrainer        # noqa: E402
    from data.loader import DataLoader          # noqa: E402
ref_tree.Tree.head = None


class _Vocab:
    def __init__(self, words):
        self.id2word = words
        self.word2id = {w: i for i, w in enumerate(words)}
        self.size = len(words)


def ref_adjacency(head, subj_pos, obj_pos, deprel, lens, k, words=None):
    """What inputs_to_tree_reps computes (/root/reference/model/gcn.py:102-108), via the reference functions."""
    maxlen = int(max(lens))
    words = np.zeros_like(head) if words is None else words
    out = []
    for i in range(len(lens)):
        t = ref_tree.head_to_tree(head[i], words[i], int(lens[i]), k, subj_pos[i], obj_pos[i], deprel[i])
        out.append(ref_tree.tree_to_adj(maxlen, t, directed=False, self_loop=True)[None])
    return np.concatenate(out, 0)


def digest(adj):
    u8 = adj.astype(np.uint8)
    assert (u8 == adj).all()
    return hashlib.sha256(u8.tobytes()).hexdigest()[:16], int(adj.sum())


# SURVEY.md §8c, recorded by an independent probe of the reference
SURVEY_HASHES = {
    ('train', -1): ('358c483221819bb8', 100802), ('train', 0): ('afe69c45eb63bb06', 17026),
    ('train', 1): ('4b7ad19e26cbeb62', 45710), ('train', 2): ('25528f5617bc7890', 57830),
    ('dev', -1): ('3ca75a7f4c283e09', 95434), ('dev', 0): ('549ed9385b3073d7', 11836),
    ('dev', 1): ('8283c38dafc184a9', 36432), ('dev', 2): ('4e8653a1ea56ab3a', 53774),
    ('test', -1): ('73c3114a62007f8a', 100720), ('test', 0): ('30d29f6bae5d8a33', 13934),
    ('test', 1): ('dba54554fd642010', 46252), ('test', 2): ('2aa160408d024f8b', 62218),
}


def main():
    adj_out, model_out = {}, {}
    raw = {s: json.load(open('%s/dataset/tacred/%s.json' % (REF, s))) for s in cases.SPLITS}
    vocab = _Vocab(cases.bundled_vocab(raw))
    print('bundled vocab size', vocab.size)
    loader_opt = dict(lower=False, word_dropout=0.04, use_bert_embeddings=False)
    split_batches = {}
    for s in cases.SPLITS:
        with contextlib.redirect_stdout(io.StringIO()):
            dl = DataLoader('%s/dataset/tacred/%s.json' % (REF, s), 50, loader_opt, vocab, evaluation=True)
        batch = dl[0]
        split_batches[s] = batch
        names = ('words', 'masks', 'pos', 'ner', 'deprel', 'head', 'subj_pos', 'obj_pos', 'rels')
        for name, t in zip(names, batch[:9]):
            if name != 'masks':
                adj_out['%s/%s' % (s, name)] = t.numpy().astype(np.int16)
        adj_out['%s/orig_idx' % s] = np.asarray(batch[9], dtype=np.int16)
        lens = (~batch[1]).sum(1).numpy()
        for k in cases.PRUNE_KS:
            a = ref_adjacency(batch[5].numpy(), batch[6].numpy(), batch[7].numpy(), batch[4].numpy(), lens, k,
                              batch[0].numpy())
            adj_out['%s/adj_k%d' % (s, k)] = a.astype(np.uint8)
            if (s, k) in SURVEY_HASHES:
                assert digest(a) == SURVEY_HASHES[(s, k)], (s, k, digest(a))
    adj_out['vocab_size'] = np.int64(vocab.size)
    print('bundled splits: SURVEY §8c hashes reproduced')

    # synthetic batches: digest only
    for seed in cases.SYNTH_ADJ_SEEDS:
        b = synth.make_batch(seed, batch_size=50)
        lens = synth.batch_lengths(b).numpy()
        for k in cases.PRUNE_KS:
            a = ref_adjacency(b[5].numpy(), b[6].numpy(), b[7].numpy(), b[4].numpy(), lens, k)
            h, sm = digest(a)
            adj_out['synth/%d/k%d/sha' % (seed, k)] = np.frombuffer(h.encode(), dtype=np.uint8)
            adj_out['synth/%d/k%d/sum' % (seed, k)] = np.int64(sm)
    # one 512-token batch (cfg5 shape, reduced B)
    b = synth.make_batch(900, batch_size=6, fixed_len=512)
    for k in (-1, 1):
        a = ref_adjacency(b[5].numpy(), b[6].numpy(), b[7].numpy(), b[4].numpy(), [512] * 6, k)
        h, sm = digest(a)
        adj_out['synth512/k%d/sha' % k] = np.frombuffer(h.encode(), dtype=np.uint8)
        adj_out['synth512/k%d/sum' % k] = np.int64(sm)

    # hand-written edge cases
    for name, (head, subj, obj, deprel) in cases.EDGE_TREES.items():
        n = len(head)
        args = (np.asarray([head]), cases.positions(subj, n)[None], cases.positions(obj, n)[None],
                np.asarray([deprel]), [n])
        for k in cases.PRUNE_KS:
            adj_out['edge/%s/k%d' % (name, k)] = ref_adjacency(*args, k)[0].astype(np.uint8)

    # model cases
    for name, (over, source, wseed) in cases.MODEL_CASES.items():
        if source[0] == 'split':
            batch = split_batches[source[1]]
            over = dict(over, vocab_size=vocab.size)
        else:
            batch = synth.make_batch(source[1], batch_size=source[2], vocab_size=over['vocab_size'],
                                     num_class=over.get('num_class', 42), dataset=over.get('dataset', 'tacred'))
        opt = synth.tacred_opt(**over)
        with contextlib.redirect_stdout(io.StringIO()):
            trainer = GCNTrainer(dict(opt))
        state = {k: torch.from_numpy(v) for k, v in weights.make_state(opt, wseed).items()}
        trainer.model.load_state_dict(state)
        trainer.model.eval()
        with torch.no_grad():
            logits, h_out = trainer.model(list(batch[:-2]))
            loss = trainer.update(batch)
        model_out['%s/logits' % name] = logits.numpy()
        model_out['%s/h_out' % name] = h_out.numpy()
        model_out['%s/eval_loss' % name] = np.float32(loss.item())
        preds, probs, ploss = trainer.predict(batch)
        model_out['%s/pred' % name] = np.asarray(preds, dtype=np.int16)
        model_out['%s/probs' % name] = np.asarray(probs, dtype=np.float32)
        model_out['%s/predict_loss' % name] = np.float32(ploss)
        if name in cases.GRAD_CASES:
            trainer.model.train()
            trainer.model.zero_grad()
            torch.manual_seed(cases.DROPOUT_SEED)
            loss = trainer.update(batch)
            loss.backward()
            model_out['%s/train_loss' % name] = np.float32(loss.item())
            seen = set()
            for key, p in trainer.model.named_parameters():
                if p.grad is None or id(p) in seen:
                    continue
                seen.add(id(p))
                sample, norm, total = weights.grad_digest(p.grad.numpy())
                model_out['%s/grad/%s/sample' % (name, key)] = sample
                model_out['%s/grad/%s/norm' % (name, key)] = norm
                model_out['%s/grad/%s/sum' % (name, key)] = total
        print('model case', name, 'loss', float(loss.item()))

    np.savez_compressed(os.path.join(HERE, 'adjacency.npz'), **adj_out)
    np.savez_compressed(os.path.join(HERE, 'model.npz'), **model_out)
    for f in ('adjacency.npz', 'model.npz'):
        print(f, os.path.getsize(os.path.join(HERE, f)), 'bytes')


if __name__ == '__main__':
    main()
