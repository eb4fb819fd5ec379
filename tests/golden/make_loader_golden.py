#!/usr/bin/env python
"""Golden vectors for the device-resident batch builder (gcn_over_pruned_trees_b200/data/loader.py, K9).

Runs the UNMODIFIED reference loader (/root/reference/data/loader.py:13-141, semeval_loader.py) in THIS container on a
small synthetic sample in the TACRED json schema (tests/golden/loader_sample.json, written here; the reference's own
bundled sample is licensed data and is not copied) and stores the batches it emits in tests/golden/loader.npz:

  eval/<i>/...        DataLoader(sample, 16, opt, vocab, evaluation=True)          no shuffle, no word dropout
  train/<i>/...       random.seed(5); np.random.seed(7); evaluation=False, word_dropout=0.2, lower=True
  semeval/<i>/...     semeval_loader.DataLoader on the same sentences relabelled with the 10 SemEval classes

    python tests/golden/make_loader_golden.py        (needs /root/reference; the tests only read the .npz / .json)
"""
import contextlib
import io
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)
from gcn_over_pruned_trees_b200 import constant  # noqa: E402

WORDS = ['The', 'company', 'said', 'its', 'founder', 'was', 'born', 'in', 'Paris', ',', 'France', 'and', 'works',
         'for', 'a', 'large', 'bank', 'since', '1999', '.', 'She', 'He', 'married', 'his', 'wife', 'who', 'is', 'an',
         'engineer', 'at', 'Acme', 'Corp', 'of', 'New', 'York', 'city', 'on', 'Monday', 'OOVword', 'Zyx']
SEMEVAL_LABELS = ['Other', 'Entity-Destination', 'Cause-Effect', 'Member-Collection', 'Entity-Origin', 'Message-Topic',
                  'Component-Whole', 'Instrument-Agency', 'Product-Producer', 'Content-Container']


def make_sample(n=37, seed=3):
    rng = np.random.default_rng(seed)
    pos_names = [k for k in constant.POS_TO_ID if k not in constant.VOCAB_PREFIX] + ['XX-unknown-tag']
    ner_names = [k for k in constant.NER_TO_ID if k not in constant.VOCAB_PREFIX]
    dep_names = [k for k, v in constant.DEPREL_TO_ID.items() if 2 <= v < 42 and k != 'ROOT'] + ['weird:dep']
    labels = list(constant.LABEL_TO_ID)
    out = []
    for i in range(n):
        L = int(rng.integers(4, 30))
        perm = rng.permutation(L)
        head = np.zeros(L, dtype=np.int64)
        for j in range(1, L):
            head[perm[j]] = perm[rng.integers(0, j)] + 1
        deprel = [dep_names[int(rng.integers(0, len(dep_names)))] for _ in range(L)]
        deprel[int(perm[0])] = 'ROOT'
        ss = int(rng.integers(0, L - 3))
        se = min(L - 1, ss + int(rng.integers(0, 3)))
        cand = [s for s in range(L) if s > se or s + 2 < ss]
        os_ = int(cand[int(rng.integers(0, len(cand)))])
        oe = os_ if os_ < ss else min(L - 1, os_ + int(rng.integers(0, 3)))
        if os_ < ss:
            oe = min(ss - 1, os_ + int(rng.integers(0, 2)))
        out.append({
            'id': 'synthetic%04d' % i, 'relation': labels[int(rng.integers(0, len(labels)))],
            'token': [WORDS[int(rng.integers(0, len(WORDS)))] for _ in range(L)],
            'subj_start': ss, 'subj_end': se, 'obj_start': os_, 'obj_end': oe,
            'subj_type': ['PERSON', 'ORGANIZATION'][int(rng.integers(0, 2))],
            'obj_type': ['PERSON', 'LOCATION', 'DATE', 'NUMBER'][int(rng.integers(0, 4))],
            'stanford_pos': [pos_names[int(rng.integers(0, len(pos_names)))] for _ in range(L)],
            'stanford_ner': [ner_names[int(rng.integers(0, len(ner_names)))] for _ in range(L)],
            'stanford_head': [str(int(h)) for h in head], 'stanford_deprel': deprel})
    return out


def vocab_words():
    """<PAD>, <UNK>, the entity mask tokens, and the sample's words except two that stay out of vocabulary; the
    lower-cased forms are included so that lower=True maps them too."""
    masks = ['SUBJ-PERSON', 'SUBJ-ORGANIZATION', 'OBJ-PERSON', 'OBJ-LOCATION', 'OBJ-DATE', 'OBJ-NUMBER']
    words = [w for w in WORDS if w not in ('OOVword', 'Zyx')]
    low = [w.lower() for w in words if w.lower() not in words]
    return list(constant.VOCAB_PREFIX) + masks + [m.lower() for m in masks] + words + low


class _Vocab:
    def __init__(self, words):
        self.id2word = words
        self.word2id = {w: i for i, w in enumerate(words)}
        self.size = len(words)


def store(out, prefix, loader, names):
    out[prefix + '/n_batches'] = np.int64(len(loader))
    for i in range(len(loader)):
        batch = loader[i]
        for name, t in zip(names, batch[:len(names)]):
            out['%s/%d/%s' % (prefix, i, name)] = t.numpy().astype(np.uint8 if name == 'masks' else np.int16)
        out['%s/%d/orig_idx' % (prefix, i)] = np.asarray(batch[-1], dtype=np.int16)


def main():
    sample = make_sample()
    path = os.path.join(HERE, 'loader_sample.json')
    with open(path, 'w') as f:
        json.dump(sample, f)
    sem = [dict(d, relation=SEMEVAL_LABELS[i % len(SEMEVAL_LABELS)]) for i, d in enumerate(sample)]
    sem_path = os.path.join(HERE, 'loader_sample_semeval.json')
    with open(sem_path, 'w') as f:
        json.dump(sem, f)
    sys.path.insert(0, REF)
    from data.loader import DataLoader
    from data import semeval_loader
    vocab = _Vocab(vocab_words())
    out = {}
    tacred = ('words', 'masks', 'pos', 'ner', 'deprel', 'head', 'subj_pos', 'obj_pos', 'rels')
    semeval = ('words', 'masks', 'pos', 'deprel', 'head', 'subj_pos', 'obj_pos', 'rels')
    with contextlib.redirect_stdout(io.StringIO()):
        dl = DataLoader(path, 16, dict(lower=False, word_dropout=0.04, use_bert_embeddings=False), vocab,
                        evaluation=True)
        store(out, 'eval', dl, tacred)
        random.seed(5)
        np.random.seed(7)
        dl = DataLoader(path, 16, dict(lower=True, word_dropout=0.2, use_bert_embeddings=False), vocab,
                        evaluation=False)
        store(out, 'train', dl, tacred)       # one pass in batch order: that is how np.random is consumed
        out['train/gold'] = np.asarray([constant.LABEL_TO_ID[x] for x in dl.gold()], dtype=np.int16)
        dl = semeval_loader.DataLoader(sem_path, 16, dict(lower=False, word_dropout=0.0, use_bert_embeddings=False),
                                       vocab, evaluation=True)
        store(out, 'semeval', dl, semeval)
    np.savez_compressed(os.path.join(HERE, 'loader.npz'), **out)
    print('wrote loader.npz with', len(out), 'arrays;', len(sample), 'sentences')


if __name__ == '__main__':
    main()
