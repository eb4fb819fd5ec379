"""Synthetic TACRED-/SemEval-shaped batches (SURVEY.md §8d generator).

There is no TACRED here (LDC licence) and no network, so throughput and full-size parity runs use random
dependency trees with the batch layout the reference's loaders emit
(/root/reference/data/loader.py:81-141 -> 10-tuple, /root/reference/data/semeval_loader.py:75-119 -> 9-tuple):

    (words, masks, pos, [ner,] deprel, head, subj_pos, obj_pos, rels, orig_idx)

* rows sorted by length, descending; padded to the batch maximum
* ``head`` is 1-based, 0 marks the root *and* padding; ``subj_pos/obj_pos`` are 0 inside the span,
  padding is filled with 150 (loader.py:120-121); ``masks`` is True on padding (``words == 0``)
* trees are uniform random recursive trees over a random permutation of the tokens
* ``deprel`` is drawn from [2, 41], the root gets ROOT (11); spans are 1-3 tokens and disjoint
"""
import numpy as np
import torch

from . import constant

POS_FILL = 150


def _positions(start, end, length):
    # same sequence as get_positions (/root/reference/data/loader.py:162-165): ..., -2, -1, 0, .., 0, 1, 2, ...
    idx = np.arange(length)
    return np.where(idx < start, idx - start, np.where(idx > end, idx - end, 0))


def random_tree(rng, length):
    """1-based head array of a uniform random recursive tree with exactly one root (head == 0)."""
    perm = rng.permutation(length)
    head = np.zeros(length, dtype=np.int64)
    for j in range(1, length):
        head[perm[j]] = perm[rng.integers(0, j)] + 1
    return head


def random_spans(rng, length):
    """Two disjoint spans of 1-3 tokens: (subj_start, subj_end, obj_start, obj_end), ends inclusive."""
    if length < 5:          # both starts are drawn from [0, length - 3): below 5 tokens the spans can never be disjoint
        raise ValueError('random_spans needs sentences of at least 5 tokens, got %d' % length)
    while True:
        ss = int(rng.integers(0, max(length - 3, 1)))
        se = min(ss + int(rng.integers(1, 4)) - 1, length - 1)
        os_ = int(rng.integers(0, max(length - 3, 1)))
        oe = min(os_ + int(rng.integers(1, 4)) - 1, length - 1)
        if se < os_ or oe < ss:
            return ss, se, os_, oe


def sample_lengths(rng, batch_size, mean_len=36, min_len=8, max_len=96, fixed_len=None):
    if fixed_len is not None:
        return np.full(batch_size, fixed_len, dtype=np.int64)
    return np.clip(rng.poisson(mean_len, size=batch_size), min_len, max_len).astype(np.int64)


def make_batch(seed, batch_size=50, vocab_size=50000, num_class=42, dataset='tacred',
               mean_len=36, min_len=8, max_len=96, fixed_len=None, pad_to=None):
    """One loader-shaped batch of CPU tensors.  ``pad_to`` forces the padded width (default: batch max)."""
    rng = np.random.default_rng(seed)
    lens = np.sort(sample_lengths(rng, batch_size, mean_len, min_len, max_len, fixed_len))[::-1]
    width = int(lens.max()) if pad_to is None else int(pad_to)
    assert width >= lens.max()

    words = np.zeros((batch_size, width), dtype=np.int64)
    pos = np.zeros_like(words)
    ner = np.zeros_like(words)
    deprel = np.zeros_like(words)
    head = np.zeros_like(words)
    subj_pos = np.full_like(words, POS_FILL)
    obj_pos = np.full_like(words, POS_FILL)
    for b, n in enumerate(lens):
        n = int(n)
        words[b, :n] = rng.integers(2, vocab_size, size=n)
        pos[b, :n] = rng.integers(2, constant.NUM_POS, size=n)
        ner[b, :n] = rng.integers(2, constant.NUM_NER, size=n)
        h = random_tree(rng, n)
        head[b, :n] = h
        rel = rng.integers(2, constant.DEPREL_FORWARD_BOUND, size=n)
        rel[h == 0] = constant.ROOT_DEPREL_ID
        deprel[b, :n] = rel
        ss, se, os_, oe = random_spans(rng, n)
        subj_pos[b, :n] = _positions(ss, se, n)
        obj_pos[b, :n] = _positions(os_, oe, n)
    rels = rng.integers(0, num_class, size=batch_size).astype(np.int64)
    orig_idx = [int(i) for i in rng.permutation(batch_size)]

    t = torch.from_numpy
    words_t = t(words)
    fields = [words_t, words_t.eq(0), t(pos)]
    if dataset == 'tacred':
        fields.append(t(ner))
    fields += [t(deprel), t(head), t(subj_pos), t(obj_pos), t(rels), orig_idx]
    return tuple(fields)


def batch_lengths(batch):
    return (~batch[1]).sum(1)


def tacred_opt(**overrides):
    """The ``opt`` dict train.py builds for train_gcn.sh (/root/reference/train.py:49-134, train_gcn.sh:4)."""
    opt = dict(
        dataset='tacred', emb_dim=300, ner_dim=30, pos_dim=30, hidden_dim=200, num_layers=2,
        input_dropout=0.5, gcn_dropout=0.5, word_dropout=0.04, topn=1e10, lower=False,
        prune_k=1, conv_l2=0.0, pooling='max', pooling_l2=0.003, mlp_layers=2, no_adj=False,
        rnn=False, rnn_hidden=200, rnn_layers=1, rnn_dropout=0.5,
        lr=0.3, lr_decay=0.9, decay_epoch=5, optim='sgd', num_epoch=100, batch_size=50, max_grad_norm=5.0,
        adj_type='regular', deprel_emb_dim=200, deprel_dropout=0.5, deprel_self_loop=True, deprel_directed=False,
        use_bert_embeddings=False, emb_dropout=0.0, deprel_attn=False, deprel_alpha=1.0,
        edge_keep_prob=1.0, deprel_keep_prop=1.0, deprel_max_depth=2,
        vocab_size=50000, num_class=42, cuda=False,
    )
    opt.update(overrides)
    return opt


def make_batch_torch(seed, batch_size, length, vocab_size=50000, num_class=42, device='cuda'):
    """Fixed-length batch (cfg5 shape: every sentence exactly ``length`` tokens) built with vectorised torch ops on
    ``device`` -- same field layout and value ranges as make_batch, same uniform random recursive trees; the two
    spans are drawn from different halves of the sentence (then randomly swapped) so they are disjoint."""
    g = torch.Generator(device=device).manual_seed(seed)
    B, T = batch_size, length

    def randint(lo, hi, shape):
        return torch.randint(lo, hi, shape, device=device, generator=g)

    perm = torch.rand((B, T), device=device, generator=g).argsort(1)
    j = torch.arange(T, device=device)
    r = (torch.rand((B, T), device=device, generator=g) * j).long().clamp_(max=T - 1)
    r = torch.minimum(r, (j - 1).clamp_(min=0))
    parent = perm.gather(1, r)
    head = torch.zeros((B, T), dtype=torch.int64, device=device)
    head.scatter_(1, perm[:, 1:], parent[:, 1:] + 1)
    deprel = randint(2, constant.DEPREL_FORWARD_BOUND, (B, T))
    deprel[head == 0] = constant.ROOT_DEPREL_ID
    half = T // 2
    a0 = randint(0, max(half - 3, 1), (B, 1))
    b0 = randint(half, max(T - 3, half + 1), (B, 1))
    a1 = a0 + randint(0, 3, (B, 1))
    b1 = (b0 + randint(0, 3, (B, 1))).clamp_(max=T - 1)
    swap = randint(0, 2, (B, 1)).bool()
    ss, se = torch.where(swap, b0, a0), torch.where(swap, b1, a1)
    os_, oe = torch.where(swap, a0, b0), torch.where(swap, a1, b1)
    idx = j[None, :]

    def positions(s, e):
        return torch.where(idx < s, idx - s, torch.where(idx > e, idx - e, torch.zeros_like(idx)))

    words = randint(2, vocab_size, (B, T))
    return (words, words.eq(0), randint(2, constant.NUM_POS, (B, T)), randint(2, constant.NUM_NER, (B, T)), deprel,
            head, positions(ss, se), positions(os_, oe), randint(0, num_class, (B,)), list(range(B)))
