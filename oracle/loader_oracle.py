"""CPU restatement of the reference loader's per-batch work -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module; the product path
(gcn_over_pruned_trees_b200/data/loader.py + csrc/batch.cu) never does.

Restates DataLoader.__getitem__ (/root/reference/data/loader.py:81-141; semeval_loader.py:75-119 for 9-tuples) the way
the reference does it: Python lists, ``sorted(zip(...), reverse=True)`` (sort_all, loader.py:176-180), per-token
``np.random.random()`` word dropout (loader.py:181-188), ``get_long_tensor`` padding (loader.py:167-174) with fill
150 for the two position fields (loader.py:125-126).  Pinned: tests/test_loader.py checks it against batches recorded
from the unmodified reference loader (tests/golden/loader.npz, made by tests/golden/make_loader_golden.py).
"""
import numpy as np
import torch

PAD_ID, UNK_ID = 0, 1


def get_long_tensor(tokens_list, batch_size, fill_value=PAD_ID):
    """loader.py:167-174."""
    width = max(len(x) for x in tokens_list)
    out = torch.LongTensor(batch_size, width).fill_(fill_value)
    for i, s in enumerate(tokens_list):
        out[i, :len(s)] = torch.LongTensor(s)
    return out


def sort_all(batch, lens):
    """loader.py:176-180: every field sorted by descending length; ties fall to the later original index."""
    order = [i for _, i in sorted(zip(lens, range(len(lens))), reverse=True)]
    return [[field[i] for i in order] for field in batch], order


def word_dropout(tokens, rate):
    """loader.py:181-188: one numpy draw per token that is not already <UNK>."""
    return [UNK_ID if x != UNK_ID and np.random.random() < rate else x for x in tokens]


def get_batch(examples, evaluation, word_dropout_rate, with_ner=True):
    """examples: list of (words, pos, ner, deprel, head, subj_pos, obj_pos, relation) id lists (ner None when the
    dataset has no NER field) -> the reference's batch tuple of CPU tensors."""
    n = len(examples)
    fields = [0, 1, 2, 3, 4, 5, 6] if with_ner else [0, 1, 3, 4, 5, 6]
    batch = [[e[f] for e in examples] for f in fields]
    rels = [e[7] for e in examples]
    lens = [len(x) for x in batch[1]]
    (*batch, rels), orig_idx = sort_all(batch + [rels], lens)
    words = batch[0] if evaluation else [word_dropout(s, word_dropout_rate) for s in batch[0]]
    words = get_long_tensor(words, n)
    masks = torch.eq(words, 0)
    rest = [get_long_tensor(f, n) for f in batch[1:-2]]
    subj = get_long_tensor(batch[-2], n, fill_value=150)
    obj = get_long_tensor(batch[-1], n, fill_value=150)
    return (words, masks, *rest, subj, obj, torch.LongTensor(rels), orig_idx)
