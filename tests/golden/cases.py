"""Case tables shared by make_golden.py (reference side) and the tests (oracle / CUDA side)."""
import numpy as np
import torch

PRUNE_KS = (-1, 0, 1, 2, 3)
SPLITS = ('train', 'dev', 'test')
SYNTH_ADJ_SEEDS = tuple(range(100, 108))           # 8 batches x 50 sentences x len(PRUNE_KS)

# name -> (opt overrides, batch source, weight seed).  Batch source: ('split', name) = bundled sample JSON,
# ('synth', seed, batch_size) = gcn_over_pruned_trees_b200.synth.make_batch.
SMALL_VOCAB = 400
MODEL_CASES = {
    'cfg1_train_json_k1':   (dict(prune_k=1), ('split', 'train'), 11),
    'cfg2_synth_kfull':     (dict(prune_k=-1, vocab_size=SMALL_VOCAB), ('synth', 201, 50), 12),
    'cfg2_synth_k0':        (dict(prune_k=0, vocab_size=SMALL_VOCAB), ('synth', 202, 50), 12),
    'cfg2_synth_k1':        (dict(prune_k=1, vocab_size=SMALL_VOCAB), ('synth', 203, 50), 12),
    'cfg2_synth_k2':        (dict(prune_k=2, vocab_size=SMALL_VOCAB), ('synth', 204, 50), 12),
    'cfg3_cgcn_k1':         (dict(prune_k=1, rnn=True, vocab_size=SMALL_VOCAB), ('synth', 205, 50), 13),
    'cfg4_semeval_k1':      (dict(prune_k=1, dataset='semeval', num_class=19, vocab_size=SMALL_VOCAB),
                             ('synth', 206, 50), 14),
    'avg_pool_conv_l2':     (dict(prune_k=1, pooling='avg', conv_l2=0.01, vocab_size=SMALL_VOCAB),
                             ('synth', 207, 24), 15),
    'sum_pool_3layer_mlp1': (dict(prune_k=2, pooling='sum', num_layers=3, mlp_layers=1, pooling_l2=0.0,
                                  vocab_size=SMALL_VOCAB), ('synth', 208, 24), 16),
    'no_adj_1layer':        (dict(prune_k=1, no_adj=True, num_layers=1, vocab_size=SMALL_VOCAB),
                             ('synth', 209, 24), 17),
    'hidden64_k0':          (dict(prune_k=0, hidden_dim=64, emb_dim=50, pos_dim=5, ner_dim=0,
                                  vocab_size=SMALL_VOCAB), ('synth', 210, 16), 18),
}
# relation-aware adjacency modes.  full_deprel shares ONE Linear(in, D*H) between the
# layers, so the reference only runs when the GCN input width equals hidden_dim.
_IN64 = dict(emb_dim=40, pos_dim=12, ner_dim=12, hidden_dim=64, vocab_size=SMALL_VOCAB)
DEPREL_CASES = {
    'full_k1_d8':          (dict(_IN64, adj_type='full_deprel', deprel_emb_dim=8, prune_k=1), ('synth', 301, 24), 21),
    'full_kfull_d16':      (dict(_IN64, adj_type='full_deprel', deprel_emb_dim=16, prune_k=-1), ('synth', 302, 16), 22),
    'full_directed':       (dict(_IN64, adj_type='full_deprel', deprel_emb_dim=8, prune_k=2, deprel_directed=True),
                            ('synth', 303, 16), 23),
    'full_no_self_loop':   (dict(_IN64, adj_type='full_deprel', deprel_emb_dim=8, prune_k=1, deprel_self_loop=False),
                            ('synth', 304, 16), 24),
    'full_depth1_3layer':  (dict(_IN64, adj_type='full_deprel', deprel_emb_dim=8, prune_k=1, num_layers=3,
                                 deprel_max_depth=1), ('synth', 305, 16), 25),
    'full_cgcn_h64':       (dict(_IN64, adj_type='full_deprel', deprel_emb_dim=8, prune_k=1, rnn=True, rnn_hidden=32),
                            ('synth', 306, 16), 26),
    'full_semeval':        (dict(_IN64, adj_type='full_deprel', deprel_emb_dim=8, prune_k=1, dataset='semeval',
                                 emb_dim=52, num_class=19), ('synth', 307, 16), 27),
    'full_split_train':    (dict(_IN64, adj_type='full_deprel', deprel_emb_dim=12, prune_k=1), ('split', 'train'), 28),
    'diag_k1':             (dict(adj_type='diagonal_deprel', prune_k=1, hidden_dim=64, vocab_size=SMALL_VOCAB),
                            ('synth', 311, 24), 31),
    'diag_kfull_3layer':   (dict(adj_type='diagonal_deprel', prune_k=-1, hidden_dim=48, num_layers=3,
                                 vocab_size=SMALL_VOCAB), ('synth', 312, 16), 32),
    'diag_cgcn':           (dict(adj_type='diagonal_deprel', prune_k=1, hidden_dim=64, rnn=True, rnn_hidden=40,
                                 vocab_size=SMALL_VOCAB), ('synth', 313, 16), 33),
}
# batches whose sentence 0 gets a second root AFTER the first one, over an entity-free subtree: for prune_k < 0 the
# reference keeps only the LAST root's component (tree.py:76-77), so the entity tokens are outside the tree (empty
# adjacency rows) while the subject / object pools still read them -- in these modes they carry relu(self loop)
DEPREL_CASES['full_entities_outside_tree'] = (
    dict(_IN64, adj_type='full_deprel', deprel_emb_dim=8, prune_k=-1), ('synth_second_root', 308, 12), 29)
DEPREL_CASES['diag_entities_outside_tree'] = (
    dict(adj_type='diagonal_deprel', prune_k=-1, hidden_dim=64, vocab_size=SMALL_VOCAB), ('synth_second_root', 314, 12), 34)
DEPREL_GRAD_CASES = ('full_k1_d8', 'full_directed', 'full_depth1_3layer', 'full_cgcn_h64', 'diag_k1')
# train-mode cases of the reference with edge dropout / relation forgetting drawn from torch.manual_seed(DROPOUT_SEED)
DEPREL_RANDOM_CASES = {
    'full_edge_drop':      (dict(_IN64, adj_type='full_deprel', deprel_emb_dim=8, prune_k=1, edge_keep_prob=0.7),
                            ('synth', 321, 16), 41),
    'full_forget':         (dict(_IN64, adj_type='full_deprel', deprel_emb_dim=8, prune_k=1, deprel_keep_prop=0.6),
                            ('synth', 322, 16), 42),
}

# cases that additionally store train-mode (dropout drawn from torch.manual_seed(DROPOUT_SEED)) loss + gradients
GRAD_CASES = ('cfg1_train_json_k1', 'cfg2_synth_k1', 'cfg3_cgcn_k1', 'cfg4_semeval_k1', 'avg_pool_conv_l2',
              'sum_pool_3layer_mlp1')
DROPOUT_SEED = 777

# hand-written trees: (head, subj token ids, obj token ids, deprel) -- heads are 1-based, 0 = root
EDGE_TREES = {
    'chain5':              ([0, 1, 2, 3, 4], [4], [0], [11, 2, 3, 4, 5]),
    'subj_is_ancestor':    ([0, 1, 2, 3, 4], [1], [4], [11, 2, 3, 4, 5]),
    'same_token':          ([2, 0, 2, 3], [3], [3], [5, 11, 6, 7]),            # singleton tree -> all-zero adj
    'star':                ([0, 1, 1, 1, 1, 1, 1], [2], [5], [11, 2, 3, 4, 5, 6, 7]),
    'multi_token_spans':   ([2, 0, 2, 3, 3, 5, 5, 2], [3, 4], [6, 7], [7, 11, 10, 3, 4, 5, 6, 2]),
    'two_roots_ent_in_2nd': ([0, 1, 0, 3, 4, 3], [4], [5], [11, 2, 11, 4, 5, 6]),
    'two_roots_ent_in_1st': ([0, 1, 1, 0, 4], [1], [2], [11, 2, 3, 11, 5]),
    'len1':                ([0], [0], [0], [11]),
    'len2':                ([2, 0], [0], [1], [7, 11]),
    'deep_left_comb':      ([2, 3, 4, 5, 6, 7, 8, 0], [0], [7], [2, 3, 4, 5, 6, 7, 8, 11]),
    'obj_empty':           ([0, 1, 2, 2], [3], [], [11, 2, 3, 4]),             # reference runs with S only
    'deprel_pad_id':       ([0, 1, 1, 2], [3], [2], [11, 0, 5, 0]),            # deprel 0: forward entry is 0
    'wide_and_deep':       ([3, 3, 0, 3, 4, 4, 6, 6, 8, 8, 10, 10], [11], [6], [2, 3, 11, 4, 5, 6, 7, 8, 9, 10, 12, 13]),
}


def add_second_root(batch, row=0):
    """Loader batch with sentence ``row`` given a second root: the token with the largest index that lies after the
    root, has at least one child and no entity token in its subtree gets head 0."""
    tacred = len(batch) >= 10
    off = 5 if tacred else 4
    head = batch[off].clone()
    subj_pos, obj_pos = batch[off + 1], batch[off + 2]
    n = int((~batch[1][row]).sum())
    h = head[row, :n].tolist()
    entity = [subj_pos[row, t].item() == 0 or obj_pos[row, t].item() == 0 for t in range(n)]
    children = [[] for _ in range(n)]
    for t, p in enumerate(h):
        if p > 0:
            children[p - 1].append(t)
    root = h.index(0)

    def subtree(v):
        out, stack = [], [v]
        while stack:
            u = stack.pop()
            out.append(u)
            stack.extend(children[u])
        return out

    for v in range(n - 1, root, -1):
        sub = subtree(v)
        if len(sub) >= 2 and not any(entity[u] for u in sub):
            head[row, v] = 0
            return batch[:off] + (head,) + batch[off + 1:]
    raise ValueError('sentence %d has no entity-free subtree after its root' % row)


def make_case_batch(source, over, golden_adj=None):
    """The batch a (DEPREL_)MODEL_CASES source describes."""
    from gcn_over_pruned_trees_b200 import synth
    if source[0] == 'split':
        return batch_from_npz(golden_adj, source[1])
    batch = synth.make_batch(source[1], batch_size=source[2], vocab_size=over['vocab_size'],
                             num_class=over.get('num_class', 42), dataset=over.get('dataset', 'tacred'))
    return add_second_root(batch) if source[0] == 'synth_second_root' else batch


def positions(span_tokens, length, fill=150, width=None):
    """subj_pos / obj_pos row for a (possibly non-contiguous) token set: 0 inside, nonzero elsewhere."""
    width = length if width is None else width
    row = np.full(width, fill, dtype=np.int64)
    row[:length] = np.arange(1, length + 1)
    for t in span_tokens:
        row[t] = 0
    return row


def bundled_vocab(raw_splits):
    """Stub vocab over the bundled sample (entity tokens anonymised as loader.py:49-53 does)."""
    words = set()
    for data in raw_splits.values():
        for d in data:
            toks = list(d['token'])
            toks[d['subj_start']:d['subj_end'] + 1] = ['SUBJ-' + d['subj_type']] * (d['subj_end'] - d['subj_start'] + 1)
            toks[d['obj_start']:d['obj_end'] + 1] = ['OBJ-' + d['obj_type']] * (d['obj_end'] - d['obj_start'] + 1)
            words.update(toks)
    return ['<PAD>', '<UNK>'] + sorted(words)


def batch_from_npz(z, prefix):
    """Rebuild the loader 10-tuple stored by make_golden.py under ``prefix``."""
    def g(name):
        return torch.from_numpy(z['%s/%s' % (prefix, name)].astype(np.int64))
    words = g('words')
    return (words, words.eq(0), g('pos'), g('ner'), g('deprel'), g('head'), g('subj_pos'), g('obj_pos'),
            g('rels'), [int(i) for i in z['%s/orig_idx' % prefix]])
