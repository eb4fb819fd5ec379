"""Deterministic parameter sets in the reference's checkpoint layout ("Checkpoint" row).

Golden fixtures store only a seed, never the weights: both ``make_golden.py`` (which loads them into the real
reference model) and the tests (which load them into the oracle and into the CUDA model) rebuild the same
``state_dict`` from the seed with numpy's PCG64 stream.
"""
import zlib

import numpy as np

N_POS, N_NER, N_DEPREL = 47, 15, 85


def state_shapes(opt):
    """{key: shape} of GCNClassifier.state_dict() (incl. the duplicated embedding keys); adj_type 'regular',
    'full_deprel' (one shared Linear, gcn.py:164-167) or 'diagonal_deprel' (a preprocessor, no W, gcn.py:153-155)."""
    tacred = opt['dataset'] == 'tacred'
    hidden = opt['hidden_dim']
    shapes = {'gcn_model.emb.weight': (opt['vocab_size'], opt['emb_dim'])}
    if opt['pos_dim'] > 0:
        shapes['gcn_model.pos_emb.weight'] = (N_POS, opt['pos_dim'])
    if opt['ner_dim'] > 0:
        shapes['gcn_model.ner_emb.weight'] = (N_NER, opt['ner_dim'])
    adj_type = opt.get('adj_type', 'regular')
    side = {'regular': 1, 'diagonal_deprel': hidden}.get(adj_type, opt.get('deprel_emb_dim', 1))
    shapes['gcn_model.deprel_emb.weight'] = (N_DEPREL, side)
    width = opt['emb_dim'] + opt['pos_dim'] + (opt['ner_dim'] if tacred else 0)
    if opt.get('rnn', False):
        rh = opt['rnn_hidden']
        for layer in range(opt['rnn_layers']):
            fan = width if layer == 0 else 2 * rh
            for suffix in ('', '_reverse'):
                shapes['gcn_model.gcn.rnn.weight_ih_l%d%s' % (layer, suffix)] = (4 * rh, fan)
                shapes['gcn_model.gcn.rnn.weight_hh_l%d%s' % (layer, suffix)] = (4 * rh, rh)
                shapes['gcn_model.gcn.rnn.bias_ih_l%d%s' % (layer, suffix)] = (4 * rh,)
                shapes['gcn_model.gcn.rnn.bias_hh_l%d%s' % (layer, suffix)] = (4 * rh,)
        width = 2 * rh
    if adj_type == 'diagonal_deprel':
        shapes['gcn_model.gcn.preprocessor.weight'] = (hidden, width)
        shapes['gcn_model.gcn.preprocessor.bias'] = (hidden,)
    elif adj_type == 'full_deprel':
        shapes['gcn_model.gcn.W.weight'] = (side * hidden, width)
        shapes['gcn_model.gcn.W.bias'] = (side * hidden,)
    else:
        for layer in range(opt['num_layers']):
            shapes['gcn_model.gcn.W.%d.weight' % layer] = (hidden, width if layer == 0 else hidden)
            shapes['gcn_model.gcn.W.%d.bias' % layer] = (hidden,)
    shapes['gcn_model.out_mlp.0.weight'] = (hidden, 3 * hidden)
    shapes['gcn_model.out_mlp.0.bias'] = (hidden,)
    for i in range(1, opt['mlp_layers']):
        shapes['gcn_model.out_mlp.%d.weight' % (2 * i)] = (hidden, hidden)
        shapes['gcn_model.out_mlp.%d.bias' % (2 * i)] = (hidden,)
    shapes['classifier.weight'] = (opt['num_class'], hidden)
    shapes['classifier.bias'] = (opt['num_class'],)
    return shapes


def make_state(opt, seed):
    """{key: float32 ndarray}; embeddings ~U(-1,1) (gcn.py:74-75), the rest ~U(-1/sqrt(fan_in), 1/sqrt(fan_in))."""
    state = {}
    for key, shape in state_shapes(opt).items():
        rng = np.random.default_rng([seed, zlib.crc32(key.encode())])
        if 'deprel_emb.weight' in key and opt.get('adj_type', 'regular') == 'full_deprel':
            bound = 1.5 / np.sqrt(shape[1])               # relation vectors: keeps the sum over D terms O(1)
        elif 'emb.weight' in key:
            bound = 1.0
        elif len(shape) == 2:
            bound = 1.0 / np.sqrt(shape[1])
        else:
            bound = 1.0 / np.sqrt(opt['hidden_dim'])
        w = rng.uniform(-bound, bound, size=shape).astype(np.float32)
        if key == 'gcn_model.emb.weight' or key == 'gcn_model.deprel_emb.weight':
            w[0] = 0.0                                    # padding_idx rows
        state[key] = w
    for name in ('emb', 'pos_emb', 'ner_emb', 'deprel_emb'):     # shared tables appear under both modules
        src = 'gcn_model.%s.weight' % name
        if src in state:
            state['gcn_model.gcn.%s.weight' % name] = state[src]
    return state


def grad_digest(g, limit=2048):
    """Compact fingerprint of a gradient tensor: strided sample + l2 norm + sum (float64 reductions)."""
    flat = np.asarray(g, dtype=np.float32).reshape(-1)
    stride = max(1, flat.size // limit)
    return flat[::stride].copy(), np.float64(np.sqrt((flat.astype(np.float64) ** 2).sum())), \
        np.float64(flat.astype(np.float64).sum())
