"""Device-resident data loader: the reference's ``data.loader.DataLoader`` interface, batches assembled on the GPU.

Same constructor, ``len()``, ``[i]``, iteration and ``gold()`` as /root/reference/data/loader.py:13-141 (TACRED
10-tuples) and /root/reference/data/semeval_loader.py:13-119 (9-tuples without NER, ``dataset='semeval'``), so
``train.py:79-81`` / ``eval.py:50`` keep working.  What changes is where the work happens:

* construction: json -> ids exactly as ``preprocess`` (loader.py:43-72: optional lower-casing, entity tokens masked as
  ``SUBJ-<type>`` / ``OBJ-<type>`` for TACRED, ``<UNK>`` for unknown words and tags, position sequences of ``get_positions``
  loader.py:162-165), shuffled with ``random.shuffle`` like the reference (same permutation for the same seed), then
  the whole corpus is uploaded ONCE as an int32 token arena together with every batch's length-sorted sentence list
  (``sort_all`` loader.py:176-180: descending length, ties by descending position in the batch);
* ``loader[i]``: one launch of K9 (csrc/batch.cu) writes the padded int64 fields, the pad mask and the labels into a
  ``PackedBatch`` on the device.  No Python loop over tokens, no host-to-device copy per step; the tuple it returns
  holds CUDA tensors, which ``unpack_batch`` / ``FusedTrainStep`` take as they are.

Word dropout (train mode, loader.py:181-188) is drawn on the device from Philox by default.  ``host_word_dropout=True``
instead replays the reference's own ``np.random.random()`` stream on the host (one draw per non-<UNK> token, in batch
order) and uploads the affected batch's word ids, which makes the batches bit-identical to the reference's for a given
``np.random.seed`` -- used by the parity tests.  CUDA only: there is no CPU fallback.
"""
import json
import os
import random

import numpy as np
import torch

try:                                    # imported as gcn_over_pruned_trees_b200.data.loader
    from .. import _lib, constant, ops
    from ..engine import PackedBatch
except ImportError:                     # imported as top-level `data.loader` (package dir on PYTHONPATH: train.py:21)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from gcn_over_pruned_trees_b200 import _lib, constant, ops
    from gcn_over_pruned_trees_b200.engine import PackedBatch

SEMEVAL_LABELS = ['Other', 'Entity-Destination', 'Cause-Effect', 'Member-Collection', 'Entity-Origin',
                  'Message-Topic', 'Component-Whole', 'Instrument-Agency', 'Product-Producer', 'Content-Container']
SEMEVAL_LABEL_TO_ID = {name: i for i, name in enumerate(SEMEVAL_LABELS)}     # utils/constant_semeval.py:29
POSITION_FILL = 150                                                           # loader.py:125-126


def map_to_ids(tokens, vocab):
    """loader.py:158-160: unknown entries map to <UNK>."""
    return [vocab.get(t, constant.UNK_ID) for t in tokens]


def get_positions(start_idx, end_idx, length):
    """loader.py:162-165: ..., -2, -1, 0, ..., 0, 1, 2, ... around the span [start_idx, end_idx]."""
    return list(range(-start_idx, 0)) + [0] * (end_idx - start_idx + 1) + list(range(1, length - end_idx))


def preprocess(data, word2id, opt, label2id, with_ner=True):
    """json examples -> (words, pos, ner, deprel, head, subj_pos, obj_pos, relation) id lists (loader.py:43-72)."""
    out = []
    for d in data:
        tokens = list(d['token'])
        if opt['lower']:
            tokens = [t.lower() for t in tokens]
        ss, se, os_, oe = d['subj_start'], d['subj_end'], d['obj_start'], d['obj_end']
        if with_ner:     # TACRED only: entity tokens are anonymised (loader.py:52-56; semeval_loader.py:48-55 keeps them)
            tokens[ss:se + 1] = ['SUBJ-' + d['subj_type']] * (se - ss + 1)
            tokens[os_:oe + 1] = ['OBJ-' + d['obj_type']] * (oe - os_ + 1)
        head = [int(x) for x in d['stanford_head']]
        if not any(x == 0 for x in head):
            raise AssertionError('sentence %r has no root' % d.get('id'))        # loader.py:64
        n = len(tokens)
        out.append((map_to_ids(tokens, word2id), map_to_ids(d['stanford_pos'], constant.POS_TO_ID),
                    map_to_ids(d['stanford_ner'], constant.NER_TO_ID) if with_ner else None,
                    map_to_ids(d['stanford_deprel'], constant.DEPREL_TO_ID), head,
                    get_positions(ss, se, n), get_positions(os_, oe, n), label2id[d['relation']]))
    return out


def sorted_rows(lens):
    """Row order of a batch after sort_all (loader.py:176-180): sorted((len, index, ...), reverse=True)."""
    return sorted(range(len(lens)), key=lambda i: (lens[i], i), reverse=True)


class DataLoader(object):
    """Load data from a json file, keep it on the GPU, emit the reference's batches."""

    def __init__(self, filename, batch_size, opt, vocab, evaluation=False, bert_embeddings=None, dataset=None,
                 device=None, host_word_dropout=False, seed=None, _processed=None):
        if bert_embeddings is not None:         # train.py:79-84 passes the keyword (None unless --use_bert_embeddings)
            raise NotImplementedError('pre-computed BERT token vectors (data/loader.py:90-104) are outside the built '
                                      'path (SURVEY.md section 2: out of scope); use the reference loader for them')
        self.batch_size = batch_size
        self.opt = opt
        self.vocab = vocab
        self.eval = evaluation
        dataset = dataset or opt.get('dataset', 'tacred')
        self.tacred = dataset == 'tacred'
        self.label2id = constant.LABEL_TO_ID if self.tacred else SEMEVAL_LABEL_TO_ID
        self.host_word_dropout = host_word_dropout
        if _processed is None:
            with open(filename) as infile:
                data = json.load(infile)
            self.raw_data = data
            data = preprocess(data, vocab.word2id, opt, self.label2id, with_ner=self.tacred)
        else:
            self.raw_data = None
            data = list(_processed)
        if not evaluation:                                  # shuffle for training (loader.py:31-34)
            indices = list(range(len(data)))
            random.shuffle(indices)
            data = [data[i] for i in indices]
        self.id2label = {v: k for k, v in self.label2id.items()}
        self.labels = [self.id2label.get(d[-1], d[-1]) for d in data]
        self.num_examples = len(data)
        self._plan(data)
        self._upload(data, device)
        # (not drawn from `random` / `np.random`: their streams stay exactly as the reference loader leaves them)
        self.seed = int.from_bytes(os.urandom(8), 'little') if seed is None else int(seed)
        self.draws = 0                                      # batches emitted so far: the dropout stream id
        print("{} batches created for {}".format(len(self.batches), filename))

    @classmethod
    def from_processed(cls, examples, batch_size, opt, evaluation=False, **kw):
        """Loader over already tokenised examples: (words, pos, ner | None, deprel, head, subj_pos, obj_pos, label id)
        id lists, i.e. what ``preprocess`` returns (synthetic corpora: bench.py, tests)."""
        return cls('<%d processed examples>' % len(examples), batch_size, opt, None, evaluation=evaluation,
                   _processed=examples, **kw)

    # ---- host side: what the reference computes per batch, computed once ------------------------------------------
    def _plan(self, data):
        self.lens = [len(d[0]) for d in data]
        self.batches = []                                   # (first sentence, size, T, sorted sentence ids, orig_idx)
        for i in range(0, len(data), self.batch_size):
            lens = self.lens[i:i + self.batch_size]
            rows = sorted_rows(lens)
            self.batches.append((i, len(lens), max(lens), [i + r for r in rows], rows))
        self._host_words = [d[0] for d in data] if (self.host_word_dropout and not self.eval) else None

    def _upload(self, data, device):
        if device is None:
            if not torch.cuda.is_available():
                raise _lib.GptError('the device-resident loader needs a CUDA device: there is no CPU fallback')
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(device)
        fields = [0, 1, 2, 3, 4, 5, 6] if self.tacred else [0, 1, 3, 4, 5, 6]
        offsets = np.zeros(len(data) + 1, dtype=np.int64)
        np.cumsum(self.lens, out=offsets[1:])
        self.arena = [None] * 7
        for f in fields:
            flat = np.fromiter((v for d in data for v in d[f]), dtype=np.int32, count=int(offsets[-1]))
            self.arena[f] = torch.from_numpy(flat).to(self.device)
        self.offsets = torch.from_numpy(offsets).to(self.device)
        self.label_ids = torch.tensor([d[-1] for d in data], dtype=torch.int32, device=self.device)
        sel = np.zeros((max(len(self.batches), 1), self.batch_size), dtype=np.int32)
        for k, (_, n, _, ids, _) in enumerate(self.batches):
            sel[k, :n] = ids
        self.sel = torch.from_numpy(sel).to(self.device)
        import ctypes
        self._arena_arr = (ctypes.c_void_p * 7)(*[None if t is None else t.data_ptr() for t in self.arena])
        self._offsets_ptr, self._labels_ptr = self.offsets.data_ptr(), self.label_ids.data_ptr()
        self._sel_ptr = self.sel.data_ptr()
        self._rings, self._args = {}, {}

    # ---- the reference's interface ------------------------------------------------------------------------------
    def gold(self):
        """Gold labels as a list (loader.py:74-76)."""
        return self.labels

    def __len__(self):
        return len(self.batches)

    RING = 4        # batches of one shape handed out before a buffer is reused

    def _target(self, n, T, out):
        """(PackedBatch, cached ctypes arguments) to write a batch of n sentences x T tokens into: ``out`` if given,
        else the next buffer of this shape's ring."""
        import ctypes
        if out is None:
            ring = self._rings.setdefault((n, T), [[], 0])
            if len(ring[0]) < self.RING:
                ring[0].append(PackedBatch(like=_Shape(n, T, 8 if self.tacred else 7), device=self.device))
                pb = ring[0][-1]
            else:
                pb = ring[0][ring[1] % self.RING]
                ring[1] += 1
        else:
            pb = out
            if pb.key != (n, T, 8 if self.tacred else 7) or pb.buf.device != self.device:
                raise _lib.GptError('target batch has layout %r on %s, need %r on %s'
                                    % (pb.key, pb.buf.device, (n, T, 8 if self.tacred else 7), self.device))
        args = self._args.get(id(pb))
        if args is None or args[0] is not pb:
            fields = pb.fields                              # words, masks, pos, [ner,] deprel, head, subj_pos, obj_pos
            ptrs = [None] * 7
            ptrs[0] = fields[0].data_ptr()
            for f, t in zip([1, 2, 3, 4, 5, 6] if self.tacred else [1, 3, 4, 5, 6], fields[2:]):
                ptrs[f] = t.data_ptr()
            args = (pb, (ctypes.c_void_p * 7)(*ptrs), fields[1].data_ptr(), pb.labels.data_ptr())
            self._args[id(pb)] = args
        return args

    def packed(self, key, out=None):
        """Batch ``key`` as a device-resident PackedBatch: one launch of K9, no host-to-device copy.  ``out``: write
        into that PackedBatch (e.g. the static input buffer of a captured step) instead of a buffer of the loader's
        own ring -- a ring buffer is valid until RING more batches of the same shape have been requested."""
        if not isinstance(key, int):
            raise TypeError
        if key < 0 or key >= len(self.batches):
            raise IndexError
        first, n, T, ids, rows = self.batches[key]
        pb, out_arr, masks_ptr, rels_ptr = self._target(n, T, out)
        p = 0.0 if (self.eval or self.host_word_dropout) else float(self.opt['word_dropout'])
        ops.build_batch_raw(self._arena_arr, self._offsets_ptr, self._labels_ptr,
                            self._sel_ptr + 4 * self.batch_size * key, n, T, p, self.seed, self.draws, out_arr,
                            masks_ptr, rels_ptr)
        self.draws += 1
        if self._host_words is not None:                    # the reference's numpy stream, replayed on the host
            rate = self.opt['word_dropout']
            words = np.zeros((n, T), dtype=np.int64)
            for r, s in enumerate(ids):
                sent = self._host_words[s]
                words[r, :len(sent)] = [constant.UNK_ID if x != constant.UNK_ID and np.random.random() < rate else x
                                        for x in sent]
            pb.fields[0].copy_(torch.from_numpy(words))
        pb.orig_idx = rows
        return pb

    def __getitem__(self, key):
        """(words, masks, pos, ner, deprel, head, subj_pos, obj_pos, rels, orig_idx) -- CUDA tensors (loader.py:140).
        The tuple owns its storage (it is not one of the ring buffers)."""
        if not isinstance(key, int):
            raise TypeError
        if key < 0 or key >= len(self.batches):
            raise IndexError
        _, n, T, _, rows = self.batches[key]
        own = PackedBatch(like=_Shape(n, T, 8 if self.tacred else 7), device=self.device)
        self.packed(key, out=own)
        self._args.pop(id(own), None)
        return own.as_tuple()[:-1] + (list(rows),)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class _Shape(object):
    """Just enough of a PackedBatch to allocate another one of the same layout."""

    def __init__(self, B, T, n_fields):
        self.key = (B, T, n_fields)
        self.dtypes = [torch.int64, torch.bool] + [torch.int64] * (n_fields - 2)
