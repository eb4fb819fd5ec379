"""CPU, world_size 2, gloo: the data-parallel plumbing (sentence sharding + gradient all-reduce).

The CUDA model has no CPU path, so the replicas here are the oracle's dense CPU model; what is under test is
gcn_over_pruned_trees_b200.parallel: after reduce(), every rank holds the gradients of the *whole* batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import weights
from gcn_over_pruned_trees_b200 import parallel, synth
from oracle import gcn_oracle


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _opt(dataset):
    return synth.tacred_opt(vocab_size=1200, prune_k=1, input_dropout=0.0, gcn_dropout=0.0, dataset=dataset,
                            num_class=19 if dataset == 'semeval' else 42)


def _grads(model):
    return {k: (None if p.grad is None else p.grad.clone()) for k, p in model.named_parameters()}


def _worker(rank, world, port, dataset, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = parallel.init_from_env('gloo')
    assert (r, w) == (rank, world)
    opt = _opt(dataset)
    torch.manual_seed(0)
    model = gcn_oracle.DenseClassifier(opt)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.make_state(opt, 3).items()})
    model.train()
    batch = synth.make_batch(77, batch_size=48, vocab_size=1200, dataset=dataset, num_class=opt['num_class'])
    shard = parallel.shard_batch(batch, rank, world)
    assert shard[0].shape[0] == 24 and len(shard[-1]) == 24
    loss, _ = model.loss(shard)
    loss.backward()
    reducer = parallel.GradAllReducer(model.parameters())
    reducer.reduce()
    t = parallel.max_over_ranks(float(rank + 1), torch.device('cpu'))
    assert t == float(world)
    parallel.barrier()
    torch.save({k: v for k, v in _grads(model).items()}, os.path.join(out_dir, 'grads_%d.pt' % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize('dataset', ('tacred', 'semeval'))
def test_allreduced_grads_equal_full_batch_grads(tmp_path, dataset):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), dataset, str(tmp_path)), nprocs=world, join=True)
    opt = _opt(dataset)
    model = gcn_oracle.DenseClassifier(opt)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.make_state(opt, 3).items()})
    model.train()
    batch = synth.make_batch(77, batch_size=48, vocab_size=1200, dataset=dataset, num_class=opt['num_class'])
    loss, _ = model.loss(batch)
    loss.backward()
    want = _grads(model)
    got = [torch.load(str(tmp_path / ('grads_%d.pt' % r))) for r in range(world)]
    n_none = 0
    for k, w in want.items():
        for g in got:
            if w is None:                       # deprel_emb (and ner_emb on SemEval) never receive gradients
                assert g[k] is None
                n_none += 1
                continue
            scale = max(float(w.abs().max()), 1e-12)
            assert float((g[k] - w).abs().max()) <= 2e-5 * scale, k
    assert n_none >= world                      # the None-gradient parameters were skipped consistently
    for k in want:
        if want[k] is not None:
            assert torch.equal(got[0][k], got[1][k])    # replicas stay bit-identical


def test_shard_batch_partitions_sentences():
    batch = synth.make_batch(5, batch_size=10)
    parts = [parallel.shard_batch(batch, r, 3) for r in range(3)]
    assert sorted(i for p in parts for i in p[-1]) == sorted(batch[-1])
    assert sum(p[0].shape[0] for p in parts) == 10
    assert torch.equal(parts[1][0], batch[0][1::3][:, :parts[1][0].shape[1]])


def test_single_process_is_a_no_op():
    opt = _opt('tacred')
    model = gcn_oracle.DenseClassifier(opt)
    loss, _ = model.loss(synth.make_batch(1, batch_size=4, vocab_size=1200))
    loss.backward()
    before = _grads(model)
    parallel.GradAllReducer(model.parameters()).reduce()
    for k, v in _grads(model).items():
        assert v is None or torch.equal(v, before[k])
