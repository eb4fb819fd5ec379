#!/usr/bin/env python
"""Timeline of ONE replay of the fused training step's CUDA graph (torch.profiler / CUPTI activity records): start
offset, duration and stream of every kernel as it actually ran -- warm caches, overlapping branches, launch gaps --
which the serialised, cold-cache ncu launch list cannot show."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcn_over_pruned_trees_b200 import parallel, synth  # noqa: E402
from gcn_over_pruned_trees_b200.engine import FusedTrainStep, PackedBatch  # noqa: E402
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer  # noqa: E402


def main():
    # under torchrun: the data-parallel step (K8 exchange over NVLink peer memory); rank 0 prints its own timeline
    rank, local_rank, world = parallel.init_from_env('nccl')
    torch.cuda.set_device(local_rank)
    torch.manual_seed(0)
    sys.stdout = open(os.devnull, 'w') if rank else sys.stdout
    tr = GCNTrainer(synth.tacred_opt(vocab_size=50000, cuda=True, gemm_mode='tf32x3', prune_k=1))
    tr.model.train()
    eng = FusedTrainStep(tr, data_parallel=world > 1, max_rows=6400)
    if eng.exchange is not None:
        print('exchange: regions in %s, push through %s' % (
            'torch symmetric memory' if eng.exchange.symm is not None else 'cudaMalloc + cudaIpc',
            'NVSwitch multicast (multimem.st)' if eng.exchange.multicast else 'one store per peer'))
    batches = [PackedBatch(parallel.shard_batch(synth.make_batch(1000 + i, batch_size=50 * world, vocab_size=50000), rank,
                                                world), device='cpu').to('cuda') for i in range(4)]
    flush = torch.empty(64 << 20, dtype=torch.float32, device='cuda')
    for _ in range(6):
        for b in batches:
            eng(b)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(4):
            flush.fill_(0.0)
            eng(batches[i])
        torch.cuda.synchronize()
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA],
                 key=lambda e: e.time_range.start)
    # split into replays at the flush kernels; print the last replay
    cuts = [i for i, e in enumerate(evs) if 'fill' in e.name.lower() or 'FillFunctor' in e.name]
    last = evs[cuts[-1] + 1:]
    last = [e for e in last if 'Memcpy' not in e.name or True]
    t0 = last[0].time_range.start
    print('%-58s %9s %9s' % ('kernel (last replay, L2 flushed before it)', 'start us', 'dur us'))
    end = 0.0
    for e in last:
        s = e.time_range.start - t0
        d = e.time_range.end - e.time_range.start
        end = max(end, s + d)
        print('%-58s %9.1f %9.1f' % (e.name.replace('(anonymous namespace)::', '')[:58], s, d))
    print('replay span %.1f us, sum of kernel durations %.1f us (world %d)' % (
        end, sum(e.time_range.end - e.time_range.start for e in last), world))
    if world > 1:
        parallel.barrier()
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
