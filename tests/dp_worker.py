"""Worker of tests/test_gpu_multi.py: run under torchrun with one process per GPU.

Every rank trains on its shard (rows rank::world) of the same global batches through FusedTrainStep(data_parallel=True)
(K8: exchange over NVLink peer memory); rank 0 also trains a single-process replica on the full batches.  Checks:
replicas bit-identical across ranks; parameters equal to the full-batch run within 5e-5 (dropout off)."""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from gcn_over_pruned_trees_b200 import parallel, synth  # noqa: E402
from gcn_over_pruned_trees_b200.engine import FusedTrainStep  # noqa: E402
from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer  # noqa: E402


def main():
    rank, local_rank, world = parallel.init_from_env('nccl')
    torch.cuda.set_device(local_rank)
    over = dict(vocab_size=1500, cuda=True, input_dropout=0.0, gcn_dropout=0.0, gemm_mode='fp32')
    steps = 8
    batches = [synth.make_batch(300 + i, batch_size=50 * world, vocab_size=1500) for i in range(3)]
    torch.manual_seed(11)
    tr = GCNTrainer(synth.tacred_opt(**over))
    tr.model.train()
    eng = FusedTrainStep(tr, data_parallel=True, max_rows=8192)
    losses = []
    for s in range(steps):
        shard = parallel.shard_batch(batches[s % 3], rank, world)
        losses.append(float(eng(shard)))
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().reshape(-1) for p in tr.model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    identical = all(torch.equal(g, gathered[0]) for g in gathered)
    mean_loss = torch.tensor(losses, device='cuda')
    dist.all_reduce(mean_loss)
    mean_loss /= world
    ok = identical
    msg = 'replicas bit-identical: %s' % identical
    if rank == 0:
        torch.manual_seed(11)
        ref = GCNTrainer(synth.tacred_opt(**over))
        ref.model.train()
        ref_eng = FusedTrainStep(ref)
        ref_losses = [float(ref_eng(batches[s % 3])) for s in range(steps)]
        worst = 0.0
        for (k, a), (_, b) in zip(tr.model.state_dict().items(), ref.model.state_dict().items()):
            d = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
            worst = max(worst, d)
        loss_err = max(abs(a - b) / abs(b) for a, b in zip(mean_loss.tolist(), ref_losses))
        ok = ok and worst < 5e-5 and loss_err < 2e-5
        msg += '; vs full-batch single process: params rel %.2e, loss rel %.2e' % (worst, loss_err)
        line = 'DP_CHECK %s world=%d %s' % ('OK' if ok else 'FAIL', world, msg)
        print(line, flush=True)
        out_dir = os.path.join(REPO, 'gpurun_out')
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, 'dp_check_w%d.txt' % world), 'w') as f:
                f.write(line + '\n')
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == '__main__':
    main()
