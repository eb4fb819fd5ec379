"""Worker of tests/test_gpu_multi.py: run under torchrun with one process per GPU.

Every rank trains on its shard (rows rank::world) of the same global batches through FusedTrainStep(data_parallel=True)
(K8: exchange over NVLink peer memory); rank 0 also trains a single-process replica on the full batches.  Checks, every
step: replicas bit-identical across ranks; parameters after the step equal to the full-batch step within 5e-6.

The two runs are compared in LOCKSTEP -- before every step all ranks load the single-process replica's parameters --
because free-running trajectories are not comparable at this tolerance: a pre-activation of the output MLP that is
zero to within rounding falls on either side of the ReLU under a 1e-7 perturbation, and that sentence's whole
contribution to the unit's gradient row flips with it (measured: 2e-7 relative noise on the parameters moves them by
up to 6e-3 after 8 steps at these sizes, tools/relu_knife_edge.py).  From identical parameters the
forward is bit-identical per sentence, so only the summation order of the gradient differs.  The exchange state
(step parity, slot tables, owner marks, gradient buffers) still carries over from step to step."""
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
    gemm = os.environ.get('GPT_DP_GEMM', 'fp32')          # 'tf32x3' = the mode bench.py runs
    over = dict(vocab_size=1500, cuda=True, input_dropout=0.0, gcn_dropout=0.0, gemm_mode=gemm)
    steps = 8
    batches = [synth.make_batch(300 + i, batch_size=50 * world, vocab_size=1500) for i in range(3)]
    torch.manual_seed(11)
    tr = GCNTrainer(synth.tacred_opt(**over))
    tr.model.train()
    eng = FusedTrainStep(tr, data_parallel=True, max_rows=8192)
    params = list(tr.model.parameters())
    sizes = [p.numel() for p in params]
    ref = ref_eng = None
    if rank == 0:
        torch.manual_seed(11)
        ref = GCNTrainer(synth.tacred_opt(**over))
        ref.model.train()
        ref_eng = FusedTrainStep(ref)
    ok, worst, loss_err, identical = True, 0.0, 0.0, True
    for s in range(steps):
        vec = torch.empty(sum(sizes), device='cuda')
        if rank == 0:
            vec.copy_(torch.cat([p.detach().reshape(-1) for p in ref.model.parameters()]))
        dist.broadcast(vec, 0)
        for p, chunk in zip(params, vec.split(sizes)):
            p.data.copy_(chunk.view_as(p))                    # in place: the step graphs keep their addresses
        shard = parallel.shard_batch(batches[s % 3], rank, world)
        loss = eng(shard).clone()
        torch.cuda.synchronize()
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        identical = identical and all(torch.equal(g, gathered[0]) for g in gathered)
        dist.all_reduce(loss)
        loss /= world
        if rank == 0:
            ref_loss = float(ref_eng(batches[s % 3]))
            for a, b in zip(params, ref.model.parameters()):
                worst = max(worst, float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)))
            loss_err = max(loss_err, abs(float(loss) - ref_loss) / abs(ref_loss))
    ok = identical
    if rank == 0:
        ok = ok and worst < 5e-6 and loss_err < 2e-5
        line = ('DP_CHECK %s world=%d steps=%d gemm=%s (lockstep) replicas bit-identical: %s; vs full-batch single process: '
                'params rel %.2e, loss rel %.2e' % ('OK' if ok else 'FAIL', world, steps, gemm, identical, worst, loss_err))
        print(line, flush=True)
        out_dir = os.path.join(REPO, 'gpurun_out')
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, 'dp_check_w%d_%s.txt' % (world, gemm)), 'w') as f:
                f.write(line + '\n')
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == '__main__':
    main()
