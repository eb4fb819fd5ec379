#!/usr/bin/env python
"""bench.py -- GCN training throughput on synthetic TACRED-shaped data (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload = "tacred_b50_k1", BASELINE.json configs[1] at prune_k=1): batches of 50 synthetic
sentences per GPU (lengths clip(Poisson(36), 8, 96), uniform random recursive trees), 2-layer GCN, 360-d input,
200-d hidden, V=50000 random-init 300-d embeddings, dropout on.  One step = what /root/reference/train.py:213-227 does
per batch: zero_grad, GCNTrainer.update (forward + loss), backward, [gradient all-reduce when N > 1], global-norm
clip at 5.0, SGD step.

  value   sentences/s with the batches already resident in HBM (device-timed with CUDA events, max over ranks,
          L2 flushed between steps)
  e2e     same step driven through the reference-facing API from pinned HOST batches: H2D copy of the 9 batch
          tensors + loss.item() D2H inside the timed region (wall clock around each step, L2 flushed between)
  roofline      K2 aggregation forward at the large synthetic shape (BASELINE.json configs[4]: 512-token trees,
                B=4096, H=512 -- the only shape where an HBM fraction means anything), timed live
  cpu_baseline  the reference's own GCNTrainer (baseline/_ref, unmodified; kind "reference") timed on this box's host
                cores; the oracle's dense restatement (kind "port") only when baseline/_ref is not installed
  dp_check      (N > 1) lockstep correctness of the gradient exchange, run before the timed region: replicas
                bit-identical, parameters equal to a single-process step on the concatenated batch
  extra         BASELINE.json configs[4] (large512: 512-token trees, B=4096 per GPU, H=512; NCCL all-reduce when N > 1)
                and configs[3] (semeval_b50: 9-tuple batches, 19 classes) as sub-records, at every N
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = 'gcn_train_sentences_per_sec'
UNIT = 'sentences/s'
BATCH = 50
VOCAB = 50000
N_BATCHES = 16          # distinct synthetic batches per rank, cycled
L2_FLUSH_BYTES = 256 << 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--prune_k', type=int, default=1)
    ap.add_argument('--gemm', default='tf32x3', choices=['fp32', 'tf32x3', 'tf32'])
    ap.add_argument('--eager', action='store_true', help='time the eager five-call step instead of the CUDA graph')
    ap.add_argument('--autograd-engine', action='store_true',
                    help='use GraphedTrainStep (autograd under capture) instead of FusedTrainStep')
    ap.add_argument('--no-roofline', action='store_true', help='skip the large-shape aggregation roofline run')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-loader', action='store_true', help='skip the K9 device-loader leg (short profiler passes)')
    ap.add_argument('--roofline-batch', type=int, default=4096)
    ap.add_argument('--cpu-steps', type=int, default=60)
    ap.add_argument('--no-extra', action='store_true', help='skip the large512 / semeval_b50 legs')
    ap.add_argument('--no-dp-check', action='store_true')
    ap.add_argument('--large-batch', type=int, default=4096, help='sentences per GPU of the large512 leg')
    ap.add_argument('--large-steps', type=int, default=5)
    return ap.parse_args()


def workload_config(args, world):
    """What is computed -- identical for the b200 arm and the reference arm (the driver compares the two dicts)."""
    return {'workload': 'tacred_b50_k%d' % args.prune_k, 'batch_per_gpu': BATCH, 'global_batch': BATCH * world,
            'len': 'clip(Poisson(36),8,96)', 'layers': 2, 'in_dim': 360, 'hidden': 200, 'vocab': VOCAB,
            'prune_k': args.prune_k, 'parallelism': 'dp%d' % world,
            'sharding': 'rows r::N of one length-sorted global batch per step',
            'step': 'zero_grad+fwd+loss+bwd+allreduce+clip5+sgd',
            'l2': 'GPU arm: flushed between timed steps (256 MiB fill); CPU arm: not applicable'}


def engine_config(args, world):
    """How the b200 arm computes it (not part of `config`: the reference arm has no such keys)."""
    return {'gemm': args.gemm,
            'engine': 'eager' if args.eager else ('cuda_graph_autograd' if args.autograd_engine else
                                                  'cuda_graph_fused_step'),
            'exchange': 'none' if world == 1 else ('nccl_allreduce' if args.autograd_engine or args.eager else
                                                   'nvlink_peer_memory (K8)')}


# ------------------------------------------------------------------------------------------------ clocks -------

class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference ----

REF_DIR = os.path.join(REPO, 'baseline', '_ref')


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, 'model', 'trainer.py'))


def cpu_reference_run(args, steps, warmup):
    """The reference's own training step on the host cores.  kind "reference": the UNMODIFIED reference from
    baseline/_ref (tools/install_reference.py) -- its GCNTrainer, its head_to_tree / tree_to_adj, torch CPU kernels, the
    five calls of train.py:213-227.  kind "port" (only when baseline/_ref is absent): the oracle's restatement."""
    import torch
    from gcn_over_pruned_trees_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    opt = synth.tacred_opt(vocab_size=VOCAB, prune_k=args.prune_k, cuda=False)
    batches = [synth.make_batch(1000 + i, batch_size=BATCH, vocab_size=VOCAB) for i in range(min(N_BATCHES, 8))]
    if reference_available():
        import contextlib
        import io
        sys.path.insert(0, REF_DIR)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                from model import tree as ref_tree
                from model.trainer import GCNTrainer as RefTrainer
                if args.prune_k < 0:
                    ref_tree.Tree.head = None       # the one attribute the reference forgets
                trainer = RefTrainer(dict(opt))
        finally:
            sys.path.remove(REF_DIR)
        trainer.model.train()
        kind = 'reference'

        def step(b):
            trainer.optimizer.zero_grad()
            loss = trainer.update(b)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(trainer.model.parameters(), opt['max_grad_norm'])
            trainer.optimizer.step()

        def adjacency(b):                           # model/gcn.py:96-110, the host section of the reference's forward
            import numpy as np
            words, masks, deprel, head, subj_pos, obj_pos = (b[i].numpy() for i in (0, 1, 4, 5, 6, 7))
            lens = (masks == 0).astype(np.int64).sum(1)
            maxlen = int(max(lens))
            trees = [ref_tree.head_to_tree(head[i], words[i], lens[i], opt['prune_k'], subj_pos[i], obj_pos[i], deprel[i])
                     for i in range(len(lens))]
            return np.concatenate([ref_tree.tree_to_adj(maxlen, t, directed=False, self_loop=True).reshape(1, maxlen, maxlen)
                                   for t in trees], axis=0)
    else:
        from oracle import gcn_oracle
        model = gcn_oracle.DenseClassifier(opt)
        model.train()
        optim = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=opt['lr'])
        kind = 'port'

        def step(b):
            gcn_oracle.train_step(model, optim, b, opt['max_grad_norm'])

        def adjacency(b):
            return model.gcn_model.adjacency(list(b[:-2]))
    for i in range(warmup):
        step(batches[i % len(batches)])
    t0 = time.perf_counter()
    for i in range(steps):
        step(batches[i % len(batches)])
    dt = time.perf_counter() - t0
    # the host-side tree + dense adjacency section alone (gcn.py:105-107), single Python thread
    t1 = time.perf_counter()
    n_adj = max(1, min(steps, 10))
    for i in range(n_adj):
        adjacency(batches[i % len(batches)])
    adj_ms = (time.perf_counter() - t1) / n_adj * 1e3
    return {'value': BATCH * steps / dt, 'ms_per_step': dt / steps * 1e3, 'cores': cores, 'tree_adj_ms': adj_ms,
            'kind': kind,
            'sample': '%d steps x %d sentences after %d warm-up, V=%d, train mode (dropout on), %s' % (
                steps, BATCH, warmup, VOCAB,
                'unmodified reference GCNTrainer from baseline/_ref, torch CPU' if kind == 'reference' else
                'oracle port (baseline/_ref not installed)')}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = cpu_reference_run(args, args.steps, max(args.warmup, 1))
    line = {'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, max(args.gpus, 1)),
            'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': r['kind'],
                             'sample': r['sample'] + '; one host process regardless of --gpus',
                             'tree_adj_ms_per_batch': r['tree_adj_ms']},
            'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ b200 ---------

def aggregation_roofline(args, peaks):
    """K2 forward at the large synthetic shape, timed alone with CUDA events on the launching stream."""
    import torch
    from gcn_over_pruned_trees_b200 import ops, synth
    B, T, H = args.roofline_batch, 512, 512
    batch = synth.make_batch_torch(7, B, T, device='cuda')
    csr = ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], -1)
    y = torch.randn(B * T, H, device='cuda')
    bias = torch.zeros(H, device='cuda')
    rng = torch.tensor([1, 1], dtype=torch.int64, device='cuda')
    n_rows = int((csr.flags != 0).sum())
    nnz = int(csr.rowptr[:, T].sum())
    results = {}
    # 'fwd_train' is the variant the training step launches for every layer but the last: dropout + activation bit mask
    for name, drop_p, want_act in (('fwd', 0.0, False), ('fwd_dropout', 0.5, False), ('fwd_train', 0.5, True)):
        for _ in range(3):
            out = ops.aggregate_fwd(y, csr, bias, drop_p=drop_p, rng_state=rng, want_act=want_act)
        reps = 10
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        torch.cuda.synchronize()
        for a, b in ev:                      # y + out = 8.6 GB >> 126 MB of L2: every launch streams from HBM
            a.record()
            out = ops.aggregate_fwd(y, csr, bias, drop_p=drop_p, rng_state=rng, want_act=want_act)
            b.record()
        torch.cuda.synchronize()
        results[name] = sum(a.elapsed_time(b) for a, b in ev) / reps
        del out
    gout = torch.randn(B, T, H, device='cuda')
    out, act = ops.aggregate_fwd(y, csr, bias, drop_p=0.5, rng_state=rng, want_act=True)
    for _ in range(2):                       # as the model runs it: [out > 0] from the 1-bit activation mask
        ops.aggregate_bwd(gout, None, csr, drop_p=0.5, act=act)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record()
        ops.aggregate_bwd(gout, None, csr, drop_p=0.5, act=act)
        b.record()
    torch.cuda.synchronize()
    results['bwd'] = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    db = torch.zeros(H, device='cuda')      # as FusedTrainStep runs it: prologue fused into the producer of gout
    for _ in range(2):
        ops.aggregate_bwd_pre(gout, csr, dbias_out=db)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record()
        ops.aggregate_bwd_pre(gout, csr, dbias_out=db)
        b.record()
    torch.cuda.synchronize()
    results['bwd_pre'] = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    bytes_bwd_pre = 2 * B * T * H * 4 + 4 * (B * (T + 1)) + 4 * nnz + 4 * B * T
    # algorithmic bytes: read each projected row once + write each output row once + CSR + denom
    bytes_fwd = 2 * B * T * H * 4 + 4 * (B * (T + 1)) + 4 * nnz + 4 * B * T + B * T
    bytes_bwd = 2 * B * T * H * 4 + B * T * H // 8 + 4 * (B * (T + 1)) + 4 * nnz + 4 * B * T
    peak = peaks['hbm_gbs']
    traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture
    tpath = os.path.join(REPO, 'profiles', 'r02_k2_fwd_traffic.json')
    if os.path.exists(tpath) and B == 4096:
        t = json.load(open(tpath))
        traffic = t['dram_bytes_read'] + t['dram_bytes_write']
    # K1 and K4 forward at the same shape (SURVEY.md 8d names them as HBM-bound kernels too); never lets the K2 record down
    k1k4 = {}
    try:
        def _avg_ms(fn, reps=10, warm=3):
            for _ in range(warm):
                fn()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            torch.cuda.synchronize()
            for a, b in evs:
                a.record()
                fn()
                b.record()
            torch.cuda.synchronize()
            return sum(a.elapsed_time(b) for a, b in evs) / reps
        t_k1 = _avg_ms(lambda: ops.prune_csr(batch[5], batch[6], batch[7], batch[4], batch[1], -1))
        k1_bytes = B * T * 33 + 4 * B * (T + 1) + 5 * nnz + 5 * B * T + 8 * B
        t_k4 = _avg_ms(lambda: ops.pool3_fwd(out, csr, ops.POOL_TYPES['max']))
        k4_bytes = B * T * H * 4 + B * 3 * H * 8 + B * T
        k1k4 = {'k1_prune_csr_ms': t_k1, 'k1_bytes_per_launch': k1_bytes, 'k1_gbs': k1_bytes / (t_k1 * 1e-3) / 1e9,
                'k1_frac': k1_bytes / (t_k1 * 1e-3) / 1e9 / peak,
                'k4_pool3_fwd_ms': t_k4, 'k4_bytes_per_launch': k4_bytes, 'k4_gbs': k4_bytes / (t_k4 * 1e-3) / 1e9,
                'k4_frac': k4_bytes / (t_k4 * 1e-3) / 1e9 / peak}
    except Exception as exc:
        k1k4 = {'k1_k4_error': repr(exc)}
    bytes_train = bytes_fwd + B * T * H // 8             # + the 1-bit activation mask the backward reads
    ach = bytes_train / (results['fwd_train'] * 1e-3) / 1e9
    del y, out, gout, act
    torch.cuda.empty_cache()
    gbs = lambda nbytes, ms: nbytes / (ms * 1e-3) / 1e9  # noqa: E731
    return {'bound': 'hbm', 'kernel': 'aggregate_fwd_kernel<8,512,1,1,0,1> (K2 forward as the training step runs it: '
                                      'dropout 0.5 + activation bit mask)',
            'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'peak_source': peaks['source'],
            'traffic': traffic,
            'workload': 'large512: B=%d x T=512 trees, H=512, prune_k=-1, rows=%d, nnz=%d' % (B, n_rows, nnz),
            'bytes_per_launch': bytes_train, 'ms_per_launch': results['fwd_train'],
            'frac_of_nominal_8000': ach / 8000.0,
            'other': {'fwd_no_dropout_ms': results['fwd'], 'fwd_no_dropout_gbs': gbs(bytes_fwd, results['fwd']),
                      'fwd_dropout_no_mask_ms': results['fwd_dropout'],
                      'fwd_dropout_no_mask_gbs': gbs(bytes_fwd, results['fwd_dropout']),
                      'fwd_dropout_no_mask_frac': gbs(bytes_fwd, results['fwd_dropout']) / peak,
                      'bwd_ms': results['bwd'], 'bwd_gbs': gbs(bytes_bwd, results['bwd']),
                      'bwd_bytes_per_launch': bytes_bwd,
                      'bwd_pre_scaled_ms': results['bwd_pre'],
                      'bwd_pre_scaled_gbs': gbs(bytes_bwd_pre, results['bwd_pre']),
                      'bwd_pre_scaled_frac': gbs(bytes_bwd_pre, results['bwd_pre']) / peak, **k1k4}}


def load_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': float(p['hbm_gbs']), 'bf16_tflops': float(p['bf16_tflops']),
                'bf16_tflops_sustained': float(p.get('bf16_tflops_sustained', p['bf16_tflops'])),
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0,
            'source': 'fallback (B200_PROFILING.md)'}


def dp_lockstep_check(args, rank, world, dev, steps=5):
    """Correctness of the data-parallel step at this world size, before anything is timed (reduced
    gradients == single-process gradients on the concatenated batch).  Every rank steps on rows rank::world of the same
    global batches through FusedTrainStep(data_parallel=True) -- the K8 exchange the timed region uses, same GEMM mode,
    same vocabulary -- and rank 0 also steps a single-process replica on the full batches.  LOCKSTEP: before every step
    all ranks load the single-process replica's parameters (free-running trajectories are not comparable at 1e-5: a
    pre-activation that is zero to within rounding flips a ReLU).  Dropout is off here: the Philox
    streams are keyed by the sentence's row in the batch, which sharding changes.  Steps 1-2 run eagerly, step 3 is
    captured and replayed, steps 4-5 are replays: every launch mode of the timed engine is covered."""
    import torch
    import torch.distributed as dist
    from gcn_over_pruned_trees_b200 import synth
    from gcn_over_pruned_trees_b200.engine import FusedTrainStep
    from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
    over = dict(vocab_size=VOCAB, prune_k=args.prune_k, cuda=True, input_dropout=0.0, gcn_dropout=0.0,
                gemm_mode=args.gemm)
    width = 96
    batches = [synth.make_batch(300 + i, batch_size=BATCH * world, vocab_size=VOCAB, pad_to=width) for i in range(3)]
    _stdout = sys.stdout
    sys.stdout = open(os.devnull, 'w')
    torch.manual_seed(11)
    tr = GCNTrainer(synth.tacred_opt(**over))
    tr.model.train()
    eng = FusedTrainStep(tr, data_parallel=True, max_rows=BATCH * 128)
    ref = ref_eng = None
    if rank == 0:
        torch.manual_seed(11)
        ref = GCNTrainer(synth.tacred_opt(**over))
        ref.model.train()
        ref_eng = FusedTrainStep(ref)
    sys.stdout = _stdout
    params = list(tr.model.parameters())
    sizes = [p.numel() for p in params]
    worst, loss_err, identical = 0.0, 0.0, True
    for s in range(steps):
        vec = torch.empty(sum(sizes), device=dev)
        if rank == 0:
            vec.copy_(torch.cat([p.detach().reshape(-1) for p in ref.model.parameters()]))
        dist.broadcast(vec, 0)
        for p, chunk in zip(params, vec.split(sizes)):
            p.data.copy_(chunk.view_as(p))                    # in place: the step graphs keep their addresses
        full = batches[s % 3]
        shard = tuple(t[rank::world].contiguous() if torch.is_tensor(t) else t[rank::world] for t in full)
        loss = eng(shard).clone()
        torch.cuda.synchronize()
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        identical = identical and bool(torch.equal(lo, hi))   # bitwise: min == max over ranks for every element
        dist.all_reduce(loss)
        loss /= world
        if rank == 0:
            ref_loss = float(ref_eng(full))
            for a, b in zip(params, ref.model.parameters()):
                worst = max(worst, float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)))
            loss_err = max(loss_err, abs(float(loss) - ref_loss) / abs(ref_loss))
    out = torch.tensor([worst, loss_err, 1.0 if identical else 0.0], dtype=torch.float64, device=dev)
    dist.broadcast(out, 0)
    worst, loss_err = float(out[0]), float(out[1])
    ok = identical and worst < 2e-5 and loss_err < 2e-5
    del eng, ref_eng, tr, ref
    torch.cuda.empty_cache()
    return {'world': world, 'steps': steps, 'mode': 'lockstep, dropout off, gemm %s, K8 peer-memory exchange' % args.gemm,
            'replicas_bit_identical': identical, 'params_rel': worst, 'loss_rel': loss_err,
            'tolerance': 2e-5, 'ok': ok}


def semeval_leg(args, rank, world, dev, flush):
    """BASELINE.json configs[3]: SemEval-shaped batches (9-tuples without NER, 19 classes, in_dim 330, lengths
    clip(Poisson(19),5,97); train_semeval.py:195-222's step), 50 sentences per GPU, same engine as the headline."""
    import torch
    from gcn_over_pruned_trees_b200 import parallel, synth
    from gcn_over_pruned_trees_b200.engine import FusedTrainStep, PackedBatch
    from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
    torch.manual_seed(1234)
    opt = synth.tacred_opt(vocab_size=VOCAB, prune_k=args.prune_k, cuda=True, gemm_mode=args.gemm, dataset='semeval',
                           num_class=19)
    _stdout = sys.stdout
    sys.stdout = open(os.devnull, 'w')
    tr = GCNTrainer(opt)
    sys.stdout = _stdout
    tr.model.train()
    eng = FusedTrainStep(tr, data_parallel=world > 1, max_rows=BATCH * 128)
    nb = 8
    host = [parallel.shard_batch(synth.make_batch(2000 + i, batch_size=BATCH * world, vocab_size=VOCAB, dataset='semeval',
                                                  num_class=19, mean_len=19, min_len=5, max_len=97), rank, world)
            for i in range(nb)]
    res = [PackedBatch(b, device='cpu').to(dev) for b in host]
    for _ in range(4):
        for b in res:
            eng(b)
    torch.cuda.synchronize()
    steps = min(args.steps, 50)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    parallel.barrier()
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(ev):
        flush.fill_(0.0)
        a.record()
        loss = eng(res[i % nb])
        b.record()
    torch.cuda.synchronize()
    parallel.barrier()
    ms = parallel.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), dev) / steps
    out = {'workload': 'semeval_b50_k%d_19cls' % args.prune_k, 'value': BATCH * world / ms * 1e3, 'unit': UNIT,
           'n_gpus': world, 'ms_per_step': ms, 'steps': steps, 'batch_per_gpu': BATCH, 'in_dim': 330, 'num_class': 19,
           'len': 'clip(Poisson(19),5,97)', 'loss': float(loss), 'launches_per_step': eng.launches_per_replay(res[0]),
           'exchange': 'none' if world == 1 else 'nvlink_peer_memory (K8)', 'scaling': 'weak',
           'projection': 'fp32 FFMA (K=330: row pitch 1320 B is not a multiple of 16 B, no TMA descriptor)'}
    del eng, tr, res
    torch.cuda.empty_cache()
    return out


def large512_leg(args, rank, world, dev, peaks):
    """BASELINE.json configs[4]: every sentence 512 tokens, B=4096 per GPU, 2-layer GCN, 360 -> H=512, k=-1, the whole
    training step.  N > 1: one NCCL all-reduce of the flat gradient buffer (dense word-embedding gradient included:
    2M tokens touch the whole vocabulary) -- the exchange north_star names; the step is tens of milliseconds, so it is
    launched eagerly (~20 launches)."""
    import torch
    from gcn_over_pruned_trees_b200 import ops, parallel, synth
    from gcn_over_pruned_trees_b200.engine import FusedTrainStep
    from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer
    B, T, H = args.large_batch, 512, 512
    torch.manual_seed(1234)
    opt = synth.tacred_opt(vocab_size=VOCAB, cuda=True, hidden_dim=H, prune_k=-1, gemm_mode=args.gemm)
    _stdout = sys.stdout
    sys.stdout = open(os.devnull, 'w')
    tr = GCNTrainer(opt)
    sys.stdout = _stdout
    tr.model.train()
    eng = FusedTrainStep(tr, data_parallel='nccl' if world > 1 else False, capture=False)
    batch = synth.make_batch_torch(5 + rank, B, T, device=dev)
    inputs, labels = list(batch[:-2]), batch[-2]
    torch.cuda.reset_peak_memory_stats()
    with torch.no_grad():
        for _ in range(3):
            eng._run(inputs, labels)
        torch.cuda.synchronize()
        steps = args.large_steps
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        parallel.barrier()
        torch.cuda.synchronize()
        for a, b in ev:                         # one step streams ~35 GB: nothing of it survives in the 126 MB L2
            a.record()
            loss, _ = eng._run(inputs, labels)
            b.record()
        torch.cuda.synchronize()
        parallel.barrier()
        times = [a.elapsed_time(b) for a, b in ev]
        ms = parallel.max_over_ranks(sum(times), dev) / steps
        kernels = None
        if world == 1:                          # per-entry-point device time (separate instrumented pass)
            ops.TIMER = ops.KernelTimer()
            for _ in range(2):
                eng._run(inputs, labels)
            summary = ops.TIMER.summary()
            ops.TIMER = None
            kernels = {k: {'calls_per_step': c / 2, 'ms_per_call': t / c}
                       for k, (c, t) in sorted(summary.items(), key=lambda kv: -kv[1][1])}
    n_rows = B * T
    flops = 3 * 2 * n_rows * (360 + H) * H           # fwd + dgrad + wgrad of both projections
    out = {'workload': 'large512: B=%d per GPU x T=512 trees, 360 -> H=512, 2 layers, prune_k=-1, V=%d' % (B, VOCAB),
           'value': B * world / ms * 1e3, 'unit': UNIT, 'n_gpus': world, 'ms_per_step': ms, 'steps': steps,
           'step_ms_min': min(times), 'step_ms_max': max(times), 'loss': float(loss), 'scaling': 'weak',
           'exchange': 'none' if world == 1 else 'nccl_allreduce (one flat fp32 buffer, dense embedding gradient included)',
           'engine': 'fused step, eager launches', 'gemm': args.gemm,
           'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9,
           'projection_tflops_algorithmic': flops / (ms * 1e-3) / 1e12, 'kernels': kernels}
    del eng, tr, batch, inputs, labels
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch
    from gcn_over_pruned_trees_b200 import _lib, ops, parallel, synth
    from gcn_over_pruned_trees_b200.model.trainer import GCNTrainer

    rank, local_rank, world = parallel.init_from_env('nccl')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    torch.backends.cuda.matmul.allow_tf32 = False      # fp32 parity mode for the cuBLAS tail (MLP / classifier)
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(1234)                            # identical replicas on every rank
    opt = synth.tacred_opt(vocab_size=VOCAB, prune_k=args.prune_k, cuda=True, gemm_mode=args.gemm)
    _stdout = sys.stdout
    sys.stdout = open(os.devnull, 'w')
    trainer = GCNTrainer(opt)
    sys.stdout = _stdout
    model = trainer.model
    model.train()
    reducer = parallel.GradAllReducer(model.parameters())
    dp_check = None
    if world > 1 and not args.no_dp_check and not args.autograd_engine and not args.eager:
        dp_check = dp_lockstep_check(args, rank, world, dev)
    # N ranks: every step is ONE global batch of 50 x N sentences, length-sorted as the loader sorts it, and rank r takes rows
    # r::N (parallel.shard_batch) -- the ranks' shards then have near-equal widths and no rank waits for a
    # straggler with a longer batch.  N = 1: the same 50-sentence batches as before.
    host = [parallel.shard_batch(synth.make_batch(1000 + i, batch_size=BATCH * world, vocab_size=VOCAB), rank, world)
            for i in range(N_BATCHES)]
    host = [tuple(t.pin_memory() if torch.is_tensor(t) else t for t in b) for b in host]
    resident = [tuple(t.to(dev) if torch.is_tensor(t) else t for t in b) for b in host]
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    params = list(model.parameters())

    def eager_step(batch):       # the reference's five calls, train.py:213-227
        trainer.optimizer.zero_grad(set_to_none=False)
        loss = trainer.update(batch)
        loss.backward()
        reducer.reduce()
        torch.nn.utils.clip_grad_norm_(params, opt['max_grad_norm'])
        trainer.optimizer.step()
        return loss

    from gcn_over_pruned_trees_b200.engine import FusedTrainStep, GraphedTrainStep
    fused = not args.autograd_engine and FusedTrainStep.unsupported_reason(trainer) is None
    if fused:       # N > 1: gradients meet through NVLink peer memory inside the step's own kernels (K8), no NCCL
        graphed = FusedTrainStep(trainer, data_parallel=world > 1, max_rows=BATCH * 128)
    else:
        graphed = GraphedTrainStep(trainer, reducer=reducer)
    step = eager_step if args.eager else graphed
    engine_info = engine_config(args, world)
    ex = getattr(graphed, 'exchange', None)
    if ex is not None:      # how the K8 regions are mapped on this box (parallel.PeerExchange)
        engine_info['exchange'] += ': regions in %s, push through %s' % (
            'torch symmetric memory' if ex.symm is not None else 'cudaMalloc + cudaIpc',
            'an NVSwitch multicast mapping (multimem.st)' if ex.multicast else 'one store per peer')
    host_tuples, resident_tuples = host, resident
    if fused and not args.eager:    # loader batches packed into one contiguous buffer each: one copy per step
        from gcn_over_pruned_trees_b200.engine import PackedBatch
        host = [PackedBatch(b, pin=True) for b in host_tuples]
        resident = [b.to(dev) for b in host]

    # warm-up: every batch shape runs eagerly 3x, is captured, and is replayed at least once
    for _ in range(5):
        for bt in resident:
            step(bt)
    for i in range(max(args.warmup, 3)):
        step(resident[i % N_BATCHES])
    torch.cuda.synchronize()

    # ---- value: device-resident inputs, per-step CUDA events, L2 flushed between steps -------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = _lib.lib().gpt_launch_count()
    launches = 0
    parallel.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for i, (a, b) in enumerate(events):
        flush.fill_(0.0)
        a.record()
        step(resident[i % N_BATCHES])
        b.record()
        if not args.eager:
            launches += graphed.launches_per_replay(resident[i % N_BATCHES])
    torch.cuda.synchronize()
    parallel.barrier()
    wall = time.perf_counter() - wall0
    launches += _lib.lib().gpt_launch_count() - launches0
    step_times = sorted(a.elapsed_time(b) for a, b in events)
    dev_ms = sum(step_times)
    dev_ms = parallel.max_over_ranks(dev_ms, dev)
    step_ms = {'min': step_times[0], 'median': statistics.median(step_times), 'max': step_times[-1],
               'of': 'rank 0, CUDA events around each timed step'}
    clocks = sampler.stop() if sampler else None
    ms_per_step = dev_ms / args.steps
    value = BATCH * world * args.steps / (dev_ms * 1e-3)

    # ---- e2e: pinned host batches through the public API, loss read back every step -----------------------------
    h2d = sum(t.numel() * t.element_size() for t in host_tuples[0] if torch.is_tensor(t))

    def timed_host_loop(fn, n):
        for i in range(3):
            fn(host[i % N_BATCHES]).item()
        parallel.barrier()
        torch.cuda.synchronize()
        total = 0.0
        for i in range(n):
            flush.fill_(0.0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(host[i % N_BATCHES]).item()
            total += time.perf_counter() - t0
        parallel.barrier()
        return parallel.max_over_ranks(total, dev)

    e2e_s = timed_host_loop(step, args.steps)
    e2e = {'value': BATCH * world * args.steps / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
           'd2h_bytes_per_step': 4, 'ms_per_step': e2e_s / args.steps * 1e3,
           'api': ('GCNTrainer.update(host batch) + backward + clip + SGD + loss.item()' if args.eager else
                   'train_step(pinned host batch, packed: 1 H2D copy) [one CUDA-graph replay] + loss.item()')}
    if fused and not args.eager:
        host = host_tuples
        t_s = timed_host_loop(step, min(args.steps, 50))
        e2e['loader_tuple'] = {'value': BATCH * world * min(args.steps, 50) / t_s,
                               'ms_per_step': t_s / min(args.steps, 50) * 1e3,
                               'api': 'train_step(pinned 10-tuple as the reference loader emits it: 9 H2D copies)'}
    if world == 1:
        # the reference's own five-call loop (train.py:213-227) on the new `model` package, nothing else changed:
        # update() / backward() / optimizer.step() replay captured graphs (engine.FastUpdate), clip_grad_norm_ is torch's
        n_eager = min(args.steps, 50)
        host = host_tuples
        trainer.fast_update = True
        for _ in range(4):                          # every batch shape: 2 eager steps, capture, one replay
            for bt in host_tuples:
                eager_step(bt)
        eager_s = timed_host_loop(eager_step, n_eager)
        fast = getattr(trainer, '_fast', None)
        e2e['dropin_eager'] = {'value': BATCH * world * n_eager / eager_s, 'ms_per_step': eager_s / n_eager * 1e3,
                               'api': 'reference call sequence train.py:213-227 (zero_grad, update, backward, '
                                      'clip_grad_norm_, optimizer.step) on the new model package, pinned host 10-tuples, '
                                      'loss.item() every step',
                               'path': 'engine.FastUpdate: 3 CUDA-graph replays per step' if fast is not None
                                       else 'per-op autograd', 'graph_replays': fast.replays if fast else 0}
        trainer.fast_update = False                 # the same five calls on the per-op autograd Functions
        n_auto = min(args.steps, 20)
        auto_s = timed_host_loop(eager_step, n_auto)
        trainer.fast_update = True
        e2e['dropin_autograd'] = {'value': BATCH * world * n_auto / auto_s, 'ms_per_step': auto_s / n_auto * 1e3,
                                  'api': 'same five calls with GPT_FAST_UPDATE=0: ~115 eager launches per step'}
    # ---- K9: batches assembled on the device from a resident token arena ----------------------
    loader_info = None
    if fused and not args.eager and not args.no_loader:
        from gcn_over_pruned_trees_b200.data.loader import DataLoader as DeviceLoader
        examples = []
        for bt in host_tuples:
            lens = (~bt[1]).sum(1).tolist()
            for r, n in enumerate(lens):
                examples.append(tuple(bt[f][r, :n].tolist() for f in (0, 2, 3, 4, 5, 6, 7)) + (int(bt[8][r]),))
        _stdout = sys.stdout
        sys.stdout = open(os.devnull, 'w')
        dl = DeviceLoader.from_processed(examples, BATCH, {'word_dropout': 0.04, 'lower': False, 'dataset': 'tacred'},
                                         evaluation=False, device=dev, seed=1234 + rank)
        sys.stdout = _stdout
        n_dl = len(dl)
        for i in range(4 * n_dl):                           # every batch shape captured before timing
            step.step_from(dl, i % n_dl)
        torch.cuda.synchronize()
        parallel.barrier()
        total = 0.0
        for i in range(args.steps):
            flush.fill_(0.0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step.step_from(dl, i % n_dl).item()
            total += time.perf_counter() - t0
        parallel.barrier()
        total = parallel.max_over_ranks(total, dev)
        e2e['device_loader'] = {'value': BATCH * world * args.steps / total, 'ms_per_step': total / args.steps * 1e3,
                                'h2d_bytes_per_step': 0,
                                'api': 'loss = engine.step_from(loader, i): K9 writes the batch from the resident token arena '
                                       '(word dropout on) into the static input buffer, then one CUDA-graph replay + '
                                       'loss.item()'}
        if rank == 0:
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(dl.RING * n_dl):                 # fill every shape's buffer ring first
                dl.packed(i % n_dl)
            torch.cuda.synchronize()
            ea.record()
            for i in range(200):
                dl.packed(i % n_dl)
            eb.record()
            torch.cuda.synchronize()
            tokens = sum(len(e[0]) for e in examples) / n_dl
            width = sum(b[2] for b in dl.batches) / n_dl
            loader_info = {'kernel': 'build_batch_kernel (K9)', 'us_per_batch_incl_launch_gap': ea.elapsed_time(eb) / 200 * 1e3,
                           'algorithmic_bytes_per_batch': int(tokens * 28 + 57 * BATCH * width + 12 * BATCH),
                           'bound': 'launch latency (0.2 MB per batch)'}

    # ---- BASELINE.json configs[3] and configs[4] at this N (sub-records; the headline workload is unchanged) ---------
    extra = None
    if fused and not args.eager and not args.no_extra:
        extra = {}
        for name, fn in (('semeval_b50', lambda: semeval_leg(args, rank, world, dev, flush)),
                         ('large512', lambda: large512_leg(args, rank, world, dev, load_peaks()))):
            try:
                extra[name] = fn()
            except Exception as exc:            # keep the headline line; the failure is in the record
                extra[name] = {'error': repr(exc)}
                if world > 1:
                    raise
    if world > 1:
        parallel.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return

    # ---- per-entry-point device time inside the step (separate instrumented pass, not the headline) ------------
    kernels = None
    ops.TIMER = ops.KernelTimer() if world == 1 else None
    n_prof = min(args.steps, 20) if world == 1 else 0
    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    pa.record()
    for i in range(n_prof):
        if fused and not args.eager:        # the same call sequence the graph replays, launched eagerly
            with torch.no_grad():
                graphed._run(list(resident_tuples[i % N_BATCHES][:-2]), resident_tuples[i % N_BATCHES][-2])
        else:
            eager_step(resident_tuples[i % N_BATCHES])
    pb.record()
    if ops.TIMER is not None:
        summary = ops.TIMER.summary()
        ops.TIMER = None
        prof_ms = pa.elapsed_time(pb)
        # eager launches: each interval also contains the host-side gap before the launch; the ncu launch list under
        # profiles/ holds the kernel-only durations
        kernels = {k: {'calls_per_step': c / n_prof, 'us_per_call_incl_launch_gap': ms / c * 1e3,
                       'share_of_step': ms / prof_ms}
                   for k, (c, ms) in sorted(summary.items(), key=lambda kv: -kv[1][1])}

    peaks = load_peaks()
    roof = None
    if not args.no_roofline and world == 1:
        try:
            roof = aggregation_roofline(args, peaks)
        except Exception as exc:                # keep the headline line even if the 8.6 GB run cannot be placed
            roof = {'error': repr(exc)}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(args, args.cpu_steps, 3)
        cpu = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample'],
               'ms_per_step': r['ms_per_step'], 'tree_adj_ms_per_batch': r['tree_adj_ms']}

    if loader_info is not None and not args.no_cpu_baseline and world == 1:
        from oracle import loader_oracle            # the reference loader's per-batch Python work, restated
        t0 = time.perf_counter()
        n_cpu = 0
        while time.perf_counter() - t0 < 2.0:
            k = n_cpu % (len(examples) // BATCH)
            loader_oracle.get_batch(examples[k * BATCH:(k + 1) * BATCH], False, 0.04)
            n_cpu += 1
        loader_info['cpu_port_us_per_batch'] = (time.perf_counter() - t0) / n_cpu * 1e6
        loader_info['cpu_port'] = 'oracle/loader_oracle.get_batch (data/loader.py:81-141 restated), 1 thread, host tensors only'
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args, world),
            'engine': engine_info, 'step_ms': step_ms, 'dp_check': dp_check,
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
            'gpu_launches_per_step': launches / args.steps, 'wall_ms_per_step_incl_flush': wall / args.steps * 1e3,
            'roofline': roof, 'cpu_baseline': cpu, 'extra': extra, 'loader': loader_info, 'kernels': kernels}
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
