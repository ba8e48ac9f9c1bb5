"""Multi-process sharding logic on CPU: world_size 2, gloo backend."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from salsa_b200 import sharding


def test_clip_range_partitions():
    for n in (0, 1, 7, 600, 4800, 4801):
        for world in (1, 2, 3, 8):
            ranges = [sharding.clip_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.clip_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_clips, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = sharding.clip_range(n_clips, rank, world)
        # per-clip "outputs": row i holds clip index i (what a rank would produce for its own clips)
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None, None].expand(hi - lo, 3, 4).contiguous()
        full = sharding.gather_clip_outputs(local, n_clips)
        ok = full.shape == (n_clips, 3, 4) and torch.equal(full[:, 0, 0], torch.arange(n_clips, dtype=torch.float32))
        # scaler statistics: every rank holds the sums of its own clips; the reduction must equal the global result
        from salsa_b200.features import FeatureScaler
        g = torch.Generator().manual_seed(0)
        data = torch.randn(n_clips, 4, 7, 5, generator=g, dtype=torch.float64) * 10 - 50        # (clips, ch, frames, F)
        mine = data[lo:hi]
        sums = torch.stack([mine.sum(dim=(0, 2)), (mine ** 2).sum(dim=(0, 2))], dim=-1)             # (4, F, 2)
        mean, std = FeatureScaler.reduce_statistics(sums, (hi - lo) * 7)
        ref_mean = data.permute(1, 0, 2, 3).reshape(4, -1, 5).mean(dim=1)
        ref_std = data.permute(1, 0, 2, 3).reshape(4, -1, 5).std(dim=1, unbiased=False)
        ok = ok and mean.shape == (4, 1, 5) and bool(torch.allclose(torch.from_numpy(mean[:, 0]).double(), ref_mean, atol=1e-5))
        ok = ok and bool(torch.allclose(torch.from_numpy(std[:, 0]).double(), ref_std, atol=1e-5))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and t.item() == world
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_clips', [5, 8])
def test_gather_world_size_2(n_clips):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: True, 1: True}


def test_gather_single_process():
    x = torch.zeros(4, 2)
    assert sharding.gather_clip_outputs(x, 4) is x
    with pytest.raises(ValueError):
        sharding.gather_clip_outputs(x, 5)


class _FakeExtractor:
    """The SalsaExtractor surface on CPU tensors: 'features' that identify (rank, clip) so that a gather can be checked."""
    freq_dim = 5

    def __init__(self, rank):
        self.rank = rank

    def n_frames(self, n_samples):
        return 1 + n_samples // 300

    def extract(self, audio, out=None):
        out[:] = audio[:, :1, :1, None].expand_as(out) + 1000.0 * self.rank
        return out


def _chunk_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n, chunk, n_samples = 7, 3, 900                         # 3 chunks, the last one short
        audio = torch.arange(n, dtype=torch.float32)[:, None, None].expand(n, 4, n_samples).contiguous()
        cg = sharding.ChunkedFeatureGather(_FakeExtractor(rank), chunk, n_samples, torch.device('cpu'))
        ok = cg.verify(audio[:chunk])
        seen = []
        pending = []
        for c in range(0, n, chunk):
            a = audio[c:c + chunk]
            pending.append((a.shape[0], c, cg.step(a)[1]))
            if len(pending) == 2:                                # chunk i is complete once step i + 1 returned and was waited:
                cg.work[(cg.i - 2) & 1].wait()                   # the consumer waits for the older gather, then reads it
                m, c0, buf = pending.pop(0)
                seen.append((c0, buf[:, :m, 0, 0, 0].clone()))
        cg.finish()
        for m, c0, buf in pending:
            seen.append((c0, buf[:, :m, 0, 0, 0].clone()))
        for c0, vals in seen:
            want = torch.stack([torch.arange(c0, c0 + vals.shape[1], dtype=torch.float32) + 1000.0 * r for r in range(world)])
            ok = ok and torch.equal(vals, want)
        q.put((rank, bool(ok) and len(seen) == 3))
    finally:
        dist.destroy_process_group()


def test_chunked_feature_gather_world_size_2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_chunk_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: True, 1: True}


def _train_worker(rank, world, port, q):
    """SeldTrainer's gradient plane on gloo: after `step` every rank holds the AVERAGE of the per-rank gradients (buckets
    all-reduced while the backward pass runs), hence identical parameters."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from salsa_b200 import train
        from salsa_b200.crnn import random_state_dict
        torch.manual_seed(0)
        torch.set_num_threads(2)
        sd = random_state_dict(0)
        tr = train.SeldTrainer(sd, device='cpu', wire_dtype=None, dropout=False, bucket_bytes=4 << 20)
        assert len(tr.reducer.buckets) > 3
        g = torch.Generator().manual_seed(100 + rank)                 # every rank its own batch
        x = torch.randn(1, 7, 32, 32, generator=g)
        tgt = {'event_frame_gt': (torch.rand(1, 4, 12, generator=g) > 0.5).float(), 'doa_frame_gt': torch.randn(1, 4, 36, generator=g)}
        loss = tr.step(x, tgt)
        mine = tr.flat_grad.clone()
        # the same gradients without the reducer, averaged by hand
        solo = train.SeldTrainer(sd, device='cpu', wire_dtype=None, dropout=False)
        solo.reducer.world = 1
        solo.step(x, tgt)
        local = solo.flat_grad.clone()
        both = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(both, local)
        want = sum(both) / world
        ok = bool(torch.allclose(mine, want, rtol=1e-5, atol=1e-7)) and bool(torch.isfinite(loss).all()) and float(mine.abs().max()) > 0
        ok = ok and not torch.allclose(local, want, rtol=1e-3, atol=1e-6)          # the ranks really saw different data
        # the single-call variant used after a CUDA-graph replay of the backward pass: hooks off, one all-reduce of the flat buffer
        solo.reducer.world = world
        solo.reducer.enabled = False
        solo._step_eager(x, tgt, gradients_only=True)
        ok = ok and bool(torch.allclose(solo.flat_grad, local, rtol=1e-5, atol=1e-7))      # nothing reduced yet
        solo.reducer.reduce_all()
        ok = ok and bool(torch.allclose(solo.flat_grad, want, rtol=1e-5, atol=1e-7))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_trainer_gradient_allreduce_world_size_2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: True, 1: True}
