"""The multi-GPU data plane on real NCCL (needs >= 2 visible GPUs; skipped on a one-GPU box): the final gather of per-clip
outputs (`gather_clip_outputs`) and the chunked, overlapped all-gather of features of BASELINE.json
configs[3] (`ChunkedFeatureGather`), both checked against what one process computes for all clips."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import salsa_b200
    from salsa_b200 import sharding
    from oracle import synth
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        n_clips, n_samples = 7, 24000
        clips = np.stack([synth.make_clip(80 + i, 'mic', seconds=1.0) for i in range(n_clips)])
        ex = salsa_b200.SalsaExtractor('mic', fmax_doa=4000)
        everything = ex.extract(torch.from_numpy(clips).to(dev))                  # what one GPU computes for all clips
        lo, hi = sharding.clip_range(n_clips, rank, world)
        mine = ex.extract(torch.from_numpy(clips[lo:hi]).to(dev))
        full = sharding.gather_clip_outputs(mine, n_clips)
        ok = tuple(full.shape) == tuple(everything.shape) and torch.equal(full, everything)
        # chunked gather: every rank walks the same number of chunks (here: its first 3 clips in chunks of 2)
        local = torch.from_numpy(clips[lo:lo + 3]).to(dev)
        used = []
        for transport in ('auto', 'nccl'):
            cg = sharding.ChunkedFeatureGather(ex, 2, n_samples, dev, transport=transport)
            used.append(cg.transport)
            ok = ok and cg.verify(local[:2])
            got = []
            for c in (0, 2):
                a = local[c:c + 2]
                _, buf = cg.step(a)
                cg.finish()
                got.append(buf[:, :a.shape[0]].clone())
            got = torch.cat(got, dim=1)                                               # (world, 3, 7, T, F)
            for r in range(world):
                rlo, _ = sharding.clip_range(n_clips, r, world)
                ok = ok and torch.equal(got[r], everything[rlo:rlo + 3])
            # back-to-back steps without a finish in between: buffer reuse ordered by the barriers / work handles
            for rep in range(4):
                _, buf = cg.step(local[:2])
            cg.finish()
            for r in range(world):
                rlo, _ = sharding.clip_range(n_clips, r, world)
                ok = ok and torch.equal(buf[r, :2], everything[rlo:rlo + 2])
            del cg
        print('rank {} transports: {}'.format(rank, used), flush=True)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (NCCL over NVLink)')
def test_nccl_gather_of_clip_outputs_and_chunked_feature_gather():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert results == {0: True, 1: True}
