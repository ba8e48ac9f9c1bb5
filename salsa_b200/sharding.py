"""Clip sharding across the GPUs of one box.

Clips are independent in feature extraction and in CRNN inference (the only dependencies -- frame
wrap padding, the noise-floor tracker, the BiGRU -- live inside a clip), so the path shards by
contiguous clip ranges, one process per GPU, with no data-path collective.  The single exchange step
is the final gather of per-clip outputs: `gather_clip_outputs` is an
`all_gather` over the process group (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def clip_range(n_clips: int, rank: int, world_size: int):
    """Contiguous range [lo, hi) of rank `rank`: [r*C/W, (r+1)*C/W)."""
    if not (0 <= rank < world_size):
        raise ValueError('rank {} outside world of {}'.format(rank, world_size))
    return rank * n_clips // world_size, (rank + 1) * n_clips // world_size


def gather_clip_outputs(local: torch.Tensor, n_clips: int, group=None) -> torch.Tensor:
    """local: this rank's per-clip outputs (n_local, ...) for its clip_range -> (n_clips, ...) on every
    rank, in clip order.  Shards may differ by one clip; they are padded to the largest for the
    collective and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized():
        if local.shape[0] != n_clips:
            raise ValueError('single process: expected {} clips, got {}'.format(n_clips, local.shape[0]))
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = clip_range(n_clips, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError('rank {} owns clips [{}, {}) but passed {} rows'.format(rank, lo, hi, local.shape[0]))
    n_max = max(clip_range(n_clips, r, world)[1] - clip_range(n_clips, r, world)[0] for r in range(world))
    padded = local.new_zeros((n_max,) + tuple(local.shape[1:]))
    padded[:hi - lo] = local
    out = local.new_empty((world * n_max,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    parts = []
    for r in range(world):
        rlo, rhi = clip_range(n_clips, r, world)
        parts.append(out[r * n_max:r * n_max + (rhi - rlo)])
    return torch.cat(parts, dim=0)


class ChunkedFeatureGather:
    """BASELINE.json configs[3]: clips sharded across the GPUs, features of every rank gathered on every rank over
    NVLink ("chunk and overlap with compute").

    Every rank walks its own clips in chunks.  `step(audio_chunk)` extracts chunk i and starts its exchange; the
    extraction of chunk i + 1 runs while chunk i crosses NVLink.  The gathered chunk `(world, chunk, 7, T, F)` of step i
    is valid once step i + 2 (or `finish()`) has been called, and is overwritten by step i + 2 -- the consumer (a trainer,
    a writer) takes it in between, so memory stays at two chunks whatever the number of clips.  `extractor` is anything
    with the `SalsaExtractor` surface (`extract(audio, out=)`, `n_frames(n_samples)`, `freq_dim`).

    transport:
      'p2p'   the gathered buffers are symmetric memory (every rank maps every peer's buffer): the extraction kernels
              write the chunk straight into this rank's slot of its own buffer, and the copy engines push that slot into
              the same slot of every peer's buffer over NVLink (no SM is spent on the exchange; two device-side barriers
              per chunk order the pushes against the peers' reuse of the buffer).  CUDA only.
      'nccl'  `all_gather_into_tensor` on the process group (NCCL on the GPU box, gloo in the CPU tests), asynchronous.
      'auto'  'p2p' when symmetric memory can be set up, else 'nccl'.
    """

    def __init__(self, extractor, chunk: int, n_samples: int, device, group=None, transport: str = 'auto'):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError('ChunkedFeatureGather needs an initialised process group')
        if transport not in ('auto', 'p2p', 'nccl'):
            raise ValueError('transport must be auto, p2p or nccl')
        self.ex, self.chunk, self.group = extractor, int(chunk), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device)
        self.shape = (self.chunk, 7, extractor.n_frames(n_samples), extractor.freq_dim)
        self.work = [None, None]
        self.i = 0
        self.transport = 'nccl'
        self.fallback_reason = None
        if transport != 'nccl' and self.device.type == 'cuda':
            try:
                self._setup_p2p()
                self.transport = 'p2p'
            except Exception as exc:                            # noqa: BLE001 -- no symmetric memory on this box / build
                if transport == 'p2p':
                    raise
                self.fallback_reason = repr(exc)[:200]
        elif transport == 'p2p':
            raise ValueError("transport='p2p' needs a CUDA device")
        if self.transport == 'nccl':
            self.local = [torch.empty(self.shape, dtype=torch.float32, device=self.device) for _ in range(2)]
            self.gathered = [torch.empty((self.world,) + self.shape, dtype=torch.float32, device=self.device) for _ in range(2)]

    # ---- peer-to-peer transport --------------------------------------------------------------------------------
    def _setup_p2p(self):
        import torch.distributed._symmetric_memory as symm
        group = self.group if self.group is not None else dist.group.WORLD
        full = (self.world,) + self.shape
        self.gathered = [symm.empty(full, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.handles = [symm.rendezvous(g, group) for g in self.gathered]
        # peer[b][r] = rank r's gathered[b], mapped into this process
        self.peer = [[h.get_buffer(r, full, torch.float32) for r in range(self.world)] for h in self.handles]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.extracted = [torch.cuda.Event() for _ in range(2)]
        self.arrived = [torch.cuda.Event() for _ in range(2)]
        self.local = [g[self.rank] for g in self.gathered]       # this rank's slot: the extraction writes it in place

    def _push(self, b: int, n: int):
        """On the copy stream: wait for the chunk, meet the peers (their previous use of buffer b is over: every rank enters
        this barrier only after its own consumer released b), push, meet again (every push has landed everywhere)."""
        cur = torch.cuda.current_stream(self.device)
        self.extracted[b].record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.extracted[b])
            self.handles[b].barrier(channel=0)
            src = self.gathered[b][self.rank, :n]
            for d in range(1, self.world):
                r = (self.rank + d) % self.world                 # staggered: at any moment every peer receives from one sender
                self.peer[b][r][self.rank, :n].copy_(src, non_blocking=True)
            self.handles[b].barrier(channel=1)
            self.arrived[b].record(self.copy_stream)
        self.work[b] = self.arrived[b]

    def _wait(self, b: int):
        if self.work[b] is None:
            return
        if self.transport == 'p2p':
            torch.cuda.current_stream(self.device).wait_event(self.work[b])
        else:
            self.work[b].wait()
        self.work[b] = None

    # ---- public surface ----------------------------------------------------------------------------------------------
    def step(self, audio_chunk: torch.Tensor, gather: bool = True):
        """Extracts `audio_chunk` (n <= chunk clips) and starts its exchange.  Returns (local features (n, 7, T, F), the
        gathered buffer (world, chunk, 7, T, F) this chunk will arrive in; rows >= n of every rank's part are stale)."""
        b = self.i & 1
        self._wait(b)                             # chunk i - 2 has left this buffer pair
        n = audio_chunk.shape[0]
        if n > self.chunk:
            raise ValueError('chunk of {} clips, buffers hold {}'.format(n, self.chunk))
        out = self.ex.extract(audio_chunk, out=self.local[b][:n])
        if gather:
            if self.transport == 'p2p':
                self._push(b, n)
            else:
                # equal sizes on every rank: a short last chunk gathers the whole buffer
                self.work[b] = dist.all_gather_into_tensor(self.gathered[b].view(-1), self.local[b].view(-1), group=self.group,
                                                           async_op=True)
        self.i += 1
        return out, self.gathered[b]

    def finish(self):
        """The current stream waits for every exchange still in flight."""
        for b in range(2):
            self._wait(b)

    def verify(self, audio_chunk: torch.Tensor) -> bool:
        """One gathered chunk checked by a checksum of checksums: every rank sums the bit patterns of its own features
        (exact, order independent), the sums are exchanged, and each rank compares them with the sums of what it received."""
        n = audio_chunk.shape[0]
        out, gathered = self.step(audio_chunk)
        self.finish()
        mine = out.contiguous().view(torch.int32).sum(dtype=torch.int64).reshape(1)
        sums = torch.empty(self.world, dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(sums, mine, group=self.group)
        got = torch.stack([gathered[r, :n].contiguous().view(torch.int32).sum(dtype=torch.int64) for r in range(self.world)])
        ok = torch.tensor([int(torch.equal(got, sums) and torch.equal(gathered[self.rank, :n], out))], device=mine.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        return bool(ok.item())
