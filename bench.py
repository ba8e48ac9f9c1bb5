#!/usr/bin/env python
"""bench.py -- SALSA feature extraction throughput on B200 (audio-minutes / second).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference ...                           # the reference's CPU algorithm (oracle port)

Workload (BASELINE.json configs[1]): SALSA FOA, 600 synthetic 4-channel 24 kHz 60 s clips per GPU,
n_fft 512, hop 300.  One "step" = one pass of the hot path (audio -> (7, 4801, 200) features) over
that batch.  1 clip = 1 audio-minute, so clips/s == audio-minutes/s.

* `value`      : device-resident (audio and features stay in HBM), CUDA events, max over ranks.
* `e2e`        : the same metric through the host-buffer C-ABI entry point `salsa_extract_host`
                 (pinned host audio in, pinned host features out, H2D / kernels / D2H pipelined).
* `roofline`   : dominant kernel (stft_kernel on the default split arrangement) against the measured HBM copy
                 bandwidth in MEASURED_PEAKS.json; algorithmic bytes = the part of "audio read once + features written
                 once" that kernel moves; `roofline.path` gives the same for the whole step.
* `cpu_baseline`: the oracle (a line-by-line port of the reference's per-bin LAPACK loop) on all
                 host cores over a bounded sample of the same clips.

Under torchrun (N > 1) every rank processes its own 600 clips (weak scaling, clips are independent,
no data-path collective); NCCL is used for the barrier and the max-over-ranks time only.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS, N_FFT, HOP = 24000, 512, 300
CLIP_SECONDS = 60
N_SAMPLES = FS * CLIP_SECONDS
N_FRAMES = 1 + N_SAMPLES // HOP
AUDIO_BYTES_PER_CLIP = 4 * N_SAMPLES * 4                    # 23 040 000
METRIC = 'audio-minutes/sec SALSA feature extraction'
UNIT = 'audio-min/s'


def feature_bytes_per_clip(freq_dim):
    return 7 * N_FRAMES * freq_dim * 4                       # 26 885 600 for F = 200


# ------------------------------------------------------------------------------------------------
# synthetic clips, generated on the device (recipe of SURVEY.md section 8d / oracle/synth.py:
# two band-passed (300-6000 Hz) gated noise sources, amplitude 0.1, FOA gains or tetrahedral-array
# delays, 1e-3 sensor noise)
# ------------------------------------------------------------------------------------------------
def make_clips(torch, n_clips, audio_format, device, seed, n_samples=N_SAMPLES, chunk=40):
    out = torch.empty((n_clips, 4, n_samples), dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(2021 + seed)
    f = torch.fft.rfftfreq(n_samples, 1.0 / FS).to(device)
    # magnitude of a 4th-order Butterworth band-pass 300-6000 Hz (zero phase)
    band = 1.0 / torch.sqrt(1.0 + (f / 6000.0) ** 8) * (1.0 / torch.sqrt(1.0 + (300.0 / f.clamp(min=1e-3)) ** 8))
    t = torch.arange(n_samples, device=device, dtype=torch.float32)
    ramp = 0.005 * FS
    mic_dirs = torch.tensor([[45, 35], [-45, -35], [135, -35], [-135, 35]], dtype=torch.float32, device=device)
    mic_dirs = torch.deg2rad(mic_dirs)
    mic_unit = torch.stack([torch.cos(mic_dirs[:, 0]) * torch.cos(mic_dirs[:, 1]),
                            torch.sin(mic_dirs[:, 0]) * torch.cos(mic_dirs[:, 1]), torch.sin(mic_dirs[:, 1])], dim=1)
    for c0 in range(0, n_clips, chunk):
        b = min(chunk, n_clips - c0)
        mix = torch.zeros((b, 4, n_samples), dtype=torch.float32, device=device)
        for _src in range(2):
            S = torch.fft.rfft(torch.randn((b, n_samples), generator=g, device=device)) * band
            s = torch.fft.irfft(S, n_samples)
            s = s * (0.1 / s.std(dim=1, keepdim=True).clamp(min=1e-12))
            gate = torch.zeros((b, n_samples), device=device)
            for _seg in range(3):
                a = torch.rand((b, 1), generator=g, device=device) * (n_samples * 7 / 8)
                ln = n_samples / 8 + torch.rand((b, 1), generator=g, device=device) * (n_samples * 3 / 8)
                on = (torch.rand((b, 1), generator=g, device=device) < (1.0 if _seg == 0 else 0.5)).float()
                seg = torch.minimum((t[None] - a) / ramp, (a + ln - t[None]) / ramp).clamp(0.0, 1.0)
                gate = torch.maximum(gate, seg * on)
            s = s * gate
            az = (torch.rand((b,), generator=g, device=device) * 360.0 - 180.0) * (3.141592653589793 / 180.0)
            el = (torch.rand((b,), generator=g, device=device) * 90.0 - 45.0) * (3.141592653589793 / 180.0)
            if audio_format == 'foa':
                gains = torch.stack([torch.ones_like(az), torch.sin(az) * torch.cos(el), torch.sin(el),
                                     torch.cos(az) * torch.cos(el)], dim=1)             # W, Y, Z, X
                mix += gains[:, :, None] * s[:, None, :]
            else:
                u = torch.stack([torch.cos(az) * torch.cos(el), torch.sin(az) * torch.cos(el), torch.sin(el)], dim=1)
                tau = -0.042 * (u @ mic_unit.T) / 343.0                                  # (b, 4) seconds
                S = torch.fft.rfft(s)
                for m in range(4):
                    ph = torch.exp(-2j * 3.141592653589793 * f[None] * tau[:, m:m + 1])
                    mix[:, m] += torch.fft.irfft(S * ph, n_samples)
        mix += 1e-3 * torch.randn(mix.shape, generator=g, device=device)
        # the dataset's clips are 16-bit wav files (librosa.load -> sample / 32768, salsa_feature_extraction.py:353):
        # the synthetic clips live on the same grid, so the float32 and the 16-bit PCM entry points see the same audio
        out[c0:c0 + b] = torch.round(mix * 32768.0).clamp_(-32768.0, 32767.0) / 32768.0
    return out


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix='clocks_', suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, reasons, power = [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        with open(self.path) as fh:
            for line in fh:
                parts = [p.strip() for p in line.split(',')]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    out['sm_max_mhz'] = float(parts[2])
                    power.append(float(parts[3]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out['sm_mhz'] = sm[len(sm) // 2]
            out['power_w_max'] = max(power)
        out['reasons'] = sorted(reasons)
        out['samples'] = len(sm)
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's faithful port of the reference loop (one LAPACK SVD per selected bin)
# ------------------------------------------------------------------------------------------------
def _cpu_init():
    from oracle import salsa as osalsa  # noqa: F401  (import cost outside the clock)


def _cpu_worker(args):
    audio, fmt, fmax, batched = args
    from oracle import salsa as osalsa
    t0 = time.perf_counter()
    feat = osalsa.salsa_clip(audio, fmt, fmax_doa=fmax, batched=batched)
    return time.perf_counter() - t0, float(feat[4:].astype(bool).mean())


def cpu_baseline(clips, audio_format, fmax_doa, batched=False, cores=None):
    """clips: list of (4, n) float32 arrays (one per worker).  Returns audio-min/s over all workers."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[k] = '1'
    ctx = mp.get_context('spawn')
    jobs = [(c, audio_format, fmax_doa, batched) for c in clips]
    with ctx.Pool(min(cores, len(jobs)), initializer=_cpu_init) as pool:
        pool.map(abs, range(4 * cores))                              # workers up, imports done
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    minutes = sum(c.shape[1] for c in clips) / FS / 60.0
    return minutes / wall, wall, min(cores, len(jobs)), sum(r[1] for r in res) / len(res)


def host_sample_clips(n_workers, seconds, audio_format, seed=0):
    """The same recipe as the device generator, drawn on the host by oracle.synth (bounded sample)."""
    from oracle import synth
    return [synth.make_clip(seed * 1000 + i, audio_format, seconds=seconds) for i in range(n_workers)]


# ------------------------------------------------------------------------------------------------
def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def reference_sample_clips(args, cores):
    """The bounded sample both arms time on the CPU: one window of --cpu-seconds per host core, cut at offsets spread over
    the clip length out of the FIRST clips of the benchmark batch (rank 0's generator, first chunk of 40 clips: the same
    audio the CUDA arm processes).  Falls back to the host generator (same recipe, other draws) without a CUDA device."""
    import numpy as np
    n_win = int(args.cpu_seconds * FS)
    n = min(cores, 40)
    try:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('no CUDA device')
        audio = make_clips(torch, 40, args.format, torch.device('cuda', 0), seed=0)[:n]
        clips = []
        for i in range(n):
            off = int((i + 0.5) / n * max(1, N_SAMPLES - n_win))
            clips.append(np.ascontiguousarray(audio[i, :, off:off + n_win].cpu().numpy()))
        del audio
        torch.cuda.empty_cache()
        return clips, 'windows of the first {} benchmark clips'.format(n)
    except Exception as exc:                                   # noqa: BLE001 -- a CPU-only box still gets its baseline
        return host_sample_clips(n, args.cpu_seconds, args.format), 'oracle.synth clips ({})'.format(repr(exc)[:60])


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is Python + librosa and cannot
    travel to the GPU box) on all host cores.  Each step = one bounded sample of the benchmark workload."""
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    fmax = 9000 if args.format == 'foa' else 4000
    clips, origin = reference_sample_clips(args, cores)
    vals = []
    for step in range(args.warmup + args.steps):
        v, wall, used, valid = cpu_baseline(clips, args.format, fmax, batched=False, cores=cores)
        if step >= args.warmup:
            vals.append((v, wall))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(w for _, w in vals) / len(vals)
    sample = ('{} x {:.0f} s {} per step (one per core, offsets spread over the clip), valid-bin fraction {:.2f}; oracle loop '
              'form = the reference algorithm (1 LAPACK SVD per selected bin); rate extrapolates linearly to the 600-clip batch'
              ).format(len(clips), args.cpu_seconds, origin, valid)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, world),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': used, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, world):
    """Identical in both arms (the driver compares them): the workload, and how the CPU arm samples it."""
    lite = getattr(args, 'feature', 'salsa') == 'salsa_lite'
    return {
        'workload': '{} {} batch: {} synthetic 4-ch 24 kHz 60 s clips per GPU, n_fft=512 hop=300 '
                    '(BASELINE.json configs[{}])'.format('SALSA-Lite' if lite else 'SALSA', args.format.upper(), args.clips, 2 if lite else 1),
        'clips_per_gpu': args.clips, 'clips_total': args.clips * world, 'audio_format': args.format,
        'stft_precision': args.stft_precision,
        'l2_policy': 'inputs larger than L2 ({:.1f} GB audio per GPU per step)'.format(args.clips * AUDIO_BYTES_PER_CLIP / 1e9),
        'parallelism': 'clips sharded over {} GPU(s), no data-path collective in the headline step'.format(world),
        'cpu_arm_sample': '{:.0f} s windows of the first benchmark clips, one per host core per step (rate extrapolated)'.format(args.cpu_seconds),
    }


# ------------------------------------------------------------------------------------------------
# second metric of BASELINE.json: CRNN clips/sec (full-clip inference on the extracted features)
# ------------------------------------------------------------------------------------------------
CRNN_CONV_FLOP_PER_CLIP = 2 * 22.368e9 * 7.5            # SURVEY.md appendix A x (4800 / 640)


def _crnn_cpu_chunk(_):
    import torch
    from oracle import crnn as ocrnn
    sd = ocrnn.make_state_dict(0)
    x = ocrnn.model_input(1, (1, 7, 640, 200))
    t0 = time.perf_counter()
    ocrnn.forward(sd, x)
    return time.perf_counter() - t0


class numa_local:
    """While pinned host buffers are allocated: run on the CPUs NVML reports as local to this rank's GPU, so that the pages
    land on the GPU's NUMA node (first touch).  With 8 ranks streaming through host memory at once that is the difference
    between every copy crossing the socket interconnect and none.  Best effort: any failure leaves the affinity alone."""

    def __init__(self, device_index):
        self.device_index, self.saved = device_index, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device_index)
            n_words = (os.cpu_count() + 63) // 64
            words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
            allowed = os.sched_getaffinity(0)
            if cpus & allowed and (cpus & allowed) != allowed:
                self.saved = allowed
                os.sched_setaffinity(0, cpus & allowed)
        except Exception:
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except OSError:
                pass
        return False


def bench_pipeline(args, torch, dist, audio, rank, world, dev):
    """Audio in, SELD outputs out (salsa_b200.SeldPipeline: features on the fly, nothing written in between): device-resident
    and end to end from pinned host audio (H2D of the next batch overlapped, logits copied back every step)."""
    import salsa_b200
    B = min(args.crnn_batch, audio.shape[0])
    model = salsa_b200.SeldModel(salsa_b200.PannResNet22(n_input_channels=7),
                                 salsa_b200.SeldDecoder(512, n_classes=12, output_format='reg_xyz', decoder_type='bigru',
                                                        freq_pool='avg', decoder_size=256), label_rate=10, feature_rate=80.0)
    model.load_state_dict(salsa_b200.crnn.random_state_dict(0))
    pipe = salsa_b200.SeldPipeline(salsa_b200.SalsaExtractor(args.format, fmax_doa=9000 if args.format == 'foa' else 4000), model)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    a = audio[:B]
    for _ in range(2):
        out = pipe(a)
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        out = pipe(a)
    stop.record()
    barrier()
    t = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value = B * world * args.steps / (float(t.item()) / 1e3)
    # end to end from the wav files' own 16-bit samples (the synthetic clips lie on that grid): the device converts them
    pcm = torch.round(a * 32768.0).clamp_(-32768, 32767).to(torch.int16)
    same = bool(torch.equal(pcm.float() / 32768.0, a))
    with numa_local(dev.index or 0):
        h_a = torch.empty(tuple(a.shape), dtype=torch.int16, pin_memory=True)
        h_a.copy_(pcm)
        h_out = {k: torch.empty(v.shape, dtype=torch.float32, pin_memory=True) for k, v in out.items()}
    d_a = [torch.empty_like(pcm) for _ in range(2)]
    del pcm
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])
            d_a[i & 1].copy_(h_a, non_blocking=True)
            ready[i & 1].record(copy_stream)

    for e in consumed:
        e.record()
    barrier()
    n_steps = max(args.e2e_steps, 10)        # the first upload is not overlapped: enough steps that the fill is < 10 % of the run
    t0 = time.perf_counter()
    upload(0)
    for i in range(n_steps):
        if i + 1 < n_steps:
            upload(i + 1)
        torch.cuda.current_stream().wait_event(ready[i & 1])
        o = pipe(d_a[i & 1])
        consumed[i & 1].record()
        for k in o:
            h_out[k].copy_(o[k], non_blocking=True)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    return {'metric': 'clips/sec audio -> SALSA features -> CRNN outputs (SeldPipeline, 60 s clips)', 'value': value, 'unit': 'clips/s',
            'batch_per_gpu': B,
            'e2e': {'value': B * world * n_steps / float(te.item()), 'unit': 'clips/s', 'h2d_bytes_per_step': int(h_a.numel() * 2),
                    'd2h_bytes_per_step': int(sum(v.numel() for v in h_out.values()) * 4), 'clips_per_step_per_gpu': B,
                    'pcm16_equals_float_audio': same, 'steps': n_steps,
                    'note': '16-bit PCM audio (11.5 MB per clip) crosses PCIe instead of features (27 MB per clip); nothing is written in between'}}


def bench_crnn(args, torch, dist, feat, rank, world, dev, peaks):
    import salsa_b200
    from salsa_b200 import _native
    B = min(args.crnn_batch, feat.shape[0])
    model = salsa_b200.SeldModel(salsa_b200.PannResNet22(n_input_channels=7),
                                 salsa_b200.SeldDecoder(512, n_classes=12, output_format='reg_xyz', decoder_type='bigru',
                                                        freq_pool='avg', decoder_size=256), label_rate=10, feature_rate=80.0)
    model.load_state_dict(salsa_b200.crnn.random_state_dict(0))
    x = feat[:B]
    T = (x.shape[2] // 16) * 16                           # 4801 -> 4800 (database.py:205-207)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = model.forward(x, n_frames=T)
    barrier()
    _native.lib().salsa_launch_count(1)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        out = model.forward(x, n_frames=T)
    stop.record()
    barrier()
    launches = int(_native.lib().salsa_launch_count(0))
    t = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    value = B * world / (ms / 1e3)
    # end to end: pinned host features in, host logits out; the copy of step i+1 runs on a second stream while
    # step i computes (two device input buffers)
    nb = B
    with numa_local(dev.index or 0):
        h_x = torch.empty((nb,) + tuple(x.shape[1:]), dtype=torch.float32, pin_memory=True)
        h_x.copy_(x[:nb])
        h_out = {k: torch.empty(v[:nb].shape, dtype=torch.float32, pin_memory=True) for k, v in out.items()}
    d_x = [torch.empty_like(x[:nb]) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])
            d_x[i & 1].copy_(h_x, non_blocking=True)
            ready[i & 1].record(copy_stream)

    for e in consumed:
        e.record()
    upload(0)
    torch.cuda.current_stream().wait_event(ready[0])
    model.forward(d_x[0], n_frames=T)
    consumed[0].record()
    barrier()
    n_e2e = max(args.e2e_steps, 10)          # the first upload is not overlapped: enough steps that the fill is < 10 % of the run
    t0 = time.perf_counter()
    upload(0)
    for i in range(n_e2e):
        if i + 1 < n_e2e:
            upload(i + 1)
        torch.cuda.current_stream().wait_event(ready[i & 1])
        o = model.forward(d_x[i & 1], n_frames=T)
        consumed[i & 1].record()
        for k in o:
            h_out[k].copy_(o[k], non_blocking=True)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = nb * world * n_e2e / float(te.item())
    achieved = value / world * CRNN_CONV_FLOP_PER_CLIP / 1e12
    res = {
        'metric': 'CRNN clips/sec (ResNet22 + BiGRU forward, full 60 s clips)', 'value': value, 'unit': 'clips/s',
        'ms_per_step': ms, 'dtype': 'bf16', 'config': {'batch_per_gpu': B, 'input': [B, 7, T, int(x.shape[3])],
                                                        'weights': 'random init (reference state-dict keys)'},
        'gpu_launches': launches,
        'e2e': {'value': e2e, 'unit': 'clips/s', 'h2d_bytes_per_step': int(h_x.numel() * 4),
                'd2h_bytes_per_step': int(sum(v.numel() for v in h_out.values()) * 4), 'clips_per_step_per_gpu': nb, 'steps': n_e2e},
        'roofline': {'bound': 'tensor', 'kernel': 'conv_tc_kernel (all 22 convolutions)', 'achieved': achieved,
                     'peak': peaks.get('bf16_tflops_sustained', 1400.0), 'unit': 'TFLOP/s',
                     'frac': achieved / peaks.get('bf16_tflops_sustained', 1400.0), 'traffic': None,
                     'note': 'algorithmic conv FLOPs (335.5 GFLOP per clip) over the whole forward time, GRU and heads included in the time'},
    }
    # the float32-parity modes (operands as sums of bf16 planes on the same tensor-core kernels), for the record:
    # bf16x2 = two planes, three plane products per MAC; bf16x3 = three planes, six products
    del model, out
    torch.cuda.empty_cache()
    for precision, bp, note in (('bf16x2', min(B, 16), 'logits 6-8e-5 of the output scale from the float32 reference (bar 1e-4), 3x the MMA work'),
                                ('bf16x3', min(B, 8), 'logits 3-4e-5 from the float32 reference, 6x the MMA work')):
        mp = salsa_b200.SeldModel(salsa_b200.PannResNet22(n_input_channels=7),
                                  salsa_b200.SeldDecoder(512, n_classes=12, output_format='reg_xyz', decoder_type='bigru',
                                                         freq_pool='avg', decoder_size=256), precision=precision)
        mp.load_state_dict(salsa_b200.crnn.random_state_dict(0))
        xp = x[:bp]
        msp = time_steps(torch, dist, world, dev, lambda: mp.forward(xp, n_frames=T), 2, 3)
        res[precision + '_parity_mode'] = {'value': bp * world / (msp / 1e3), 'unit': 'clips/s', 'batch_per_gpu': bp, 'note': note}
        del mp
        torch.cuda.empty_cache()
    if rank == 0 and not args.no_cpu_baseline:
        # stock PyTorch / cuDNN on the same GPU, for the convolution stack only (the encoder holds 99.5 % of
        # the FLOPs; the oracle's step-by-step GRU is a Python loop and would be unfair), channels_last tensors, in the
        # three arithmetic modes that sit beside ours: bf16 autocast (beside bf16), TF32 and strict fp32 (beside the
        # parity modes).  Baselines beside the number, like the CPU leg: they execute the oracle's functional encoder.
        try:
            from oracle import crnn as ocrnn
            sd = {k: v.to(dev) for k, v in salsa_b200.crnn.random_state_dict(0).items() if k.startswith('encoder.')}
            sd = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
            bc = min(B, 8)
            xc = x[:bc, :, :T].contiguous(memory_format=torch.channels_last)
            base = {}
            for label, tf32, autocast in (('bf16_autocast', True, True), ('tf32', True, False), ('fp32', False, False)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.benchmark = True
                with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
                    msb = time_steps(torch, dist, 1, dev, lambda: ocrnn.encoder_forward(sd, xc), 2, 3)
                base[label] = bc / (msb / 1e3)
            res['cudnn_encoder_baselines'] = {
                'clips_per_s': base, 'batch': bc,
                'note': 'torch {} F.conv2d / batch_norm / avg_pool2d, channels_last, cudnn.benchmark on, encoder only (no GRU, no heads)'.format(
                    torch.__version__)}
            del sd, xc
            torch.cuda.empty_cache()
        except Exception as exc:          # the baseline must never take the product line down
            res['cudnn_encoder_baselines'] = {'unavailable': repr(exc)[:200]}
        import torch as _t
        tc = _crnn_cpu_chunk(None)
        tc = min(tc, _crnn_cpu_chunk(None))
        res['cpu_baseline'] = {'value': 1.0 / (7.5 * tc), 'unit': 'clips/s', 'cores': _t.get_num_threads(), 'kind': 'port',
                               'sample': 'oracle fp32 forward (stock PyTorch CPU ops) of one (1,7,640,200) chunk, {:.2f} s; '
                                         'a clip is 7.5 chunks'.format(tc)}
    return res


def time_steps(torch, dist, world, dev, fn, warmup, steps):
    """W warm-up calls, then K calls between two CUDA events on the current stream, bracketed by barrier + synchronize;
    returns the max over ranks of the milliseconds per step."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        fn()
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        fn()
    stop.record()
    barrier()
    t = torch.tensor([start.elapsed_time(stop) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_host_link(torch, dist, world, dev, ne, feat_bytes, chunk_clips, steps):
    """Ceiling of the end-to-end leg: the same bytes as one `extract_host` step (ne clips of audio in, ne clips of features
    out, in chunks of `chunk_clips`) as plain pinned cudaMemcpyAsync on two streams with NO kernels, all ranks at once."""
    out = {}
    for label, in_bytes in (('pcm16', AUDIO_BYTES_PER_CLIP // 2), ('f32', AUDIO_BYTES_PER_CLIP)):
        with numa_local(dev.index or 0):
            h_in = torch.empty(ne * in_bytes, dtype=torch.uint8, pin_memory=True)
            h_out = torch.empty(ne * feat_bytes, dtype=torch.uint8, pin_memory=True)
            h_in.zero_()
            h_out.zero_()
        d_in = torch.empty(chunk_clips * in_bytes, dtype=torch.uint8, device=dev)
        d_out = torch.empty(chunk_clips * feat_bytes, dtype=torch.uint8, device=dev)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def step():
            for c0 in range(0, ne, chunk_clips):
                n = min(chunk_clips, ne - c0)
                with torch.cuda.stream(s_in):
                    d_in[:n * in_bytes].copy_(h_in[c0 * in_bytes:(c0 + n) * in_bytes], non_blocking=True)
                with torch.cuda.stream(s_out):
                    h_out[c0 * feat_bytes:(c0 + n) * feat_bytes].copy_(d_out[:n * feat_bytes], non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()

        step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        sec = float(te.item()) / steps
        out[label] = {'clips_per_s_ceiling': ne * world / sec, 'h2d_gbs_per_gpu': ne * in_bytes / sec / 1e9,
                      'd2h_gbs_per_gpu': ne * feat_bytes / sec / 1e9}
        del h_in, h_out, d_in, d_out
    return out


def bench_other_configs(args, torch, dist, rank, world, dev, peak):
    """BASELINE.json configs[2] (SALSA-Lite MIC), the MIC full-eigenvector path of configs[3] on this GPU count, and -- with
    more than one rank -- configs[3]'s data plane: every rank's features all-gathered over NVLink (NCCL) in chunks of clips,
    the gather of chunk i overlapped with the extraction of chunk i + 1; reported with and without the gather."""
    import salsa_b200
    n = args.clips
    audio = make_clips(torch, n, 'mic', dev, seed=7000 + 1000 * rank)
    res = {}
    # ---- configs[2]: SALSA-Lite MIC
    lite = salsa_b200.SalsaLiteExtractor('salsa_lite')
    feat = torch.empty((n, 7, N_FRAMES, lite.freq_dim), dtype=torch.float32, device=dev)
    ms = time_steps(torch, dist, world, dev, lambda: lite.extract(audio, out=feat), 2, max(3, args.steps // 4))
    lite_bytes = AUDIO_BYTES_PER_CLIP + feature_bytes_per_clip(lite.freq_dim)
    res['salsa_lite_mic'] = {'config': 'configs[2]: SALSA-Lite MIC, {} clips per GPU'.format(n), 'value': n * world / (ms / 1e3),
                             'unit': UNIT, 'ms_per_step': ms, 'roofline_frac': n * lite_bytes / (ms / 1e3) / 1e9 / peak}
    del feat
    # ---- configs[3] without the gather: SALSA MIC full eigenvector
    ex = salsa_b200.SalsaExtractor('mic', fmax_doa=4000)
    feat = torch.empty((n, 7, N_FRAMES, ex.freq_dim), dtype=torch.float32, device=dev)
    ms = time_steps(torch, dist, world, dev, lambda: ex.extract(audio, out=feat), 2, max(3, args.steps // 4))
    fbytes = feature_bytes_per_clip(ex.freq_dim)
    res['salsa_mic'] = {'config': 'configs[3] compute: SALSA MIC full eigenvector, {} clips per GPU, no gather'.format(n),
                        'value': n * world / (ms / 1e3), 'unit': UNIT, 'ms_per_step': ms,
                        'roofline_frac': n * (AUDIO_BYTES_PER_CLIP + fbytes) / (ms / 1e3) / 1e9 / peak}
    del feat
    torch.cuda.empty_cache()
    if world == 1:
        return res
    # ---- configs[3] with its data plane
    from salsa_b200.sharding import ChunkedFeatureGather
    chunk = min(args.gather_chunk, n)
    n_chunks = (n + chunk - 1) // chunk
    recv = (world - 1) * n * fbytes
    ms_plain = None
    for transport in ('auto', 'nccl'):
        gather = ChunkedFeatureGather(ex, chunk, N_SAMPLES, dev, transport=transport)
        if transport == 'auto' and gather.transport == 'nccl':
            res['salsa_mic_gather_p2p'] = {'unavailable': gather.fallback_reason}
            del gather
            continue
        ok = gather.verify(audio[:chunk])                    # bit-level checksum of every rank's chunk after the exchange

        def chunked(with_gather):
            for c in range(n_chunks):
                gather.step(audio[c * chunk:(c + 1) * chunk], gather=with_gather)
            gather.finish()

        if ms_plain is None:
            ms_plain = time_steps(torch, dist, world, dev, lambda: chunked(False), 1, 3)
        ms_gather = time_steps(torch, dist, world, dev, lambda: chunked(True), 1, 3)
        res['salsa_mic_gather_' + gather.transport] = {
            'config': 'configs[3]: SALSA MIC full eigenvector, {} clips over {} GPUs, features of every rank gathered on every rank in '
                      'chunks of {} clips, the exchange of chunk i overlapped with the extraction of chunk i + 1'.format(n * world, world, chunk),
            'transport': 'copy-engine pushes into symmetric (peer-mapped) buffers over NVLink' if gather.transport == 'p2p'
                         else 'NCCL all_gather_into_tensor',
            'value': n * world / (ms_gather / 1e3), 'unit': UNIT, 'ms_per_step': ms_gather,
            'value_without_gather': n * world / (ms_plain / 1e3), 'ms_per_step_without_gather': ms_plain,
            'nvlink_bytes_received_per_gpu': recv, 'gather_gbs_received_per_gpu': recv / (ms_gather / 1e3) / 1e9,
            'gathered_checksums_match': bool(ok)}
        del gather
        torch.cuda.empty_cache()
    return res


def bench_train(args, torch, dist, rank, world, dev, peaks=None):
    """BASELINE.json configs[4]: CRNN training on on-the-fly SALSA features, bf16, data-parallel.  Per step and rank: 32 audio
    chunks of 8 s -> SALSA FOA features (native) -> channel-swap / frequency-shift augmentation (native) -> forward + backward
    (3x3 and 1x1 convolutions in all three directions, train-mode BatchNorm + residual + ReLU + dropout, pooling and the BiGRU native; heads: torch autograd / cuBLAS) ->
    loss (native) -> bucketed bf16 gradient all-reduce overlapped with the backward pass (NCCL) -> Adam (native)."""
    import numpy as np
    import salsa_b200
    from salsa_b200 import augment, train
    B, n_samp = args.train_batch, 8 * FS
    audio = make_clips(torch, B, 'foa', dev, seed=9000 + 1000 * rank, n_samples=n_samp, chunk=B)
    ex = salsa_b200.SalsaExtractor('foa')
    aug = augment.BatchAugment(augment.TfmapRandomSwapChannelFoa(n_classes=12), augment.RandomShiftUpDownNp(freq_shift_range=10))
    sched = salsa_b200.optim.LearningRateScheduler(steps_per_epoch=100, max_epochs=50)
    tr = train.SeldTrainer(salsa_b200.crnn.random_state_dict(0), scheduler=sched, device=dev, use_graph=not args.no_train_graph)
    g = torch.Generator(device=dev)
    g.manual_seed(77 + rank)
    tgt = {'event_frame_gt': (torch.rand((B, 80, 12), generator=g, device=dev) > 0.8).float(),
           'doa_frame_gt': torch.rand((B, 80, 36), generator=g, device=dev) * 2 - 1}
    np.random.seed(1234 + rank)
    losses = []

    def step():
        feat = ex.extract(audio)[:, :, :640]                 # 641 -> 640 frames (database.py:205-207)
        x, _, y_doa = aug(feat, tgt['event_frame_gt'], tgt['doa_frame_gt'])
        losses.append(tr.step(x, {'event_frame_gt': tgt['event_frame_gt'], 'doa_frame_gt': y_doa}))

    ms = time_steps(torch, dist, world, dev, step, 3, max(3, args.steps))
    first, last = float(losses[0][0].item()), float(losses[-1][0].item())
    # the feature part alone, for the split
    ms_feat = time_steps(torch, dist, world, dev, lambda: ex.extract(audio), 1, 3)
    n_params = int(tr.flat.numel())
    # forward + input gradient + weight gradient = 3 x the forward's convolution FLOPs (44.74 GFLOP per chunk)
    tflops = 3 * 2 * 22.368e9 * B / (ms / 1e3) / 1e12
    peak = (peaks or {}).get('bf16_tflops_sustained', 1400.0)
    return {'config': 'configs[4]: CRNN (ResNet22 + BiGRU) training step on on-the-fly SALSA FOA features, bf16 autocast, batch {} x (7, 640, 200) '
                      'per GPU, {} GPU(s) data-parallel'.format(B, world),
            'value': B * world / (ms / 1e3), 'unit': 'chunks/s (8 s each)', 'audio_min_per_s': B * world * 8 / 60.0 / (ms / 1e3),
            'ms_per_step': ms, 'ms_features': ms_feat,
            'roofline': {'bound': 'tensor', 'achieved': tflops, 'peak': peak, 'unit': 'TFLOP/s', 'frac': tflops / peak,
                         'note': 'algorithmic convolution FLOPs of forward + dgrad + wgrad per GPU over the whole step (features, BatchNorm, GRU, Adam included in the time)'}, 'cuda_graph': ('off' if args.no_train_graph else (tr.graph_error or 'forward + loss + backward replayed as one CUDA graph; all-reduce of the flat gradient (one NCCL call) and Adam follow it')), 'loss_first_step': first, 'loss_last_step': last, 'parameters': n_params,
            'allreduce': ('bf16 all-reduce of the flat gradient ({:.1f} MB on the wire per step), '.format(n_params * 2 / 1e6) +
                          ('one call after the graph replay' if (tr.use_graph and tr.graph_error is None) else 'launched per bucket during the backward pass'))
                         if world > 1 else 'single rank: none',
            'native': 'SALSA features, augmentation, every convolution forward + input gradient + weight gradient (tcgen05), train-mode BatchNorm + residual + ReLU + dropout and 2x2 pooling forward / backward, BiGRU recurrence + back-propagation through time, loss, Adam',
            'library': 'the heads (nn.Linear pairs with dropout) and the GEMMs around the GRU recurrence: torch autograd / cuBLAS; no cuDNN call in the step'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--clips', type=int, default=600, help='clips per GPU per step')
    ap.add_argument('--format', default='foa', choices=['foa', 'mic'])
    ap.add_argument('--feature', default='salsa', choices=['salsa', 'salsa_lite'],
                    help='salsa = BASELINE configs[1] (default); salsa_lite = configs[2] (SALSA-Lite MIC, no CRNN / CPU legs)')
    ap.add_argument('--stft-precision', type=int, default=64, choices=[32, 64])
    ap.add_argument('--e2e-clips', type=int, default=120, help='clips per GPU per end-to-end step (host buffers)')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--e2e-chunk', type=int, default=8, help='clips per chunk of the host pipeline')
    ap.add_argument('--cpu-seconds', type=float, default=5.0, help='clip length of the CPU baseline sample')
    ap.add_argument('--gather-chunk', type=int, default=50, help='clips per all-gather chunk of the configs[3] leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-crnn', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true', help='skip the SALSA-Lite / MIC / gather legs')
    ap.add_argument('--no-fast-mode', action='store_true', help='skip the float32-FFT comparison run')
    ap.add_argument('--crnn-batch', type=int, default=32, help='clips per CRNN forward per GPU')
    ap.add_argument('--train-batch', type=int, default=32, help='8 s chunks per training step per GPU')
    ap.add_argument('--no-train', action='store_true', help='skip the configs[4] training-step leg')
    ap.add_argument('--no-train-graph', action='store_true', help='training step as single launches instead of one CUDA graph replay')
    args = ap.parse_args()

    rank, world, local_rank = env_int('RANK', 0), env_int('WORLD_SIZE', 1), env_int('LOCAL_RANK', 0)
    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import salsa_b200
    from salsa_b200 import _native

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    peaks = json.load(open(peaks_path)) if os.path.isfile(peaks_path) else {}
    if 'hbm_gbs' in peaks:
        peak, peak_src = float(peaks['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (measured copy)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'

    fmax = 9000 if args.format == 'foa' else 4000
    if args.feature == 'salsa_lite':
        args.format, args.no_crnn, args.no_cpu_baseline, args.no_other_configs = 'mic', True, True, True
        ex_kwargs = dict(feature_type='salsa_lite', stft_precision=args.stft_precision)
        ex = salsa_b200.SalsaLiteExtractor(**ex_kwargs)
    else:
        ex_kwargs = dict(audio_format=args.format, fmax_doa=fmax, stft_precision=args.stft_precision)
        ex = salsa_b200.SalsaExtractor(**ex_kwargs)
    n_clips = args.clips
    audio = make_clips(torch, n_clips, args.format, dev, seed=1000 * rank)
    feat = torch.empty((n_clips, 7, N_FRAMES, ex.freq_dim), dtype=torch.float32, device=dev)
    feat_bytes = feature_bytes_per_clip(ex.freq_dim)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident: W warm-up steps, K timed steps --------------------------------------
    for _ in range(args.warmup):
        ex.extract(audio, out=feat)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _native.lib().salsa_launch_count(1)
    _native.profile_enable(True)
    _native.profile_read()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(args.steps):
        ex.extract(audio, out=feat)
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    launches = int(_native.lib().salsa_launch_count(0))
    kernels = _native.profile_read()
    _native.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = n_clips * world / (ms_per_step / 1e3)
    valid_frac = float((feat[: min(n_clips, 8), 4:, :, :ex.upper_bin - ex.lower_bin] != 0).float().mean().item())

    # ---- the float32-FFT variant, for the record: speed and how far it is from the float64-FFT features ----
    fast = None
    if args.stft_precision == 64 and not args.no_fast_mode:
        ex32 = type(ex)(**{**ex_kwargs, 'stft_precision': 32})
        feat32 = torch.empty_like(feat)
        t32 = time_steps(torch, dist, world, dev, lambda: ex32.extract(audio, out=feat32), 1, 2)
        n_sp = 4
        mism = int(((feat32[:, n_sp:] != 0) != (feat[:, n_sp:] != 0)).sum().item())
        both = (feat32[:, n_sp:] != 0) & (feat[:, n_sp:] != 0)
        fast = {'stft_precision': 32, 'value': n_clips * world / (t32 / 1e3), 'unit': UNIT, 'ms_per_step': t32,
                'valid_bin_mask_mismatches_vs_fp64_fft': mism, 'bins_compared': int(feat[:, n_sp:].numel()),
                'max_abs_diff_spectrogram_db': float((feat32[:, :n_sp] - feat[:, :n_sp]).abs().max().item()),
                'max_abs_diff_spatial_on_common_bins': float(((feat32[:, n_sp:] - feat[:, n_sp:]).abs() * both).max().item())}
        del feat32, both

    # ---- end to end through the host-buffer entry points --------------------------------------
    e2e = None
    if not args.no_e2e:
        ne = min(args.e2e_clips, n_clips)
        is_salsa = args.feature == 'salsa'
        with numa_local(dev.index or 0):
            h_audio = torch.empty((ne, 4, N_SAMPLES), dtype=torch.float32, pin_memory=True)
            h_audio.copy_(audio[:ne])
            h_pcm = torch.empty((ne, 4, N_SAMPLES), dtype=torch.int16, pin_memory=True)
            h_pcm.copy_(torch.round(audio[:ne] * 32768.0).to(torch.int16))          # exact: the clips live on the 16-bit grid
            h_feat = torch.empty((ne, 7, N_FRAMES, ex.freq_dim), dtype=torch.float32, pin_memory=True)
            h_feat.zero_()                                                   # first touch under the local affinity

        def e2e_run(h_in):
            ex.extract_host(h_in, out=h_feat, clips_per_chunk=args.e2e_chunk)             # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                ex.extract_host(h_in, out=h_feat, clips_per_chunk=args.e2e_chunk)        # synchronous
            torch.cuda.synchronize()
            te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(h_feat[0].to(dev), feat[0]) or torch.allclose(h_feat[0].to(dev), feat[0], equal_nan=True))
            return ne * world * args.e2e_steps / float(te.item()), same

        v32, same32 = e2e_run(h_audio)
        if is_salsa:
            v16, same16 = e2e_run(h_pcm)
        link = bench_host_link(torch, dist, world, dev, ne, feat_bytes, args.e2e_chunk, args.e2e_steps)
        if is_salsa:
            e2e = {'value': v16, 'unit': UNIT, 'h2d_bytes_per_step': ne * AUDIO_BYTES_PER_CLIP // 2, 'd2h_bytes_per_step': ne * feat_bytes,
                   'clips_per_step_per_gpu': ne, 'steps': args.e2e_steps, 'matches_device_path': same16 and same32,
                   'api': 'SalsaExtractor.extract_host(int16 PCM) -> salsa_extract_host_pcm16 (pinned host buffers: the wav files\' 16-bit '
                          'samples in, float32 features out; 3-stream pipeline)',
                   'value_f32_input': v32, 'h2d_bytes_per_step_f32_input': ne * AUDIO_BYTES_PER_CLIP,
                   'host_link_ceiling': link['pcm16']['clips_per_s_ceiling'], 'frac_of_host_link': v16 / link['pcm16']['clips_per_s_ceiling'],
                   'host_link': link}
        else:
            e2e = {'value': v32, 'unit': UNIT, 'h2d_bytes_per_step': ne * AUDIO_BYTES_PER_CLIP, 'd2h_bytes_per_step': ne * feat_bytes,
                   'clips_per_step_per_gpu': ne, 'steps': args.e2e_steps, 'matches_device_path': same32,
                   'api': 'SalsaLiteExtractor.extract_host -> salsa_lite_extract_host (pinned host buffers, 3-stream pipeline)',
                   'host_link_ceiling': link['f32']['clips_per_s_ceiling'], 'frac_of_host_link': v32 / link['f32']['clips_per_s_ceiling'],
                   'host_link': link}
        del h_audio, h_feat, h_pcm
        _native.lib().salsa_host_release()

    cpu_clips = None
    if rank == 0 and not args.no_cpu_baseline:
        # bounded CPU sample: one window of cpu_seconds per host core, taken at offsets spread over the
        # clip length (clip starts are often silent, which would flatter the CPU's masked-bin loop)
        cores = os.cpu_count() or 1
        n_win = int(args.cpu_seconds * FS)
        nc = min(cores, n_clips, 40)
        cpu_clips = []
        for i in range(nc):
            off = int((i + 0.5) / nc * max(1, N_SAMPLES - n_win))
            cpu_clips.append(np.ascontiguousarray(audio[i, :, off:off + n_win].cpu().numpy()))
    crnn = None
    pipeline = None
    if not args.no_crnn:
        pipeline = bench_pipeline(args, torch, dist, audio, rank, world, dev) if args.feature == 'salsa' else None
    del audio
    ex._workspace = None
    torch.cuda.empty_cache()
    if not args.no_crnn:
        crnn = bench_crnn(args, torch, dist, feat, rank, world, dev, peaks)
        crnn['pipeline'] = pipeline
    del feat
    torch.cuda.empty_cache()
    other = None
    if not args.no_other_configs:
        other = bench_other_configs(args, torch, dist, rank, world, dev, peak)
    train_leg = None
    if not args.no_train and args.feature == 'salsa':
        torch.cuda.empty_cache()
        train_leg = bench_train(args, torch, dist, rank, world, dev, peaks)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ------------------------------------------------------
    dom = max(kernels.items(), key=lambda kv: kv[1][0]) if kernels else (None, (0.0, 0))
    total_kernel_ms = sum(v[0] for v in kernels.values())
    # algorithmic bytes (audio read once + features written once, 49 925 600 B per SALSA clip) split over the
    # kernels that move them: stft_kernel reads the audio and writes the 4 spectrogram channels, eig_tile_kernel writes
    # the 3 spatial channels; the fused kernels move everything.  X and the masks between the kernels are NOT counted.
    path_bytes = n_clips * (AUDIO_BYTES_PER_CLIP + feat_bytes)
    eig_name = next((k for k in kernels if k.startswith('eig_') and k != 'eig_redo_kernel'), None)
    share = {'stft_kernel': (AUDIO_BYTES_PER_CLIP + feat_bytes * 4 // 7) if eig_name else 0}
    if eig_name:
        share[eig_name] = feat_bytes * 3 // 7
    roofline = None
    if dom[0]:
        avg_ms = dom[1][0] / dom[1][1]
        algo_bytes = n_clips * share.get(dom[0], AUDIO_BYTES_PER_CLIP + feat_bytes)
        achieved = algo_bytes / (avg_ms / 1e3) / 1e9
        path_achieved = path_bytes / (ms_per_step / 1e3) / 1e9
        roofline = {'bound': 'hbm', 'kernel': dom[0], 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': algo_bytes, 'avg_launch_ms': avg_ms,
                    'share_of_step': dom[1][0] / total_kernel_ms,
                    # the whole step (all kernels): audio read once + features written once over ms_per_step
                    'path_bytes_per_step': path_bytes, 'path_achieved': path_achieved, 'path_frac': path_achieved / peak}
        for k, v in kernels.items():
            roofline['ms_' + k] = round(v[0] / args.steps, 4)
        traffic_path = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.isfile(traffic_path):
            try:
                tr = json.load(open(traffic_path)).get(dom[0])
                if tr:
                    # ncu measured bytes per clip at its own (smaller) batch; scaled to this launch
                    roofline['traffic'] = tr['dram_bytes_per_clip'] * n_clips
                    roofline['traffic_note'] = 'ncu capture of {} clips scaled to {}'.format(tr.get('clips', '?'), n_clips)
            except (ValueError, KeyError):
                pass

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, wall, used, vfrac = cpu_baseline(cpu_clips, args.format, fmax, batched=False, cores=cores)
        vb, wallb, _, _ = cpu_baseline(cpu_clips, args.format, fmax, batched=True, cores=cores)
        cpu = {'value': v, 'unit': UNIT, 'cores': used, 'kind': 'port',
               'sample': '{:.0f} s windows of {} of the benchmark clips (offsets spread over the clip), one per core, '
                         '{:.1f} s wall, valid-bin fraction {:.2f}; oracle loop form (1 LAPACK SVD per selected bin, as the '
                         'reference)'.format(args.cpu_seconds, len(cpu_clips), wall, vfrac),
               'value_stacked_lapack': vb}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32 (covariance/eigenvector) + f64 (STFT, tracker)' if args.stft_precision == 64 else 'f32 (+ f64 tracker)',
        'data': 'synthetic', 'config': workload_config(args, world), 'valid_bin_fraction': round(valid_frac, 4),
        'clocks': clocks, 'gpu_launches': launches, 'fp32_fft_mode': fast, 'crnn': crnn, 'other_configs': other, 'train_step': train_leg,
        'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _main_with_clean_stdout():
    """Libraries (NCCL's version banner, for one) write to file descriptor 1; the contract is ONE JSON line on stdout.
    Everything is sent to stderr while the benchmark runs and only the result line goes to the real stdout."""
    import io
    real_stdout = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    buf = io.StringIO()
    py_stdout, sys.stdout = sys.stdout, buf
    try:
        rc = main()
    finally:
        sys.stdout = py_stdout
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    out = buf.getvalue()
    json_lines = [ln for ln in out.splitlines() if ln.startswith('{')]
    for ln in out.splitlines():
        if not ln.startswith('{'):
            print(ln, file=sys.stderr)
    if json_lines:
        print(json_lines[-1])
        sys.stdout.flush()
    return rc


if __name__ == '__main__':
    sys.exit(_main_with_clean_stdout())
