"""On-the-fly training step of the SELD CRNN (BASELINE.json configs[4]).

The reference trains through PyTorch Lightning: `training_step` (models/seld_models.py:51-76) = forward in train mode ->
`interpolate_tensor` to the label rate -> `compute_loss` (models/interfaces.py:273-355), Adam with the piecewise-linear
lr / beta1 schedule (utilities/learning_utils.py:17-52), DDP gradient all-reduce (experiments/train.py:98-104).  Here:

  native (libsalsa_b200.so)   SALSA features on the fly, augmentations, every convolution's forward (3x3, 1x1 and the 7-channel
                              first one), input gradient
                              (the tcgen05 implicit-GEMM kernel on flipped / transposed weights) and weight gradient
                              (`crnn_conv_wgrad`), the loss with its output gradients (`crnn_seld_loss`), the Adam step,
                              train-mode BatchNorm fused with the residual add, the ReLU and the encoder's dropout, forward
                              and backward (`crnn_bn_train_forward` / `_backward`), 2x2 average pooling forward and backward
  torch (library)             the heads (four pairs of nn.Linear with their dropouts) and the GEMMs around the GRU recurrence,
                              through autograd / cuBLAS -- said so wherever a number is quoted
  torch.distributed           bf16 all-reduce of the flat gradient buffer (`GradAllReduce`; NCCL on the GPU box, gloo in the
                              CPU tests): per bucket while the backward pass is still running (eager step), or one call after
                              the replay when forward + loss + backward run as ONE CUDA graph (`use_graph=True`)

Parameters live in ONE flat float32 buffer (the optimiser's view) with per-tensor views carrying the reference's state-dict
names, so `state_dict()` interchanges with reference checkpoints and with the inference model (`SeldModel.load_state_dict`).
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import crnn_ops as ops
from .optim import Adam, LearningRateScheduler

__all__ = ['SeldTrainer', 'GradAllReduce', 'NativeConv3x3', 'NativeConv1x1', 'NativeConvFirst', 'NativeBnAct', 'NativeBnActPool', 'NativeAvgPool2', 'NativeGRULayer']


class NativeConv3x3(torch.autograd.Function):
    """3x3 / pad 1 / stride 1 convolution without bias on channels_last bf16 tensors.  forward: `crnn_conv2d`.  backward:
    input gradient = `crnn_conv2d` of the output gradient with the taps flipped and Cin / Cout exchanged; weight gradient =
    `crnn_conv_wgrad` (tcgen05 GEMMs over the pixel axis on MN-major operands).  `native_wgrad = False` switches the weight
    gradient to torch.nn.grad.conv2d_weight (cuDNN) for A/B runs."""
    native_wgrad = True

    @staticmethod
    def forward(ctx, x, w):
        xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wp = w.detach().permute(2, 3, 0, 1).reshape(9, w.shape[0], w.shape[1]).to(torch.bfloat16).contiguous()
        out = ops.conv2d(xb.permute(0, 2, 3, 1), wp)                       # NHWC view of the channels_last tensor
        ctx.save_for_backward(xb, w)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xb, w = ctx.saved_tensors
        gyb = gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wt = w.detach().flip(2, 3).permute(2, 3, 1, 0).reshape(9, w.shape[1], w.shape[0]).to(torch.bfloat16).contiguous()
            dx = ops.conv2d(gyb.permute(0, 2, 3, 1), wt).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            if NativeConv3x3.native_wgrad:
                dw = ops.conv_wgrad(xb.permute(0, 2, 3, 1), gyb.permute(0, 2, 3, 1))          # (9, Cout, Cin) fp32
                dw = dw.reshape(3, 3, w.shape[0], w.shape[1]).permute(2, 3, 0, 1).to(w.dtype)
            else:
                dw = torch.nn.grad.conv2d_weight(xb, w.shape, gyb, padding=1).to(w.dtype)
        return dx, dw


class NativeConvFirst(torch.autograd.Function):
    """The first convolution (7 -> 64 channels, 3x3 / pad 1, no bias; models/model_utils.py:192-195 with in_channels = 7) on
    the float32 NCHW feature batch: `crnn_pack_input` (NHWC bf16, channels padded to 16) -> `crnn_conv_first`; the weight
    gradient is `crnn_conv_wgrad` on the 16-channel tensor (its TMA boxes span 64 channels and arrive zero-filled above
    channel 15).  The input needs no gradient."""

    @staticmethod
    def forward(ctx, x, w):
        x16 = ops.pack_input(x.detach().float(), c_pad=16)                           # (B, T, F, 16) bf16
        wp = torch.zeros((9, w.shape[0], 16), dtype=torch.bfloat16, device=w.device)
        wp[:, :, :w.shape[1]] = w.detach().permute(2, 3, 0, 1).reshape(9, w.shape[0], w.shape[1])
        ctx.save_for_backward(x16)
        ctx.wshape = tuple(w.shape)
        return ops.conv_first(x16, wp, bias=None, relu=False).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        x16, = ctx.saved_tensors
        cout, cin = ctx.wshape[:2]
        gyb = gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dw = ops.conv_wgrad(x16, gyb.permute(0, 2, 3, 1))                             # (9, 64, 16) fp32
        return None, dw[:, :, :cin].reshape(3, 3, cout, cin).permute(2, 3, 0, 1)


class NativeConv1x1(torch.autograd.Function):
    """1x1 convolution without bias (the downsample branch, models/model_utils.py:307-309) on channels_last bf16 tensors:
    forward and input gradient = `crnn_conv2d` with one tap (weights transposed for the gradient), weight gradient =
    `crnn_conv_wgrad(ksize=1)`."""

    @staticmethod
    def forward(ctx, x, w):
        xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wp = w.detach().reshape(1, w.shape[0], w.shape[1]).to(torch.bfloat16).contiguous()
        ctx.save_for_backward(xb, w)
        return ops.conv2d(xb.permute(0, 2, 3, 1), wp).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xb, w = ctx.saved_tensors
        gyb = gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wt = w.detach().reshape(w.shape[0], w.shape[1]).t().reshape(1, w.shape[1], w.shape[0]).to(torch.bfloat16).contiguous()
            dx = ops.conv2d(gyb.permute(0, 2, 3, 1), wt).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            dw = ops.conv_wgrad(xb.permute(0, 2, 3, 1), gyb.permute(0, 2, 3, 1), ksize=1).reshape(w.shape).to(w.dtype)
        return dx, dw


class NativeGRULayer(torch.autograd.Function):
    """One bidirectional nn.GRU layer (hidden size 256, batch_first) in float32: the recurrence and back-propagation through
    time are `crnn_gru_layer_train` / `crnn_gru_layer_backward` (cluster kernels, W_hh resident in shared memory); the input
    projection and the parameter / input gradients are plain float32 GEMMs and column sums around them (library GEMMs).
    Arguments after x: weight_ih, weight_hh, bias_ih, bias_hh of the forward direction, then of the reverse direction.
    `gemm_dtype`: operand type of those GEMMs (accumulation is float32 either way): bf16 like the rest of the autocast step
    by default, float32 for exactness tests."""
    gemm_dtype = torch.bfloat16

    @staticmethod
    def _mm(a, b):
        dt = NativeGRULayer.gemm_dtype
        return a @ b if dt == torch.float32 else (a.to(dt) @ b.to(dt)).float()

    @staticmethod
    @torch.amp.custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, x, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r):
        B, T, n_in = x.shape
        xf = x.contiguous()
        w_ih = torch.cat([w_ih_f, w_ih_r]).contiguous()                       # (1536, In)
        w_hh = torch.stack([w_hh_f, w_hh_r]).contiguous()                     # (2, 768, 256)
        b_hh = torch.stack([b_hh_f, b_hh_r]).contiguous()
        xproj = (NativeGRULayer._mm(xf.reshape(-1, n_in), w_ih.t()) + torch.cat([b_ih_f, b_ih_r])).view(B, T, 1536)
        y, save = ops.gru_layer_train(xproj, w_hh, b_hh)
        ctx.save_for_backward(xf, w_ih, w_hh, y, save)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type='cuda')
    def backward(ctx, dy):
        xf, w_ih, w_hh, y, save = ctx.saved_tensors
        B, T, n_in = xf.shape
        dgi, dgh = ops.gru_layer_backward(dy.float().contiguous(), y, save, w_hh)
        gi, gh = dgi.view(-1, 1536), dgh.view(-1, 1536)
        mm = NativeGRULayer._mm
        dx = mm(gi, w_ih).view(B, T, n_in)
        dw_ih, db_ih, db_hh = mm(gi.t(), xf.reshape(-1, n_in)), gi.sum(0), gh.sum(0)
        zero = y.new_zeros(B, 1, 256)
        hp_f = torch.cat([zero, y[:, :-1, :256]], 1).reshape(-1, 256)          # h_{t-1} of the forward direction
        hp_r = torch.cat([y[:, 1:, 256:], zero], 1).reshape(-1, 256)           # h_{t+1} of the reverse direction
        return (dx, dw_ih[:768], mm(gh[:, :768].t(), hp_f), db_ih[:768], db_hh[:768],
                dw_ih[768:], mm(gh[:, 768:].t(), hp_r), db_ih[768:], db_hh[768:])


class NativeBnAct(torch.autograd.Function):
    """Train-mode BatchNorm2d (+ residual add) (+ ReLU) on channels_last bf16 tensors: `crnn_bn_train_forward` /
    `crnn_bn_train_backward`.  Running statistics are updated in place in the forward pass."""

    @staticmethod
    def forward(ctx, y, gamma, beta, residual, running_mean, running_var, relu, drop=None):
        """drop = (device seed tensor, salt, p): element-wise dropout behind the ReLU (nn.Dropout after relu(bn1(.)),
        models/model_utils.py:356), recomputed from the seed in the backward pass instead of stored."""
        yb = y.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        rb = None if residual is None else residual.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        g, b = gamma.detach().contiguous(), beta.detach().contiguous()
        z, stat = ops.bn_train_forward(yb.permute(0, 2, 3, 1), g, b, None if rb is None else rb.permute(0, 2, 3, 1), relu=relu,
                                       running_mean=running_mean, running_var=running_var, drop=drop)
        ctx.relu, ctx.has_res, ctx.drop = relu, residual is not None, drop
        # without a residual the ReLU mask is a function of y: the backward pass recomputes it instead of reading z
        ctx.save_for_backward(yb, z if (relu and ctx.has_res) else None, stat, g, b)
        return z.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dz):
        yb, z, stat, g, b = ctx.saved_tensors
        dzb = dz.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dy, dres, dgamma, dbeta = ops.bn_train_backward(dzb.permute(0, 2, 3, 1), z, yb.permute(0, 2, 3, 1), stat, g, relu=ctx.relu, beta=b,
                                                        want_residual_grad=ctx.has_res and ctx.needs_input_grad[3], drop=ctx.drop)
        return (dy.permute(0, 3, 1, 2), dgamma, dbeta, None if dres is None else dres.permute(0, 3, 1, 2), None, None, None, None)


class NativeBnActPool(torch.autograd.Function):
    """relu(BatchNorm2d(y) (+ residual)) followed by F.avg_pool2d(2) in one pass each way (`crnn_bn_train_forward_pool` /
    `_backward_pool`): used where the network pools and the BatchNorm output has no other consumer, so the full-resolution
    activation is never written and the backward reads the pooled gradient directly.  Bit-identical to NativeBnAct followed
    by NativeAvgPool2."""

    @staticmethod
    def forward(ctx, y, gamma, beta, residual, running_mean, running_var):
        yb = y.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        rb = None if residual is None else residual.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        g, b = gamma.detach().contiguous(), beta.detach().contiguous()
        pooled, stat = ops.bn_train_forward_pool(yb.permute(0, 2, 3, 1), g, b, None if rb is None else rb.permute(0, 2, 3, 1),
                                                 running_mean=running_mean, running_var=running_var)
        ctx.save_for_backward(yb, rb, stat, g, b)
        return pooled.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dpool):
        yb, rb, stat, g, b = ctx.saved_tensors
        dpb = dpool.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dy, dres, dgamma, dbeta = ops.bn_train_backward_pool(dpb.permute(0, 2, 3, 1), yb.permute(0, 2, 3, 1),
                                                             None if rb is None else rb.permute(0, 2, 3, 1), stat, g, b,
                                                             want_residual_grad=rb is not None and ctx.needs_input_grad[3])
        return dy.permute(0, 3, 1, 2), dgamma, dbeta, None if dres is None else dres.permute(0, 3, 1, 2), None, None


class NativeAvgPool2(torch.autograd.Function):
    """F.avg_pool2d(x, 2) on channels_last bf16 tensors: `crnn_avgpool2` / `crnn_avgpool2_backward` (torch's NHWC pooling
    kernels took a third of the training step)."""

    @staticmethod
    def forward(ctx, x):
        xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ctx.hw = (xb.shape[2], xb.shape[3])
        return ops.avgpool2(xb.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        dyb = dy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        return ops.avgpool2_backward(dyb.permute(0, 2, 3, 1), *ctx.hw).permute(0, 3, 1, 2)


class GradAllReduce:
    """Averages a flat gradient buffer over the process group in buckets, each bucket's all-reduce started (asynchronously,
    bf16 on the wire by default) as soon as every parameter inside it has its gradient, i.e. while the backward pass of the
    earlier layers is still running.  `params` are the per-tensor views in FORWARD order; the backward pass fills them back
    to front, so buckets are cut from the end."""

    def __init__(self, flat_grad: torch.Tensor, offsets, bucket_bytes: int = 8 << 20, wire_dtype=torch.bfloat16, group=None):
        self.flat, self.group, self.wire = flat_grad, group, wire_dtype
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        # buckets = runs of whole parameters, cut from the last parameter backwards
        self.buckets = []                      # (lo, hi) element ranges
        self.bucket_of = {}                    # parameter index -> bucket index
        hi = offsets[-1][1] if offsets else 0
        cur = []
        for i in range(len(offsets) - 1, -1, -1):
            cur.append(i)
            lo = offsets[i][0]
            if (hi - lo) * 4 >= bucket_bytes or i == 0:
                for j in cur:
                    self.bucket_of[j] = len(self.buckets)
                self.buckets.append((lo, hi))
                hi, cur = lo, []
        self.pending = [0] * len(self.buckets)
        self.sizes = [sum(1 for j in self.bucket_of.values() if j == b) for b in range(len(self.buckets))]
        self.work = []
        self.enabled = True            # False: the hooks do nothing (a CUDA graph capture must not contain the collective)

    def reset(self):
        self.pending = list(self.sizes)
        self.work = []

    def ready(self, param_index: int):
        """Call when parameter `param_index` has its final gradient (a post-accumulate-grad hook)."""
        if not self.enabled:
            return
        b = self.bucket_of[param_index]
        self.pending[b] -= 1
        if self.pending[b] == 0 and self.world > 1:
            lo, hi = self.buckets[b]
            chunk = self.flat[lo:hi]
            wire = chunk if self.wire is None or chunk.dtype == self.wire else chunk.to(self.wire)
            self.work.append((dist.all_reduce(wire, group=self.group, async_op=True), wire, chunk))

    def finish(self):
        """Waits for the buckets in flight and writes the averages back into the flat buffer."""
        if not self.enabled:
            return
        for b, n in enumerate(self.pending):          # parameters that received no gradient this step
            if n > 0 and self.world > 1:
                self.pending[b] = 1
                self.ready(next(j for j, bb in self.bucket_of.items() if bb == b))
        for work, wire, chunk in self.work:
            work.wait()
            if wire is not chunk:
                chunk.copy_(wire)
            chunk.div_(self.world)
        self.work = []


    def reduce_all(self):
        """The whole flat gradient in ONE all-reduce (28 MB in bf16: a fraction of a millisecond over NVLink), for a step whose
        backward pass is replayed as a CUDA graph and therefore cannot start bucket collectives from hooks."""
        if self.world == 1:
            return
        wire = self.flat if self.wire is None or self.flat.dtype == self.wire else self.flat.to(self.wire)
        dist.all_reduce(wire, group=self.group)
        if wire is not self.flat:
            self.flat.copy_(wire)
        self.flat.div_(self.world)


class SeldTrainer:
    """PannResNet22 + SeldDecoder(bigru, avg) in train mode with the reference's state-dict names, one flat parameter buffer,
    and `step(x, target_dict)` = the reference's training step."""

    def __init__(self, state_dict, n_classes: int = 12, label_rate: int = 10, feature_rate: float = 80.0, loss_weight=(0.3, 0.7),
                 lr: float = 1e-3, device='cuda', native_conv: bool = True, group=None, scheduler: LearningRateScheduler = None,
                 bucket_bytes: int = 8 << 20, wire_dtype=torch.bfloat16, autocast: bool = True, dropout: bool = True,
                 native_bn: bool = True, use_graph: bool = False, native_gru: bool = True):
        """native_conv / native_bn / autocast / dropout = False are for tests (a pure torch float32 reference of the same step);
        wire_dtype None sends float32 gradients.  use_graph: `step` captures forward + loss + backward as ONE CUDA graph at the
        first sighting of a batch shape and replays it afterwards (about 1000 launches per step otherwise: the step is
        host-bound without it); the gradient all-reduce (one NCCL call on the flat buffer) and Adam follow the replay as
        ordinary launches -- a collective inside a capture is not portable across NCCL versions."""
        self.device = torch.device(device)
        self.n_classes, self.loss_weight = n_classes, tuple(loss_weight)
        self.ratio = 16.0 * label_rate / feature_rate                 # time_downsample_ratio * label_rate / feature_rate
        self.native_conv = native_conv and self.device.type == 'cuda'
        self.native_bn = native_bn and self.device.type == 'cuda'
        self.native_gru = native_gru and self.device.type == 'cuda'
        self.fuse_pool = True                # BatchNorm + ReLU + pooling as one pass where the network pools (False: two passes)
        self.autocast = autocast and self.device.type == 'cuda'
        self.dropout = dropout
        self.scheduler = scheduler
        names = [k for k, v in state_dict.items() if not (k.endswith('running_mean') or k.endswith('running_var') or k.endswith('num_batches_tracked'))]
        sizes = [int(torch.as_tensor(state_dict[k]).numel()) for k in names]
        total = sum(sizes)
        self.flat = torch.empty(total, dtype=torch.float32, device=self.device)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.params, self.offsets, at = {}, [], 0
        for k, n in zip(names, sizes):
            src = torch.as_tensor(state_dict[k]).detach().to(self.device, torch.float32)
            self.flat[at:at + n].copy_(src.reshape(-1))
            p = self.flat[at:at + n].view(src.shape).requires_grad_(True)
            p.grad = self.flat_grad[at:at + n].view(src.shape)        # autograd accumulates in place into the flat gradient
            self.params[k] = p
            self.offsets.append((at, at + n))
            at += n
        self.buffers = {k: torch.as_tensor(v).detach().clone().to(self.device) for k, v in state_dict.items() if k not in self.params}
        self.reducer = GradAllReduce(self.flat_grad, self.offsets, bucket_bytes=bucket_bytes, wire_dtype=wire_dtype, group=group)
        for i, p in enumerate(self.params.values()):
            p.register_post_accumulate_grad_hook(lambda _p, i=i: self.reducer.ready(i))
        self.optimizer = Adam(self.flat, lr=lr) if self.device.type == 'cuda' else None
        self.gru = torch.nn.GRU(input_size=512, hidden_size=256, num_layers=2, batch_first=True, bidirectional=True, dropout=0.3).to(self.device)
        self.training = True
        self.epoch, self.batch_idx = 0, 0
        self.use_graph = use_graph and self.device.type == 'cuda'
        # seed of the dropout fused into the BatchNorm kernels: a device counter, advanced by a launch inside the step so that a
        # replayed CUDA graph draws new masks every time
        self.drop_seed = torch.full((1,), 0x5A15A, dtype=torch.int64, device=self.device) if self.device.type == 'cuda' else None
        self._graphs = {}                    # batch shape -> (graph, static inputs, static loss)
        self._index_cache = {}
        self.graph_error = None              # why a capture fell back to eager launches (None: it did not)

    # ---- state ------------------------------------------------------------------------------------------------------
    def state_dict(self):
        out = {k: v.detach().clone() for k, v in self.params.items()}
        out.update({k: v.clone() for k, v in self.buffers.items()})
        return out

    # ---- forward (train mode) ---------------------------------------------------------------------------------------
    def _conv3(self, x, key):
        w = self.params[key]
        if self.native_conv and w.shape[1] % 64 == 0:
            return NativeConv3x3.apply(x, w)
        return F.conv2d(x, w, padding=1)

    def _pool(self, x):
        if self.native_bn and x.is_cuda and x.shape[1] % 8 == 0:
            return NativeAvgPool2.apply(x)
        return F.avg_pool2d(x, 2)

    def _bn(self, x, prefix, relu=False, residual=None, drop_p=0.0, salt=0, pool=False):
        """BatchNorm2d (+ residual) (+ ReLU) (+ dropout) (+ the 2x2 average pooling behind it): one native pass each way in
        train mode, torch ops otherwise."""
        rm, rv = self.buffers[prefix + '.running_mean'], self.buffers[prefix + '.running_var']
        w, b = self.params[prefix + '.weight'], self.params[prefix + '.bias']
        if self.native_bn and self.training and x.is_cuda and x.shape[1] in (64, 128, 256, 512):
            if pool and relu and drop_p == 0.0 and self.fuse_pool and x.shape[2] >= 2 and x.shape[3] >= 2:
                return NativeBnActPool.apply(x, w, b, residual, rm, rv)
            drop = (self.drop_seed, salt, drop_p) if drop_p > 0.0 else None
            out = NativeBnAct.apply(x, w, b, residual, rm, rv, relu, drop)
            return self._pool(out) if pool else out
        out = F.batch_norm(x, rm, rv, w, b, training=self.training, momentum=0.1, eps=1e-5)
        if residual is not None:
            out = out + residual
        out = F.relu(out) if relu else out
        out = F.dropout(out, p=drop_p, training=True) if drop_p > 0.0 else out
        return self._pool(out) if pool else out

    def forward(self, x):
        """x (B, 7, T, F) float32 -> {'event_frame_logit': (B, T/16, n), 'doa_frame_output': (B, T/16, 3n)}, with autograd."""
        tr = self.training and self.dropout
        with torch.autocast(self.device.type, dtype=torch.bfloat16, enabled=self.autocast):
            p = 'encoder.conv_block1'
            w0 = self.params[p + '.conv1.weight']
            if self.native_conv and x.is_cuda and w0.shape[0] == 64 and w0.shape[1] <= 16:
                y0 = NativeConvFirst.apply(x, w0)
            else:
                y0 = F.conv2d(x.contiguous(memory_format=torch.channels_last), w0, padding=1)
            x = self._bn(y0, p + '.bn1', relu=True)
            # ConvBlock.forward (models/model_utils.py:213-220): bn2 + ReLU + the pooling in one pass
            x = self._bn(self._conv3(x, p + '.conv2.weight'), p + '.bn2', relu=True, pool=True)
            for li in range(1, 5):
                for bi in range(2):
                    q = 'encoder.resnet.layer{}.{}'.format(li, bi)
                    identity = x
                    # _ResnetBasicBlock.forward (:345-367) pools at the entry of layers 2-4; that pooling is the only consumer of
                    # the previous block's output, so it already happened in that block's last BatchNorm pass (below)
                    pooled = x
                    # relu(bn1(conv1(.))) and the dropout behind it (:354-356) in one pass
                    out = self._bn(self._conv3(pooled, q + '.conv1.weight'), q + '.bn1', relu=True, drop_p=0.1 if tr else 0.0, salt=2 * li + bi)
                    if li > 1 and bi == 0:                                     # downsample = AvgPool2d(2) + 1x1 conv + BN (:474-481): the same pooled tensor
                        wd = self.params[q + '.downsample.1.weight']
                        ds = NativeConv1x1.apply(pooled, wd) if self.native_conv else F.conv2d(pooled, wd)
                        identity = self._bn(ds, q + '.downsample.2')
                    x = self._bn(self._conv3(out, q + '.conv2.weight'), q + '.bn2', relu=True, residual=identity,   # relu(bn2(.) + identity)
                                 pool=(bi == 1 and li < 4))
            x = torch.mean(x.float(), dim=3).transpose(1, 2)                  # SeldDecoder.forward (models/decoders.py:106-154)
            if self.native_gru and x.is_cuda:
                for layer in range(2):
                    keys = ['decoder.gru.{}_l{}{}'.format(n, layer, suf) for suf in ('', '_reverse')
                            for n in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')]
                    x = NativeGRULayer.apply(x, *[self.params[k] for k in keys])
                    if layer == 0:
                        x = F.dropout(x, 0.3, tr)                             # inter-layer dropout (models/decoders.py:44-46)
            else:
                gru_params = {k[len('decoder.gru.'):]: v for k, v in self.params.items() if k.startswith('decoder.gru.')}
                self.gru.train(self.training)
                self.gru.dropout = 0.3 if tr else 0.0
                x, _ = torch.func.functional_call(self.gru, gru_params, (x,))

            def head(name, act=None):
                h = F.relu(F.linear(F.dropout(x, 0.2, tr), self.params['decoder.{}_fc_1.weight'.format(name)], self.params['decoder.{}_fc_1.bias'.format(name)]))
                h = F.linear(F.dropout(h, 0.2, tr), self.params['decoder.{}_fc_2.weight'.format(name)], self.params['decoder.{}_fc_2.bias'.format(name)])
                return h if act is None else act(h)

            logit = head('event')
            doa = torch.cat([head('x', torch.tanh), head('y', torch.tanh), head('z', torch.tanh)], dim=-1)
        return {'event_frame_logit': logit.float(), 'doa_frame_output': doa.float()}

    # ---- one training step ------------------------------------------------------------------------------------------
    def _label_index(self, n_frames):
        """Index map of interpolate_tensor on the device, built once per sequence length (no host copy inside a capture)."""
        key = (n_frames, self.ratio)
        if key not in self._index_cache:
            self._index_cache[key] = torch.as_tensor(ops.interpolate_index(n_frames, self.ratio), device=self.device)
        return self._index_cache[key]

    def step(self, x, target_dict):
        """forward -> interpolate to the label rate -> loss -> backward (bucketed all-reduce overlapped) -> Adam.
        Returns (loss, sed_loss, doa_loss) as a float32 tensor (3,)."""
        if self.use_graph and self.graph_error is None:
            return self._step_graph(x, target_dict)
        return self._step_eager(x, target_dict)

    # ---- the step as one CUDA graph ---------------------------------------------------------------------------------------
    def _mutable_state(self):
        return [self.flat, self.optimizer.exp_avg, self.optimizer.exp_avg_sq] + list(self.buffers.values())

    def _step_graph(self, x, target_dict):
        egt, dgt = target_dict['event_frame_gt'], target_dict['doa_frame_gt']
        key = (tuple(x.shape), tuple(egt.shape), tuple(dgt.shape))
        if key not in self._graphs:
            static = [torch.empty_like(x, memory_format=torch.contiguous_format), torch.empty_like(egt), torch.empty_like(dgt)]
            for s_, t_ in zip(static, (x, egt, dgt)):
                s_.copy_(t_)
            body = lambda: self._step_eager(static[0], {'event_frame_gt': static[1], 'doa_frame_gt': static[2]}, gradients_only=True)
            # warm-up on a side stream (lazy initialisation of cuDNN / autograd / the allocator must not be captured), with
            # the trainer's state (BatchNorm running statistics) put back afterwards so that it is not part of the trajectory
            saved = [t.clone() for t in self._mutable_state()]
            self.reducer.enabled = False
            try:
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    for _ in range(2):
                        body()
                torch.cuda.current_stream(self.device).wait_stream(side)
                torch.cuda.synchronize(self.device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    loss = body()
            except Exception as e:                      # capture refused (a library call that is not capturable): eager launches
                self.graph_error = '{}: {}'.format(type(e).__name__, str(e).splitlines()[0] if str(e) else '')
                torch.cuda.synchronize(self.device)
                graph = None
            self.reducer.enabled = True
            for t, sv in zip(self._mutable_state(), saved):
                t.copy_(sv)
            if graph is None:
                return self._step_eager(x, target_dict)
            self._graphs[key] = (graph, static, loss)
        graph, static, loss = self._graphs[key]
        for s_, t_ in zip(static, (x, egt, dgt)):
            s_.copy_(t_)
        graph.replay()                                   # zero gradients, forward, loss, backward
        self.reducer.reduce_all()
        if self.scheduler is not None:
            self.scheduler.apply(self.optimizer, self.epoch, self.batch_idx)
        self.optimizer.step(self.flat_grad)
        self.batch_idx += 1
        return loss.clone()

    def _step_eager(self, x, target_dict, gradients_only=False):
        """The launches of one step; `gradients_only`: stop after the backward pass (the body a CUDA graph is captured from;
        all-reduce, schedule and Adam are the caller's)."""
        self.flat_grad.zero_()
        self.reducer.reset()
        if self.drop_seed is not None:
            self.drop_seed.add_(1)                          # a launch: part of the captured graph, so every replay advances it
        out = self.forward(x)
        idx = self._label_index(out['event_frame_logit'].shape[1])
        logit, doa = out['event_frame_logit'][:, idx], out['doa_frame_output'][:, idx]      # interpolate_tensor (model_utils.py:57-75)
        n = min(logit.shape[1], target_dict['event_frame_gt'].shape[1])
        logit, doa = logit[:, :n], doa[:, :n]
        egt, dgt = target_dict['event_frame_gt'][:, :n], target_dict['doa_frame_gt'][:, :n]
        if self.device.type == 'cuda':
            loss, g_logit, g_doa = ops.seld_loss(logit.detach(), doa.detach(), egt, dgt, loss_weight=self.loss_weight, with_grad=True)
            torch.autograd.backward([logit, doa], [g_logit, g_doa])
        else:                                                                 # CPU tests of the host logic: the same loss in torch
            sed = F.binary_cross_entropy_with_logits(logit, egt)
            d = sum(((doa[..., i * self.n_classes:(i + 1) * self.n_classes] - dgt[..., i * self.n_classes:(i + 1) * self.n_classes]).abs() * egt).sum()
                    / egt.sum().clamp(min=1e-12) for i in range(3))
            total = self.loss_weight[0] * sed + self.loss_weight[1] * d
            total.backward()
            loss = torch.stack([total.detach(), sed.detach(), d.detach()])
        if gradients_only:
            return loss
        self.reducer.finish()
        if self.scheduler is not None and self.optimizer is not None:
            self.scheduler.apply(self.optimizer, self.epoch, self.batch_idx)
        if self.optimizer is not None:
            self.optimizer.step(self.flat_grad)
        self.batch_idx += 1
        return loss
