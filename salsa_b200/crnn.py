"""Host-side mirror of the reference SELD CRNN (forward / inference path) on libsalsa_b200.so.

Same class names, constructor arguments, attributes and state-dict keys as the reference:
`models.encoders.PannResNet22` (encoders.py:26-56), `models.decoders.SeldDecoder` (decoders.py:13-154,
`decoder_type='bigru'`, `freq_pool='avg'` -- the only combination the reference configs use,
experiments/configs/seld.yml:30-32) and `models.seld_models.SeldModel.forward` (seld_models.py:39-49).
`SeldModel.load_state_dict` takes a reference checkpoint's `['state_dict']` (inference.py:115-116).

Eval mode only: BatchNorm uses running statistics (folded into the convolution weights when the state
dict is loaded) and dropout is the identity.  All compute happens in the CUDA library; there is no
PyTorch / CPU fallback.  Arithmetic: bf16 operands, fp32 accumulation (tensor cores), fp32 GRU state.
"""
import ctypes

import numpy as np
import torch

from . import _native
from . import crnn_ops as ops

BN_EPS = 1e-5


def _t(v):
    return v.detach().float().cpu() if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v)).float()


def random_state_dict(seed: int = 0) -> dict:
    """Random-init weights with the reference's state-dict keys and shapes (for benchmarks and smoke runs:
    there are no checkpoints to download).  He-normal convolutions / linears, non-trivial BatchNorm
    statistics, uniform(-1/16, 1/16) GRU weights (PyTorch's default for hidden size 256)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv_bn(wkey, bn, cout, cin, k):
        sd[wkey] = torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5
        sd[bn + '.weight'] = torch.empty(cout).uniform_(0.5, 1.5, generator=g)
        sd[bn + '.bias'] = torch.empty(cout).uniform_(-0.2, 0.2, generator=g)
        sd[bn + '.running_mean'] = torch.empty(cout).uniform_(-0.2, 0.2, generator=g)
        sd[bn + '.running_var'] = torch.empty(cout).uniform_(0.5, 1.5, generator=g)
        sd[bn + '.num_batches_tracked'] = torch.zeros((), dtype=torch.int64)

    conv_bn('encoder.conv_block1.conv1.weight', 'encoder.conv_block1.bn1', 64, 7, 3)
    conv_bn('encoder.conv_block1.conv2.weight', 'encoder.conv_block1.bn2', 64, 64, 3)
    inpl = 64
    for li, planes in zip(range(1, 5), (64, 128, 256, 512)):
        for bi in range(2):
            p = 'encoder.resnet.layer{}.{}'.format(li, bi)
            conv_bn(p + '.conv1.weight', p + '.bn1', planes, inpl if bi == 0 else planes, 3)
            conv_bn(p + '.conv2.weight', p + '.bn2', planes, planes, 3)
            if li > 1 and bi == 0:
                conv_bn(p + '.downsample.1.weight', p + '.downsample.2', planes, inpl, 1)
        inpl = planes
    for layer in range(2):
        for suffix in ('', '_reverse'):
            for name, shape in (('weight_ih', (768, 512)), ('weight_hh', (768, 256)), ('bias_ih', (768,)), ('bias_hh', (768,))):
                sd['decoder.gru.{}_l{}{}'.format(name, layer, suffix)] = torch.empty(shape).uniform_(-1 / 16, 1 / 16, generator=g)
    for head in ('event', 'x', 'y', 'z'):
        sd['decoder.{}_fc_1.weight'.format(head)] = torch.randn((256, 512), generator=g) * (2.0 / 512) ** 0.5
        sd['decoder.{}_fc_1.bias'.format(head)] = torch.zeros(256)
        sd['decoder.{}_fc_2.weight'.format(head)] = torch.randn((12, 256), generator=g) * (2.0 / 256) ** 0.5
        sd['decoder.{}_fc_2.bias'.format(head)] = torch.zeros(12)
    return sd


class PannResNet22:
    """Encoder description (encoders.py:26-46).  Holds no parameters itself: `SeldModel` owns the packed weights."""

    def __init__(self, n_input_channels: int = 1, p_dropout: float = 0.0, **kwargs):
        self.n_input_channels = n_input_channels
        self.p_dropout = p_dropout
        self.n_output_channels = 512
        self.time_downsample_ratio = 16


class SeldDecoder:
    """Decoder description (decoders.py:18-46)."""

    def __init__(self, n_output_channels, n_classes: int = 12, output_format: str = 'reg_xyz', decoder_type: str = None,
                 freq_pool: str = None, decoder_size: int = 128, **kwargs):
        assert decoder_type in ['gru', 'bigru', 'lstm', 'bilstm', 'transformer'], 'Invalid decoder type {}'.format(decoder_type)
        if decoder_type != 'bigru':
            raise NotImplementedError('decoder type: {} is not implemented (salsa_b200 implements bigru)'.format(decoder_type))
        if freq_pool != 'avg':
            raise NotImplementedError('freq pooling {} is not implemented (salsa_b200 implements avg)'.format(freq_pool))
        if decoder_size != 256 or n_output_channels != 512:
            raise NotImplementedError('salsa_b200 implements decoder_size=256 on 512 encoder channels')
        if 4 * n_classes > 64:
            raise NotImplementedError('at most 16 classes')
        self.n_classes = n_classes
        self.decoder_type = decoder_type
        self.freq_pool = freq_pool
        self.doa_format = output_format
        self.gru_size = decoder_size
        self.fc_size = 2 * decoder_size


class SeldModel:
    """Counterpart of models.seld_models.SeldModel: the forward / inference path on the native kernels, and -- after
    train() -- the training step through salsa_b200.train.SeldTrainer."""

    def __init__(self, encoder: PannResNet22, decoder: SeldDecoder, label_rate: int = 10, feature_rate: float = None,
                 device='cuda', precision: str = 'bf16', **kwargs):
        """precision: 'bf16' (bf16 operands, fp32 accumulation: the fast mode) or 'bf16x3' (every operand is the
        sum of three bf16 planes, six plane products per MAC on the same tensor-core kernels: float32-grade
        results, used for parity with the float32 reference)."""
        if precision not in ('bf16', 'bf16x2', 'bf16x3'):
            raise ValueError('precision must be bf16, bf16x2 or bf16x3')
        self.precision = precision
        self.planes = {'bf16': 1, 'bf16x2': 2, 'bf16x3': 3}[precision]
        self.encoder, self.decoder = encoder, decoder
        self.label_rate, self.feature_rate = label_rate, feature_rate
        self.time_downsample_ratio = float(encoder.time_downsample_ratio)
        self.n_classes = decoder.n_classes
        self.doa_format = decoder.doa_format
        self.device = torch.device(device)
        self.training = False
        self._w = None
        self._scaler = None
        self._native_model = None          # crnn_load_weights handle (the one-call forward)
        self._workspace = None
        self._out = None
        self._sd = None                    # the loaded state dict (float32, CPU): what a trainer starts from
        self._trainer = None               # salsa_b200.train.SeldTrainer, created by train()
        self._trainer_kwargs = {'loss_weight': tuple(kwargs.get('loss_weight', (0.3, 0.7))), 'lr': float(kwargs.get('lr', 1e-3))}

    def __del__(self):
        self._release_native()

    def _release_native(self):
        if getattr(self, '_native_model', None):
            try:
                _native.lib().crnn_free_model(self._native_model)
            except Exception:              # noqa: BLE001 -- interpreter shutdown
                pass
            self._native_model = None

    # ---- nn.Module-like surface -----------------------------------------------------------------
    def eval(self):
        """Back to the inference path; weights trained since train() are folded into it first."""
        if self.training and self._trainer is not None:
            self.load_state_dict(self._trainer.state_dict())
        self.training = False
        return self

    def train(self, mode: bool = True, **trainer_kwargs):
        """nn.Module.train(): the training side of the reference's LightningModule (models/seld_models.py:51-76,
        models/interfaces.py:85-95) is `salsa_b200.train.SeldTrainer`; train() creates one from the loaded weights (keyword
        arguments go to it: scheduler, use_graph, group, ...), `training_step` / `validation_step`-style calls go through it,
        and eval() hands the trained weights back to the inference kernels."""
        if not mode:
            return self.eval()
        if self._sd is None:
            raise RuntimeError('load_state_dict() first')
        if self._trainer is None or trainer_kwargs:
            from .train import SeldTrainer
            kw = dict(self._trainer_kwargs)
            kw.update(trainer_kwargs)
            self._trainer = SeldTrainer(self._sd, n_classes=self.n_classes, label_rate=self.label_rate,
                                        feature_rate=self.feature_rate or 80.0, device=self.device, **kw)
        self.training = True
        return self

    @property
    def trainer(self):
        return self._trainer

    def training_step(self, train_batch, batch_idx=None):
        """The reference's `training_step` (models/seld_models.py:68-76) INCLUDING what Lightning does around it (backward,
        gradient all-reduce, optimiser step): train_batch = (x, y_sed, y_doa, filenames) as the data loader yields them.
        Returns {'loss', 'sed_loss', 'doa_loss'} as 0-d CUDA tensors."""
        if not self.training or self._trainer is None:
            raise RuntimeError('call train() first')
        x, y_sed, y_doa = train_batch[0], train_batch[1], train_batch[2]
        dev = self.device
        loss = self._trainer.step(x.to(dev, torch.float32), {'event_frame_gt': y_sed.to(dev, torch.float32),
                                                             'doa_frame_gt': y_doa.to(dev, torch.float32)})
        return {'loss': loss[0], 'sed_loss': loss[1], 'doa_loss': loss[2]}

    def state_dict(self):
        """Reference state-dict keys: the trainer's current weights while training, else the loaded ones."""
        if self.training and self._trainer is not None:
            return self._trainer.state_dict()
        if self._sd is None:
            raise RuntimeError('load_state_dict() first')
        return dict(self._sd)

    def __call__(self, x):
        return self.forward(x)

    # ---- weights ----------------------------------------------------------------------------------
    def _fold(self, sd, conv_key, bn_prefix, pad_to=64):
        """conv weight (Cout,Cin,k,k) + eval BatchNorm -> bf16 (k*k, Cout, planes*Cin_pad), fp32 bias (Cout,)."""
        w = _t(sd[conv_key])
        # float32 arithmetic with a correctly rounded square root (NumPy; torch's vectorised CPU sqrt is off by an ulp for
        # some inputs), so that the library's own folding (crnn_model.cu: fold_conv) gives the same bits
        var = _t(sd[bn_prefix + '.running_var']).numpy()
        scale = torch.from_numpy(_t(sd[bn_prefix + '.weight']).numpy() / np.sqrt(var + np.float32(BN_EPS)))
        bias = _t(sd[bn_prefix + '.bias']) - _t(sd[bn_prefix + '.running_mean']) * scale
        w = w * scale[:, None, None, None]
        cout, cin, k, _ = w.shape
        cin_pad = (cin + pad_to - 1) // pad_to * pad_to
        wp = torch.zeros((k * k, cout, cin_pad))
        wp[:, :, :cin] = w.permute(2, 3, 0, 1).reshape(k * k, cout, cin)
        return ops.split_planes(wp, self.planes).to(self.device), bias.contiguous().to(self.device)

    def load_state_dict(self, state_dict, strict: bool = True):
        sd = state_dict
        W = {}
        W['cb1'] = self._fold(sd, 'encoder.conv_block1.conv1.weight', 'encoder.conv_block1.bn1', pad_to=16)
        W['cb2'] = self._fold(sd, 'encoder.conv_block1.conv2.weight', 'encoder.conv_block1.bn2')
        for li in range(1, 5):
            for bi in range(2):
                p = 'encoder.resnet.layer{}.{}'.format(li, bi)
                W[(li, bi, 1)] = self._fold(sd, p + '.conv1.weight', p + '.bn1')
                W[(li, bi, 2)] = self._fold(sd, p + '.conv2.weight', p + '.bn2')
                if li > 1 and bi == 0:
                    W[(li, bi, 'ds')] = self._fold(sd, p + '.downsample.1.weight', p + '.downsample.2')
        dev = self.device
        for layer in range(2):
            w_ih = torch.cat([_t(sd['decoder.gru.weight_ih_l{}'.format(layer)]), _t(sd['decoder.gru.weight_ih_l{}_reverse'.format(layer)])])
            b_ih = torch.cat([_t(sd['decoder.gru.bias_ih_l{}'.format(layer)]), _t(sd['decoder.gru.bias_ih_l{}_reverse'.format(layer)])])
            w_hh = torch.stack([_t(sd['decoder.gru.weight_hh_l{}'.format(layer)]), _t(sd['decoder.gru.weight_hh_l{}_reverse'.format(layer)])])
            b_hh = torch.stack([_t(sd['decoder.gru.bias_hh_l{}'.format(layer)]), _t(sd['decoder.gru.bias_hh_l{}_reverse'.format(layer)])])
            W[('gru', layer)] = (ops.split_planes(w_ih, self.planes).to(dev), b_ih.contiguous().to(dev),
                                 w_hh.contiguous().to(dev), b_hh.contiguous().to(dev))
        heads = ('event', 'x', 'y', 'z')
        n = self.n_classes
        w1 = torch.cat([_t(sd['decoder.{}_fc_1.weight'.format(h)]) for h in heads])            # (1024, 512)
        b1 = torch.cat([_t(sd['decoder.{}_fc_1.bias'.format(h)]) for h in heads])
        w2 = torch.zeros((64, 1024))                                                             # block diagonal
        b2 = torch.zeros((64,))
        for i, h in enumerate(heads):
            w2[i * n:(i + 1) * n, i * 256:(i + 1) * 256] = _t(sd['decoder.{}_fc_2.weight'.format(h)])
            b2[i * n:(i + 1) * n] = _t(sd['decoder.{}_fc_2.bias'.format(h)])
        W['fc1'] = (ops.split_planes(w1, self.planes).to(dev), b1.contiguous().to(dev))
        W['fc2'] = (ops.split_planes(w2, self.planes).to(dev), b2.contiguous().to(dev))
        self._w = W
        self._sd = {k: _t(v).clone() for k, v in sd.items()}
        self._load_native(sd)
        return self

    def _load_native(self, sd):
        """The same state dict through `crnn_load_weights` (include/salsa_crnn.h): the library folds and packs the weights
        itself and keeps them for `crnn_forward`, the one-call forward a non-Python host uses."""
        self._release_native()
        keep, table = [], []
        for name, value in sd.items():
            if not (name.startswith('encoder.') or name.startswith('decoder.')) or name.endswith('num_batches_tracked'):
                continue
            arr = np.ascontiguousarray(_t(value).numpy(), dtype=np.float32)
            keep.append(arr)
            table.append(_native.CrnnTensor(name.encode(), arr.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), arr.size))
        arr_t = (_native.CrnnTensor * len(table))(*table)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _native.check(_native.lib().crnn_load_weights(ctypes.cast(arr_t, ctypes.c_void_p), len(table), self.planes, self.n_classes,
                                                          ctypes.byref(handle)))
        self._native_model = handle

    # ---- forward ----------------------------------------------------------------------------------
    def set_scaler(self, mean, std):
        """Fuse the data layer's normalisation into the input packing: channels 0..n-1 become (x - mean) / std with
        mean, std (n, 1, F) as stored in `<fmt>_feature_scaler.h5` (database.py:87-96, :196-202).  None switches it off."""
        if mean is None:
            self._scaler = None
        else:
            self._scaler = (torch.as_tensor(np.asarray(mean), dtype=torch.float32).to(self.device),
                            torch.as_tensor(np.asarray(std), dtype=torch.float32).to(self.device))
        return self

    def encode(self, x, n_frames=None):
        """PannResNet22.forward: (B,7,T,F) fp32 CUDA -> (B, T/16, F/16, planes*512) bf16 NHWC.
        `n_frames` keeps only the first frames of x (the reference trims 4801 -> 4800 before the model,
        database.py:205-207) without a copy."""
        if self._w is None:
            raise RuntimeError('load_state_dict() first')
        if x.dim() != 4 or x.shape[1] != self.encoder.n_input_channels:
            raise ValueError('x must be (batch_size, {}, n_timesteps, n_features)'.format(self.encoder.n_input_channels))
        W, P = self._w, self.planes
        if self.encoder.n_input_channels > 16:
            raise NotImplementedError('at most 16 input channels')
        h = ops.pack_input(x.to(self.device, torch.float32), t_use=n_frames, c_pad=16, planes=P, scaler=self._scaler)
        h = ops.conv_first(h, *W['cb1'], relu=True, planes=P)
        h = ops.conv2d(h, *W['cb2'], relu=True, planes=P, pool=True)          # + F.avg_pool2d (model_utils.py:220)
        for li in range(1, 5):
            for bi in range(2):
                if li > 1 and bi == 0:
                    # `h` arrives already pooled: the producer of a stride-2 block's input applies the block's
                    # avg_pool2d (model_utils.py:349, :476) in its epilogue
                    identity = ops.conv2d(h, *W[(li, bi, 'ds')], planes=P)
                    out = ops.conv2d(h, *W[(li, bi, 1)], relu=True, planes=P)
                else:
                    identity = h
                    out = ops.conv2d(h, *W[(li, bi, 1)], relu=True, planes=P)
                feeds_stride2 = bi == 1 and li < 4
                h = ops.conv2d(out, *W[(li, bi, 2)], residual=identity, relu=True, planes=P, pool=feeds_stride2)
        return h

    def decode(self, enc):
        """SeldDecoder.forward on the NHWC encoder output."""
        W, P = self._w, self.planes
        B, T, _, _ = enc.shape
        rows = B * T
        h = ops.freq_mean(enc, planes=P)                                    # (rows_pad, P*512)
        for layer in range(2):
            w_ih, b_ih, w_hh, b_hh = W[('gru', layer)]
            xproj = ops.gemm(h, w_ih, b_ih, M=rows, out_f32=True, planes=P)
            h = ops.gru_layer(xproj, w_hh, b_hh, B, T, planes=P)
        f1 = ops.gemm(h, *W['fc1'], relu=True, M=rows, planes=P)
        z = ops.gemm(f1, *W['fc2'], M=rows, out_f32=True, planes=P)
        logits, doa = ops.head_finish(z, rows, self.n_classes)
        return {'event_frame_logit': logits.reshape(B, T, self.n_classes),
                'doa_frame_output': doa.reshape(B, T, 3 * self.n_classes)}

    def forward(self, x, n_frames=None):
        """x: (batch_size, n_channels, n_timesteps, n_features) -> the reference's output dict (seld_models.py:39-49).
        One native call (`crnn_forward`: the whole layer schedule on a workspace this object owns, replayed as a CUDA graph)."""
        if self._native_model is None:
            raise RuntimeError('load_state_dict() first')
        if x.dim() != 4 or x.shape[1] != self.encoder.n_input_channels or self.encoder.n_input_channels != 7:
            raise ValueError('x must be (batch_size, 7, n_timesteps, n_features)')
        if not x.is_cuda:
            raise ValueError('x must be a CUDA tensor (salsa_b200 has no CPU fallback)')
        x = x.to(self.device, torch.float32).contiguous()
        B, _, T_in, F = x.shape
        T = T_in if n_frames is None else int(n_frames)
        lib = _native.lib()
        with _native.device_of(x) as st:
            need = lib.crnn_workspace_bytes(self._native_model, B, T, F)
            if need == 0:
                _native.check(_native.SALSA_EINVAL)
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != x.device:
                self._workspace = torch.empty(need, dtype=torch.uint8, device=x.device)
            Tp = T // 16
            # persistent output buffers (stable pointers keep the captured graph valid); the caller gets copies
            if self._out is None or self._out[0].shape[:2] != (B, Tp) or self._out[0].device != x.device:
                self._out = (torch.empty((B, Tp, self.n_classes), dtype=torch.float32, device=x.device),
                             torch.empty((B, Tp, 3 * self.n_classes), dtype=torch.float32, device=x.device))
            logits, doa = self._out
            mean = std = None
            n_scaled = 0
            if self._scaler is not None:
                mean = self._scaler[0].to(x.device, torch.float32).reshape(-1, F).contiguous()
                std = self._scaler[1].to(x.device, torch.float32).reshape(-1, F).contiguous()
                n_scaled = mean.shape[0]
            p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
            _native.check(lib.crnn_forward(self._native_model, p(x), B, T_in, T, F, p(mean), p(std), n_scaled, p(logits), p(doa),
                                           p(self._workspace), self._workspace.numel(), st))
            self._keep = (x, mean, std)          # the asynchronous call reads them
            return {'event_frame_logit': logits.clone(), 'doa_frame_output': doa.clone()}

    def forward_ops(self, x, n_frames=None):
        """The same forward operator by operator (`encode` + `decode`): every layer is its own native call on torch-allocated
        tensors.  Bit-identical to `forward`; kept for per-layer tests and profiling."""
        return self.decode(self.encode(x, n_frames))

    def predict(self, x, n_frames=None):
        """forward + interpolate_tensor to the label rate, as `common_step` does (seld_models.py:58-64)."""
        out = self.forward(x, n_frames)
        ratio = self.time_downsample_ratio * self.label_rate / self.feature_rate
        idx = ops.interpolate_index(out['event_frame_logit'].shape[1], ratio)
        return {k: ops.gather_time(v, idx) for k, v in out.items()}

    def common_step(self, batch_data):
        """models/seld_models.py:51-66: batch (x, y_sed, y_doa, filenames) -> (target_dict, pred_dict) at the label rate."""
        x, y_sed, y_doa = batch_data[0], batch_data[1], batch_data[2]
        dev = self.device
        target = {'event_frame_gt': y_sed.to(dev, torch.float32), 'doa_frame_gt': y_doa.to(dev, torch.float32)}
        pred = self.predict(x.to(dev, torch.float32))
        n = min(pred['event_frame_logit'].shape[1], target['event_frame_gt'].shape[1])
        pred = {k: v[:, :n].contiguous() for k, v in pred.items()}
        target = {k: v[:, :n].contiguous() for k, v in target.items()}
        return target, pred

    def validation_step(self, val_batch, batch_idx=None, sed_threshold: float = 0.3):
        """models/seld_models.py:84-95 without the file system: the three losses plus, per clip, the rows the reference would
        write to the submission csv (`events`' format)."""
        target, pred = self.common_step(val_batch)
        loss, sed_loss, doa_loss = self.compute_loss(target, pred, loss_weight=self._trainer_kwargs['loss_weight'])
        return {'loss': loss, 'sed_loss': sed_loss, 'doa_loss': doa_loss,
                'events': self.events(val_batch[0].to(self.device, torch.float32), sed_threshold=sed_threshold)}

    def compute_loss(self, target_dict, pred_dict, loss_weight=(0.3, 0.7)):
        """BaseModel.compute_loss (models/interfaces.py:273-286) for reg_xyz outputs at the label rate:
        target_dict['event_frame_gt' / 'doa_frame_gt'], pred_dict['event_frame_logit' / 'doa_frame_output'] (CUDA tensors)
        -> (loss, sed_loss, doa_loss) as 0-d float32 CUDA tensors (the reference logs them in common_step,
        models/seld_models.py:51-66)."""
        out = ops.seld_loss(pred_dict['event_frame_logit'], pred_dict['doa_frame_output'], target_dict['event_frame_gt'],
                            target_dict['doa_frame_gt'], loss_weight=loss_weight)
        return out[0], out[1], out[2]

    def events(self, x, sed_threshold: float = 0.3, max_nframes_per_file: int = None, eval_version: str = '2021', n_frames=None):
        """predict + the decoding of `write_classwise_output_to_file` (models/interfaces.py:210-258) for whole-clip inputs
        (one chunk per file, as the reference's test configuration): per clip the list of rows the reference writes to the
        submission csv, [frame, class, 0, azimuth, elevation] (2021) or [frame, class, azimuth, elevation]."""
        pred = self.predict(x, n_frames)
        logit, doa = pred['event_frame_logit'], pred['doa_frame_output']
        B, T, n = logit.shape
        active, azi, ele = ops.decode_events(logit.reshape(B * T, n), doa.reshape(B * T, 3 * n), sed_threshold)
        active = active.reshape(B, T, n).cpu().numpy()
        azi, ele = azi.reshape(B, T, n).cpu().numpy(), ele.reshape(B, T, n).cpu().numpy()
        n_frames = T if max_nframes_per_file is None else max_nframes_per_file
        assert T >= n_frames, 'n_output_frames of sed < max_nframes_per_file'
        rows = []
        for b in range(B):
            fr, cl = np.nonzero(active[b, :n_frames])
            if eval_version == '2021':
                rows.append([[int(f), int(c), 0, int(azi[b, f, c]), int(ele[b, f, c])] for f, c in zip(fr, cl)])
            else:
                rows.append([[int(f), int(c), int(azi[b, f, c]), int(ele[b, f, c])] for f, c in zip(fr, cl)])
        return rows
