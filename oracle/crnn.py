"""Plain PyTorch fp32 restatement of the reference SELD CRNN forward (eval mode).

TEST INFRASTRUCTURE -- see `oracle/__init__.py`.

Functional form driven by a reference-format state dict (`SeldModel.state_dict()` keys:
`encoder.conv_block1.*`, `encoder.resnet.layer{1..4}.{0,1}.*`, `decoder.gru.*`, `decoder.*_fc_{1,2}.*`).
Follows `models/encoders.py:48-56` (PannResNet22.forward), `models/model_utils.py:213-228` (ConvBlock),
`:345-367` (_ResnetBasicBlock), `:474-481` (downsample), `models/decoders.py:106-154`
(SeldDecoder.forward, bigru + avg pooling) and `models/model_utils.py:57-75` (interpolate_tensor).

Pinned by `tests/golden/model_cases.npz`: outputs of the UNMODIFIED reference modules (run behind
the Lightning stub of `oracle/ref_import.py`) loaded with `make_state_dict(0)`.
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5           # nn.BatchNorm2d default, used by every BN of the reference
N_CLASSES = 12


def _conv_keys():
    """[(conv weight key, bn prefix)] of every conv + BN pair, in forward order."""
    keys = [('encoder.conv_block1.conv1.weight', 'encoder.conv_block1.bn1'),
            ('encoder.conv_block1.conv2.weight', 'encoder.conv_block1.bn2')]
    for li in range(1, 5):
        for bi in range(2):
            p = 'encoder.resnet.layer{}.{}'.format(li, bi)
            keys += [(p + '.conv1.weight', p + '.bn1'), (p + '.conv2.weight', p + '.bn2')]
            if li > 1 and bi == 0:
                keys.append((p + '.downsample.1.weight', p + '.downsample.2'))
    return keys


def make_state_dict(seed: int = 0) -> dict:
    """A deterministic, fully random reference-format state dict (every BN affine / running statistic
    and every bias is non-trivial, unlike the reference's own init where bn2.weight = 0 switches the
    residual branches off, model_utils.py:343)."""
    g = torch.Generator().manual_seed(1000 + seed)
    sd = {}

    def uniform(shape, lo, hi):
        return torch.empty(shape).uniform_(lo, hi, generator=g)

    planes = {1: 64, 2: 128, 3: 256, 4: 512}
    shapes = {'encoder.conv_block1.conv1.weight': (64, 7, 3, 3), 'encoder.conv_block1.conv2.weight': (64, 64, 3, 3)}
    inpl = 64
    for li in range(1, 5):
        for bi in range(2):
            p = 'encoder.resnet.layer{}.{}'.format(li, bi)
            shapes[p + '.conv1.weight'] = (planes[li], inpl if bi == 0 else planes[li], 3, 3)
            shapes[p + '.conv2.weight'] = (planes[li], planes[li], 3, 3)
            if li > 1 and bi == 0:
                shapes[p + '.downsample.1.weight'] = (planes[li], inpl, 1, 1)
        inpl = planes[li]
    for wkey, bn in _conv_keys():
        shape = shapes[wkey]
        fan_in = shape[1] * shape[2] * shape[3]
        sd[wkey] = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        c = shape[0]
        sd[bn + '.weight'] = uniform((c,), 0.5, 1.5)
        sd[bn + '.bias'] = uniform((c,), -0.2, 0.2)
        sd[bn + '.running_mean'] = uniform((c,), -0.2, 0.2)
        sd[bn + '.running_var'] = uniform((c,), 0.5, 1.5)
        sd[bn + '.num_batches_tracked'] = torch.zeros((), dtype=torch.int64)
    for layer in range(2):
        for suffix in ('', '_reverse'):
            sd['decoder.gru.weight_ih_l{}{}'.format(layer, suffix)] = uniform((768, 512), -1 / 16, 1 / 16)
            sd['decoder.gru.weight_hh_l{}{}'.format(layer, suffix)] = uniform((768, 256), -1 / 16, 1 / 16)
            sd['decoder.gru.bias_ih_l{}{}'.format(layer, suffix)] = uniform((768,), -0.1, 0.1)
            sd['decoder.gru.bias_hh_l{}{}'.format(layer, suffix)] = uniform((768,), -0.1, 0.1)
    for head in ('event', 'x', 'y', 'z'):
        sd['decoder.{}_fc_1.weight'.format(head)] = torch.randn((256, 512), generator=g) / 512 ** 0.5
        sd['decoder.{}_fc_1.bias'.format(head)] = uniform((256,), -0.1, 0.1)
        sd['decoder.{}_fc_2.weight'.format(head)] = torch.randn((N_CLASSES, 256), generator=g) / 256 ** 0.5
        sd['decoder.{}_fc_2.bias'.format(head)] = uniform((N_CLASSES,), -0.1, 0.1)
    # reference key order (encoder convs before their BNs inside conv_block1) does not matter for load_state_dict
    return sd


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], sd[prefix + '.weight'],
                        sd[prefix + '.bias'], training=False, eps=BN_EPS)


def encoder_forward(sd, x):
    """PannResNet22.forward (encoders.py:48-56), eval mode: (B,7,T,F) -> (B,512,T/16,F/16)."""
    p = 'encoder.conv_block1'
    x = F.relu(_bn(F.conv2d(x, sd[p + '.conv1.weight'], padding=1), sd, p + '.bn1'))
    x = F.relu(_bn(F.conv2d(x, sd[p + '.conv2.weight'], padding=1), sd, p + '.bn2'))
    x = F.avg_pool2d(x, kernel_size=(2, 2))
    for li in range(1, 5):
        for bi in range(2):
            p = 'encoder.resnet.layer{}.{}'.format(li, bi)
            stride2 = li > 1 and bi == 0
            identity = x
            out = F.avg_pool2d(x, kernel_size=(2, 2)) if stride2 else x
            out = F.relu(_bn(F.conv2d(out, sd[p + '.conv1.weight'], padding=1), sd, p + '.bn1'))
            out = _bn(F.conv2d(out, sd[p + '.conv2.weight'], padding=1), sd, p + '.bn2')
            if stride2:
                identity = F.avg_pool2d(identity, kernel_size=2)
                identity = _bn(F.conv2d(identity, sd[p + '.downsample.1.weight']), sd, p + '.downsample.2')
            x = F.relu(out + identity)
    return x


def gru_forward(sd, x):
    """2-layer bidirectional GRU (decoders.py:44-46, :126), eval mode (no inter-layer dropout).
    Written out step by step: gate order r, z, n; n = tanh(W_in x + b_in + r * (W_hn h + b_hn))."""
    B, T, _ = x.shape
    for layer in range(2):
        outs = []
        for suffix in ('', '_reverse'):
            w_ih, w_hh = sd['decoder.gru.weight_ih_l{}{}'.format(layer, suffix)], sd['decoder.gru.weight_hh_l{}{}'.format(layer, suffix)]
            b_ih, b_hh = sd['decoder.gru.bias_ih_l{}{}'.format(layer, suffix)], sd['decoder.gru.bias_hh_l{}{}'.format(layer, suffix)]
            xp = x @ w_ih.T + b_ih
            h = torch.zeros(B, 256, dtype=x.dtype)
            ys = [None] * T
            order = range(T - 1, -1, -1) if suffix else range(T)
            for t in order:
                hp = h @ w_hh.T + b_hh
                r = torch.sigmoid(xp[:, t, :256] + hp[:, :256])
                z = torch.sigmoid(xp[:, t, 256:512] + hp[:, 256:512])
                n = torch.tanh(xp[:, t, 512:] + r * hp[:, 512:])
                h = (1 - z) * n + z * h
                ys[t] = h
            outs.append(torch.stack(ys, dim=1))
        x = torch.cat(outs, dim=-1)
    return x


def decoder_forward(sd, x):
    """SeldDecoder.forward (decoders.py:106-154) with decoder_type='bigru', freq_pool='avg', eval mode."""
    x = torch.mean(x, dim=3).transpose(1, 2)
    x = gru_forward(sd, x)

    def head(name):
        h = F.relu(F.linear(x, sd['decoder.{}_fc_1.weight'.format(name)], sd['decoder.{}_fc_1.bias'.format(name)]))
        return F.linear(h, sd['decoder.{}_fc_2.weight'.format(name)], sd['decoder.{}_fc_2.bias'.format(name)])

    event = head('event')
    doa = torch.cat([torch.tanh(head('x')), torch.tanh(head('y')), torch.tanh(head('z'))], dim=-1)
    return {'event_frame_logit': event, 'doa_frame_output': doa}


def forward(sd, x):
    """SeldModel.forward (seld_models.py:39-49)."""
    with torch.no_grad():
        return decoder_forward(sd, encoder_forward(sd, x))


def interpolate_tensor(tensor, ratio: float = 1.0):
    """model_utils.py:57-75."""
    ratio = float(ratio)
    n_input_frames = tensor.shape[1]
    n_output_frames = int(round(n_input_frames * ratio))
    output_idx = torch.arange(n_output_frames)
    input_idx = torch.floor(output_idx / ratio).long()
    return tensor[:, input_idx]


def decode_events(event_frame_logit, doa_frame_output, sed_threshold: float = 0.3, max_nframes_per_file: int = None,
                  eval_version: str = '2021'):
    """write_classwise_output_to_file (models/interfaces.py:210-258) for ONE file given as a single chunk
    (batch dimension 1): the rows of the submission csv."""
    n_classes = event_frame_logit.shape[-1]
    doa = doa_frame_output.detach().cpu().numpy()
    event = torch.sigmoid(event_frame_logit).detach().cpu().numpy()
    assert event.shape[0] == 1
    event, doa = event[0], doa[0]
    event = (event >= sed_threshold)
    n_frames = event.shape[0] if max_nframes_per_file is None else max_nframes_per_file
    assert event.shape[0] >= n_frames, 'n_output_frames of sed < max_nframes_per_file'
    x, y, z = doa[:, :n_classes], doa[:, n_classes:2 * n_classes], doa[:, 2 * n_classes:]
    azi_out = np.around(np.arctan2(y, x) * 180.0 / np.pi)
    ele_out = np.around(np.arctan2(z, np.sqrt(x ** 2 + y ** 2)) * 180.0 / np.pi)
    outputs = []
    for iframe in np.arange(n_frames):
        for class_idx in np.where(event[iframe] == 1)[0]:
            azi = int(azi_out[iframe, class_idx])
            if azi == 180:
                azi = -180
            ele = int(ele_out[iframe, class_idx])
            if eval_version == '2021':
                outputs.append([int(iframe), int(class_idx), 0, azi, ele])
            else:
                outputs.append([int(iframe), int(class_idx), azi, ele])
    return outputs


def seld_loss(event_logit, doa_output, event_gt, doa_gt, loss_weight=(0.3, 0.7), n_classes: int = 12):
    """BaseModel.compute_loss for output_format='reg_xyz' (models/interfaces.py:273-355): binary cross-entropy with
    logits on the event activity + masked MAE on the x, y, z regressions, each normalised by the number of active
    (frame, class) cells.  -> (loss, sed_loss, doa_loss) as float32 tensors."""
    sed_loss = F.binary_cross_entropy_with_logits(input=event_logit, target=event_gt)
    n = n_classes
    N = min(doa_output.shape[1], doa_gt.shape[1])
    mask = event_gt[:, :N]
    norm = torch.sum(mask)
    doa_loss = 0.0
    for i in range(3):
        doa_loss = doa_loss + torch.sum(torch.abs(doa_output[:, :N, i * n:(i + 1) * n] - doa_gt[:, :N, i * n:(i + 1) * n]) * mask) / norm
    return loss_weight[0] * sed_loss + loss_weight[1] * doa_loss, sed_loss, doa_loss


def seld_loss_inputs(seed: int = 4, batch: int = 3, n_frames: int = 80, n_classes: int = 12):
    """Deterministic (logit, doa, event_gt, doa_gt) of the label-rate shapes (B, 80, 12) / (B, 80, 36)."""
    g = torch.Generator().manual_seed(seed)
    logit = 3.0 * torch.randn((batch, n_frames, n_classes), generator=g)
    doa = torch.tanh(torch.randn((batch, n_frames, 3 * n_classes), generator=g))
    event_gt = (torch.rand((batch, n_frames, n_classes), generator=g) < 0.2).float()
    v = torch.randn((batch, n_frames, 3, n_classes), generator=g)
    v = v / v.norm(dim=2, keepdim=True)
    doa_gt = (v * event_gt[:, :, None, :]).reshape(batch, n_frames, 3 * n_classes)
    return logit, doa, event_gt, doa_gt


def model_input(seed: int = 2, shape=(2, 7, 128, 200)) -> torch.Tensor:
    """Deterministic feature-like input: log-spectrogram-scale first four channels, [-1, 1] spatial ones."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    x[:, 4:] = torch.tanh(x[:, 4:])
    return x


def to_numpy(d):
    return {k: np.asarray(v) for k, v in d.items()}
