"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container).

TEST INFRASTRUCTURE -- see `oracle/__init__.py`.

    python -m oracle.make_golden            # needs /root/reference

Every array stored here was produced by reference code executed verbatim
(`oracle/ref_import.py`): `extract_normalized_eigenvector`, `MagStftExtractor.extract`,
the per-clip driver bodies, `interpolate_tensor` and `SeldModel.forward`.  The only
non-reference arithmetic underneath is the `librosa.stft` / `power_to_db` shim
(`oracle/stft.py`), because librosa is not installable here.
"""
import os
import sys

import numpy as np

from . import ref_import, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

DATA_CFG = {
    'foa': dict(format='foa', fs=24000, n_fft=512, win_len=512, hop_len=300, fmin_doa=50, fmax_doa=9000),
    'mic': dict(format='mic', fs=24000, n_fft=512, win_len=512, hop_len=300, fmin_doa=50, fmax_doa=4000),
    'lite': dict(format='mic', fs=24000, n_fft=512, win_len=512, hop_len=300, fmin_doa=50, fmax_doa=2000),
}


def structured_spectrum(seed=7, n_bins=24, n_frames=60, n_chans=4):
    """Random complex spectrum with a dominant rank-1 part in a random subset of TF bins, so
    that tracker, coherence test and normalisation all see both outcomes."""
    rng = np.random.default_rng(seed)
    steer = rng.standard_normal((n_bins, 1, n_chans)) + 1j * rng.standard_normal((n_bins, 1, n_chans))
    steer[:, :, 0] = 1.0 + 0.2 * rng.standard_normal((n_bins, 1))
    src = rng.standard_normal((n_bins, n_frames, 1)) + 1j * rng.standard_normal((n_bins, n_frames, 1))
    envelope = (rng.uniform(size=(n_bins, n_frames // 10 + 1, 1)) > 0.4).repeat(10, axis=1)[:, :n_frames]
    noise = rng.standard_normal((n_bins, n_frames, n_chans)) + 1j * rng.standard_normal((n_bins, n_frames, n_chans))
    return 3.0 * envelope * src * steer + 0.15 * noise


def eigvec_cases():
    ref = ref_import.features_module()
    X = structured_spectrum()
    out = {'X': X}
    for fmt in ('foa', 'mic'):
        for tag, trk, cond in (('t5', True, 5.0), ('t0', True, 0.0), ('n5', False, 5.0), ('t2', True, 2.0)):
            out['{}_{}'.format(fmt, tag)] = ref.extract_normalized_eigenvector(
                X.copy(), condition_number=cond, n_hopframes=3, is_tracking=trk, audio_format=fmt,
                fs=24000, n_fft=512, lower_bin=1)
    return out


def clip_cases(seconds=1.0):
    out = {}
    ref = ref_import.features_module()
    foa = synth.make_clip(3, 'foa', seconds=seconds)
    mic = synth.make_clip(4, 'mic', seconds=seconds)
    out['audio_foa'] = foa
    out['audio_mic'] = mic
    out['salsa_foa'] = ref_import.run_driver_body('salsa', foa, DATA_CFG['foa'])
    out['salsa_mic'] = ref_import.run_driver_body('salsa', mic, DATA_CFG['mic'])
    out['salsa_foa_notracking'] = ref_import.run_driver_body('salsa', foa, DATA_CFG['foa'], is_tracking=False)
    out['salsa_lite'] = ref_import.run_driver_body('salsa_lite', mic, DATA_CFG['lite'], feature_type='salsa_lite')
    out['salsa_ipd'] = ref_import.run_driver_body('salsa_lite', mic, DATA_CFG['lite'], feature_type='salsa_ipd')
    out['logspec_foa'] = ref.MagStftExtractor(n_fft=512, hop_length=300, win_length=512).extract(foa)
    out['logspec_foa_nocompress'] = ref.MagStftExtractor(
        n_fft=512, hop_length=300, win_length=512, is_compress_high_freq=False).extract(foa)
    out['W512'] = ref.MagStftExtractor(n_fft=512, hop_length=300).W
    out['W256'] = ref.MagStftExtractor(n_fft=256, hop_length=150).W
    return out


def model_cases():
    import torch
    models = ref_import.models_module()
    from models.model_utils import interpolate_tensor
    out = {}
    x = torch.arange(24).reshape(2, -1, 3)
    out['interp_in'] = x.numpy()
    out['interp_half'] = interpolate_tensor(x, ratio=0.5).numpy()
    out['interp_double'] = interpolate_tensor(x, ratio=2.0).numpy()
    out['interp_40_to_80'] = interpolate_tensor(torch.arange(40).reshape(1, 40, 1), ratio=2.0).numpy().ravel()

    # the reference modules, loaded with the deterministic random state dict of oracle/crnn.py (the
    # reference's own init zeroes bn2.weight, which would switch every residual branch off)
    from . import crnn as ocrnn
    model = ref_import.build_reference_seld_model()
    missing = model.load_state_dict(ocrnn.make_state_dict(0), strict=True)
    model.eval()
    xin = ocrnn.model_input(2, (2, 7, 128, 200))
    with torch.no_grad():
        y = model(xin)
        enc = model.encoder(xin)
    out['model_encoder_out'] = enc.numpy().astype(np.float16)       # checked at 1e-3: float16 is enough
    out['model_event_frame_logit'] = y['event_frame_logit'].numpy()
    out['model_doa_frame_output'] = y['doa_frame_output'].numpy()
    # a second, ragged case: F = 191 (SALSA-Lite), odd batch
    xin2 = ocrnn.model_input(3, (1, 7, 96, 191))
    with torch.no_grad():
        y2 = model(xin2)
    out['lite_event_frame_logit'] = y2['event_frame_logit'].numpy()
    out['lite_doa_frame_output'] = y2['doa_frame_output'].numpy()
    # BaseModel.compute_loss (interfaces.py:273-355) on the deterministic inputs of oracle/crnn.py: seld_loss_inputs
    logit, doa, event_gt, doa_gt = ocrnn.seld_loss_inputs()
    loss = model.compute_loss(target_dict={'event_frame_gt': event_gt, 'doa_frame_gt': doa_gt},
                              pred_dict={'event_frame_logit': logit, 'doa_frame_output': doa})
    out['loss_values'] = np.array([float(v) for v in loss], dtype=np.float64)
    return out


def augment_cases():
    """Outputs of the unmodified reference augmentation classes (utilities/transforms.py) under np.random.seed:
    the joint FOA / MIC channel swaps and the frequency shift, composed as dataset/datamodule.py:45-83 composes them for
    the SALSA features.  Inputs are regenerated from `seed`; draws are replayed by oracle/augment.py's draw_* functions."""
    T = ref_import.transforms_module()
    out = {}
    rng = np.random.default_rng(5)
    x = rng.standard_normal((7, 24, 40)).astype(np.float32)
    y_sed = (rng.random((6, 12)) > 0.5).astype(np.float32)
    y_doa = rng.standard_normal((6, 36)).astype(np.float32)
    out['x'], out['y_sed'], out['y_doa'] = x, y_sed, y_doa
    for fmt, cls in (('foa', T.TfmapRandomSwapChannelFoa), ('mic', T.TfmapRandomSwapChannelMic)):
        joint = T.ComposeMapTransform([cls(n_classes=12)])
        single = T.ComposeTransformNp([T.RandomShiftUpDownNp(freq_shift_range=10)])
        for seed in range(24):
            np.random.seed(seed)
            xa, ya_sed, ya_doa = joint(x, y_sed, y_doa)            # datamodule order: joint transform, then the shift
            xa = single(xa)
            out['{}_{}_x'.format(fmt, seed)] = np.ascontiguousarray(xa)
            out['{}_{}_y_doa'.format(fmt, seed)] = np.ascontiguousarray(ya_doa)
            assert np.array_equal(ya_sed, y_sed)
    return out


def extras_cases():
    """Round-2 additions, again from the unmodified reference: the MIC training transforms with CompositeCutout behind the
    frequency shift (dataset/datamodule.py:76-82) under np.random.seed, and LinSpecIvExtractor (dataset/feature_extraction.py:
    273-358) on the golden FOA clip."""
    T = ref_import.transforms_module()
    out = {}
    rng = np.random.default_rng(11)
    x = (rng.standard_normal((7, 32, 48)) * 10.0 - 40.0).astype(np.float32)
    y_sed = (rng.random((6, 12)) > 0.5).astype(np.float32)
    y_doa = rng.standard_normal((6, 36)).astype(np.float32)
    out['cut_x'], out['cut_y_sed'], out['cut_y_doa'] = x, y_sed, y_doa
    joint = T.ComposeMapTransform([T.TfmapRandomSwapChannelMic(n_classes=12)])
    single = T.ComposeTransformNp([T.RandomShiftUpDownNp(freq_shift_range=10),
                                   T.CompositeCutout(image_aspect_ratio=32 / 48, n_zero_channels=3)])
    for seed in range(30):
        np.random.seed(seed)
        xa, _, ya_doa = joint(x, y_sed, y_doa)
        xa = single(xa)
        out['cut_{}_x'.format(seed)] = np.ascontiguousarray(xa)
        out['cut_{}_y_doa'.format(seed)] = np.ascontiguousarray(ya_doa)
    fe = ref_import.other_features_module()
    foa = np.load(os.path.join(GOLDEN_DIR, 'clip_cases.npz'))['audio_foa']
    out['linspeciv_foa'] = fe.LinSpecIvExtractor(n_fft=512, hop_length=300, win_length=512).extract(foa).astype(np.float32)
    mic = np.load(os.path.join(GOLDEN_DIR, 'clip_cases.npz'))['audio_mic'][:, :12000]          # 0.5 s: 41 frames
    out['linspecgcc_mic'] = fe.LogSpecGccExtractor(n_fft=512, hop_length=300, win_length=512).extract(mic).astype(np.float32)
    return out


def nfft256_cases():
    """The n_fft = 256 configuration the reference accepts (salsa_feature_extraction.py:151-152, :163-170, :300-306), from
    the unmodified reference on the first 0.5 s of the golden clips: MagStftExtractor (both band layouts, a shorter window), the
    per-clip SALSA body, FOA and MIC, and the SALSA-Lite / SALSA-IPD body, at hop 150."""
    ref = ref_import.features_module()
    clips = np.load(os.path.join(GOLDEN_DIR, 'clip_cases.npz'))
    foa, mic = clips['audio_foa'][:, :12000], clips['audio_mic'][:, :12000]
    cfg = lambda base: dict(DATA_CFG[base], n_fft=256, win_len=256, hop_len=150)
    out = {}
    out['salsa_foa'] = ref_import.run_driver_body('salsa', foa, cfg('foa'))
    out['salsa_mic'] = ref_import.run_driver_body('salsa', mic, cfg('mic'))
    out['logspec_foa'] = ref.MagStftExtractor(n_fft=256, hop_length=150, win_length=256).extract(foa)
    out['logspec_foa_nocompress'] = ref.MagStftExtractor(n_fft=256, hop_length=150, win_length=256, is_compress_high_freq=False).extract(foa)
    out['logspec_foa_win200'] = ref.MagStftExtractor(n_fft=256, hop_length=150, win_length=200).extract(foa)
    lite_cfg = dict(DATA_CFG['lite'], n_fft=256, win_len=256, hop_len=150)
    out['salsa_lite'] = ref_import.run_driver_body('salsa_lite', mic, lite_cfg, feature_type='salsa_lite')
    out['salsa_ipd'] = ref_import.run_driver_body('salsa_lite', mic, lite_cfg, feature_type='salsa_ipd')
    return out


def main(argv=None):
    if not ref_import.available():
        print('reference checkout not found; golden vectors can only be generated in the build container')
        return 1
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if not argv or 'features' in argv:
        np.savez_compressed(os.path.join(GOLDEN_DIR, 'eigvec_cases.npz'), **eigvec_cases())
        np.savez_compressed(os.path.join(GOLDEN_DIR, 'clip_cases.npz'), **clip_cases())
    if argv and 'model' in argv:
        # inputs and weights are regenerated from their seeds (oracle/crnn.py); only outputs are stored
        np.savez_compressed(os.path.join(GOLDEN_DIR, 'model_cases.npz'), **model_cases())
    if argv and 'augment' in argv:
        np.savez_compressed(os.path.join(GOLDEN_DIR, 'augment_cases.npz'), **augment_cases())
    if argv and 'extras' in argv:
        np.savez_compressed(os.path.join(GOLDEN_DIR, 'extras_cases.npz'), **extras_cases())
    if argv and 'nfft256' in argv:
        np.savez_compressed(os.path.join(GOLDEN_DIR, 'nfft256_cases.npz'), **nfft256_cases())
    for fn in sorted(os.listdir(GOLDEN_DIR)):
        print('{:32s} {:10d} B'.format(fn, os.path.getsize(os.path.join(GOLDEN_DIR, fn))))
    return 0


if __name__ == '__main__':
    sys.exit(main(sys.argv[1:]))
