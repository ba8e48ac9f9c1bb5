"""Audio in, SELD outputs out, without leaving the GPU: SALSA features computed on the fly and fed to the CRNN.

The reference goes through the file system: `make salsa` writes (7, 4801, 200) float32 features to h5
(dataset/salsa_feature_extraction.py:377-382), `Database` reads them back, normalises channels 0..3 with the scaler and
trims to a multiple of the label resolution (dataset/database.py:196-207), and `SeldModel.forward` consumes them
(models/seld_models.py:39-49).  Here the same three steps are three stages on device memory: per 60 s clip 23 MB of audio
cross PCIe instead of 27 MB of features, and nothing is written in between (BASELINE.json config 5, inference side;
SURVEY.md section 8 f2); with the wav files' own 16-bit samples as input, 11.5 MB."""
import torch

from .crnn import SeldModel
from .features import SalsaExtractor

__all__ = ['SeldPipeline']


class SeldPipeline:
    def __init__(self, extractor: SalsaExtractor, model: SeldModel, scaler=None):
        """scaler: (mean, std) as `compute_scaler` / `FeatureScaler.finalize` return them, fused into the CRNN's input
        packing (`SeldModel.set_scaler`); None when the model already has one or the features are not normalised."""
        self.extractor, self.model = extractor, model
        if scaler is not None:
            model.set_scaler(*scaler)
        self._features = None
        self._audio = None

    def features(self, audio: torch.Tensor) -> torch.Tensor:
        """(B, 4, N) float32 CUDA -- or int16, the 16-bit PCM samples of the wav files, converted on the device like
        librosa.load does on the host (dataset/salsa_feature_extraction.py:353): half the bytes over PCIe -- ->
        (B, 7, T, F) float32 CUDA, into a buffer that is reused between calls."""
        if audio.dtype == torch.int16:
            from .driver import pcm16_to_float
            if self._audio is None or self._audio.shape != audio.shape or self._audio.device != audio.device:
                self._audio = torch.empty(audio.shape, dtype=torch.float32, device=audio.device)
            audio = pcm16_to_float(audio, out=self._audio)
        B, T = audio.shape[0], self.extractor.n_frames(audio.shape[2])
        shape = (B, 7, T, self.extractor.freq_dim)
        if self._features is None or tuple(self._features.shape) != shape or self._features.device != audio.device:
            self._features = torch.empty(shape, dtype=torch.float32, device=audio.device)
        return self.extractor.extract(audio, out=self._features)

    def _n_frames(self, T: int) -> int:
        r = int(self.model.time_downsample_ratio)
        return (T // r) * r                         # 4801 -> 4800 (database.py:205-207)

    def forward(self, audio: torch.Tensor):
        """-> {'event_frame_logit': (B, T/16, n_classes), 'doa_frame_output': (B, T/16, 3 n_classes)}"""
        x = self.features(audio)
        return self.model.forward(x, n_frames=self._n_frames(x.shape[2]))

    __call__ = forward

    def predict(self, audio: torch.Tensor):
        """forward + interpolate_tensor to the label rate (seld_models.py:58-64)."""
        x = self.features(audio)
        return self.model.predict(x, n_frames=self._n_frames(x.shape[2]))

    def events(self, audio: torch.Tensor, **kwargs):
        """predict + the csv rows of write_classwise_output_to_file (models/interfaces.py:210-258), per clip."""
        x = self.features(audio)
        return self.model.events(x, n_frames=self._n_frames(x.shape[2]), **kwargs)
