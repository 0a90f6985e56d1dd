"""Spectrogram front-end of the reference on the device.

The reference builds its input spectrograms with madmom (tutorials/Embedding Tutorial.ipynb cell 28; the microphone
stream of audio_sheet_server.py:44-60 uses the same chain):

    SignalProcessor(num_channels=1, sample_rate=22050)
    FramedSignalProcessor(frame_size=2048, fps=20, origin='future')
    FilteredSpectrogramProcessor(LogarithmicFilterbank, num_bands=16, fmin=30, fmax=6000)
    LogarithmicSpectrogramProcessor()

`LogSpectrogramProcessor` is that chain from PCM samples on (decoding audio files is not part of it): the filterbank
is built once on the host exactly as madmom 0.15.1 does (log-spaced centre frequencies, closest FFT bins made unique,
unit-area triangles -- 92 bands for the parameters above, the "92 Frequency Bins" of the notebook), framing, Hann
window, FFT magnitude, filtering and log10(1 + x) run in one kernel (asr_log_spectrogram).
"""
import numpy as np
import torch

from .. import _lib

SAMPLE_RATE, FRAME_SIZE, FPS = 22050, 2048, 20


def _log_frequencies(bands_per_octave, fmin, fmax, fref=440.0):
    left = np.floor(np.log2(float(fmin) / fref) * bands_per_octave)
    right = np.ceil(np.log2(float(fmax) / fref) * bands_per_octave)
    freqs = fref * 2.0 ** (np.arange(left, right) / float(bands_per_octave))
    freqs = freqs[np.searchsorted(freqs, fmin):]
    return freqs[:np.searchsorted(freqs, fmax, 'right')]


def logarithmic_filterbank(sample_rate=SAMPLE_RATE, frame_size=FRAME_SIZE, num_bands=16, fmin=30.0, fmax=6000.0):
    """(frame_size / 2, n_bands) float32: LogarithmicFilterbank(..., norm_filters=True, unique_filters=True)."""
    n_bins = frame_size >> 1
    bin_freqs = np.fft.fftfreq(frame_size, 1.0 / sample_rate)[:n_bins]
    freqs = _log_frequencies(num_bands, fmin, fmax)
    idx = np.clip(bin_freqs.searchsorted(freqs), 1, n_bins - 1)
    idx = idx - (freqs - bin_freqs[idx - 1] < bin_freqs[idx] - freqs)        # the closer of the two neighbouring bins
    bins = np.unique(idx)
    if len(bins) < 3:
        raise ValueError("not enough FFT bins between fmin and fmax for a triangular filterbank")
    fb = np.zeros((n_bins, len(bins) - 2), np.float64)
    for b in range(len(bins) - 2):
        start, center, stop = int(bins[b]), int(bins[b + 1]), int(bins[b + 2])
        if stop - start < 2:
            center, stop = start + 1, start + 2
        tri = np.concatenate([np.linspace(0, 1, center - start, endpoint=False),
                              np.linspace(1, 0, stop - center, endpoint=False)])
        fb[start:stop, b] = np.maximum(fb[start:stop, b], tri / tri.sum())
    return fb.astype(np.float32)


class LogSpectrogramProcessor(object):
    """processor.process(samples) -> (n_bands, n_frames) float32, like `processor.process(path).T` of the reference."""

    def __init__(self, sample_rate=SAMPLE_RATE, frame_size=FRAME_SIZE, fps=FPS, num_bands=16, fmin=30.0, fmax=6000.0,
                 device=None):
        self.sample_rate, self.frame_size, self.fps = int(sample_rate), int(frame_size), float(fps)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        fb = logarithmic_filterbank(sample_rate, frame_size, num_bands, fmin, fmax)
        nz = fb > 0
        lo = np.argmax(nz, axis=0).astype(np.int32)
        hi = (fb.shape[0] - np.argmax(nz[::-1], axis=0)).astype(np.int32)
        self.n_bands = int(fb.shape[1])
        self.filterbank = torch.as_tensor(fb).to(self.device).contiguous()
        self._lo = torch.as_tensor(lo).to(self.device)
        self._hi = torch.as_tensor(hi).to(self.device)

    def num_frames(self, n_samples):
        return int(_lib.lib.asr_spectrogram_num_frames(int(n_samples), self.sample_rate, self.fps))

    def process_device(self, samples):
        """samples: 1-D mono signal (NumPy or torch; int16 is scaled by 1 / 32768 as madmom scales its window, floats
        are taken as they are).  -> (n_bands, n_frames) float32 CUDA tensor."""
        t = torch.as_tensor(samples)
        if t.dim() != 1 or t.numel() == 0:
            raise ValueError("expected a non-empty mono signal")
        if t.dtype == torch.int16:
            t = t.to(self.device).to(torch.float32) / 32768.0
        else:
            t = t.to(self.device, torch.float32)
        t = t.contiguous()
        n_frames = self.num_frames(t.numel())
        out = torch.empty((self.n_bands, n_frames), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.asr_log_spectrogram(_lib.dptr(t), int(t.numel()), self.sample_rate, self.frame_size, self.fps,
                                                _lib.dptr(self.filterbank), _lib.dptr(self._lo), _lib.dptr(self._hi),
                                                self.n_bands, _lib.dptr(out), _lib.stream_ptr()))
        return out

    def process(self, samples):
        return self.process_device(samples).cpu().numpy()
