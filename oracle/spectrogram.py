"""CPU restatement of the reference's spectrogram front-end -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference computes its input spectrograms with the third-party package madmom (requirements.txt: madmom==0.15.1,
absent from this image), configured in tutorials/Embedding Tutorial.ipynb cell 28:

    SignalProcessor(num_channels=1, sample_rate=22050)
    FramedSignalProcessor(frame_size=2048, fps=20, origin='future')
    FilteredSpectrogramProcessor(LogarithmicFilterbank, num_bands=16, fmin=30, fmax=6000)
    LogarithmicSpectrogramProcessor()                      # log10(1 * x + 1)
    spectrogram = processor.process(audio_path).T          # (92, n_frames)

This file restates madmom 0.15.1's published algorithm for exactly that chain (madmom/audio/signal.py: FramedSignal /
signal_frame; madmom/audio/stft.py; madmom/audio/filters.py: log_frequencies, frequencies2bins, TriangularFilter,
LogarithmicFilterbank; madmom/audio/spectrogram.py).  PARITY UNPINNED against madmom itself (not installable here, and
tutorials/audio.mp3 cannot be decoded); what pins the restatement is the reference's own notebook: the filterbank built
below has exactly the 92 bands cell 29 / 31 report ("92 Frequency Bins", excerpt shape (100, 1, 92, 42)).
"""
import numpy as np

SAMPLE_RATE, FRAME_SIZE, FPS = 22050, 2048, 20
NUM_BANDS, FMIN, FMAX, FREF = 16, 30.0, 6000.0, 440.0


def log_frequencies(bands_per_octave, fmin, fmax, fref=FREF):
    """madmom.audio.filters.log_frequencies"""
    left = np.floor(np.log2(float(fmin) / fref) * bands_per_octave)
    right = np.ceil(np.log2(float(fmax) / fref) * bands_per_octave)
    frequencies = fref * 2.0 ** (np.arange(left, right) / float(bands_per_octave))
    frequencies = frequencies[np.searchsorted(frequencies, fmin):]
    frequencies = frequencies[:np.searchsorted(frequencies, fmax, 'right')]
    return frequencies


def frequencies2bins(frequencies, bin_frequencies, unique_bins=False):
    """madmom.audio.filters.frequencies2bins: index of the closest FFT bin"""
    frequencies = np.asarray(frequencies)
    bin_frequencies = np.asarray(bin_frequencies)
    indices = bin_frequencies.searchsorted(frequencies)
    indices = np.clip(indices, 1, len(bin_frequencies) - 1)
    left = bin_frequencies[indices - 1]
    right = bin_frequencies[indices]
    indices -= frequencies - left < right - frequencies
    if unique_bins:
        indices = np.unique(indices)
    return indices


def triangular_filter(start, center, stop, norm):
    """madmom.audio.filters.TriangularFilter: rising [start, center), falling [center, stop), optionally unit area"""
    center, stop = center - start, stop - start
    data = np.zeros(stop)
    data[:center] = np.linspace(0, 1, center, endpoint=False)
    data[center:] = np.linspace(1, 0, stop - center, endpoint=False)
    if norm:
        data /= np.sum(data)
    return data


def log_filterbank(sample_rate=SAMPLE_RATE, frame_size=FRAME_SIZE, num_bands=NUM_BANDS, fmin=FMIN, fmax=FMAX):
    """LogarithmicFilterbank(bin_frequencies, num_bands, fmin, fmax, norm_filters=True, unique_filters=True)
    -> (frame_size / 2, n_bands) float32"""
    num_bins = frame_size >> 1
    bin_frequencies = np.fft.fftfreq(frame_size, 1.0 / sample_rate)[:num_bins]
    bins = frequencies2bins(log_frequencies(num_bands, fmin, fmax), bin_frequencies, unique_bins=True)
    if len(bins) < 3:
        raise ValueError("not enough bins to create a TriangularFilter")
    fb = np.zeros((num_bins, len(bins) - 2))
    for band in range(len(bins) - 2):
        start, center, stop = (int(b) for b in bins[band:band + 3])
        if stop - start < 2:                      # overlap=True: too small filters get the minimum width
            center, stop = start + 1, start + 2
        filt = triangular_filter(start, center, stop, norm=True)
        np.maximum(filt, fb[start:stop, band], out=fb[start:stop, band])      # Filterbank._put_filter
    return fb.astype(np.float32)


def num_frames(n_samples, sample_rate=SAMPLE_RATE, fps=FPS):
    """FramedSignal(end='normal'): ceil(len(signal) / hop_size), hop_size = sample_rate / fps (a float)"""
    return int(np.ceil(n_samples / (sample_rate / float(fps))))


def log_spectrogram(signal, sample_rate=SAMPLE_RATE, frame_size=FRAME_SIZE, fps=FPS, filterbank=None):
    """signal: mono float samples in [-1, 1) (madmom scales the window by 1 / 32768 for int16 audio, which is the same
    thing) -> (n_bands, n_frames) float32 = processor.process(path).T of cell 28.
    origin='future': frame i starts AT int(i * hop_size) (signal_frame: start = ref - frame_size // 2 - origin with
    origin = -(frame_size / 2)); samples beyond the end are zeros."""
    signal = np.asarray(signal, dtype=np.float64)
    fb = log_filterbank(sample_rate, frame_size) if filterbank is None else np.asarray(filterbank)
    hop = sample_rate / float(fps)
    n = num_frames(len(signal), sample_rate, fps)
    window = np.hanning(frame_size)
    frames = np.zeros((n, frame_size))
    for i in range(n):
        start = int(i * hop)
        seg = signal[start:start + frame_size]
        frames[i, :len(seg)] = seg
    spec = np.abs(np.fft.fft(frames * window, frame_size, axis=1)[:, :frame_size >> 1])
    return np.log10(np.dot(spec, fb.astype(np.float64)) + 1.0).T.astype(np.float32)
