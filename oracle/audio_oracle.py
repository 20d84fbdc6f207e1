"""CPU oracle for the mel front-end / Griffin-Lim back-end.  TEST INFRASTRUCTURE ONLY.

Restates the arithmetic of ``/root/reference/data_utils.py``:
  inv_mel_matrix :16-31     m = librosa.filters.mel(sr, n_fft, n_mels); pinv-like m.T diag(1 / sum(m m.T))
  file2mel       :65-117    (after librosa.load / effects.trim, which stay host I/O) pre-emphasis, |STFT|, mel basis,
                            20 log10(max(1e-5, .)), clip((mel - ref_db + max_db) / max_db, 1e-8, 1), transpose
  mel2wav        :120-165   inverse of the dB scaling, inv_mel_matrix, griffin_lim, lfilter([1], [1, -preemph])
  griffin_lim    :168-197   n_iter x (istft -> stft -> unit phase x magnitude), final istft

**Parity unpinned against librosa.**  The algorithms live in librosa (absent from /root/reference and not installed here; the
reference pins no version, and calls ``librosa.stft(y, n_fft, hop_length, win_length)`` / ``librosa.filters.mel(sr, n_fft,
n_mels)`` positionally, i.e. the <= 0.9 API: center=True, pad_mode="reflect", periodic Hann window zero-padded to n_fft,
Slaney mel scale with Slaney normalisation).  This file restates those published algorithms in numpy; tests pin it against two
independent implementations that ARE installed -- ``scipy.signal.stft / istft`` (same framing when configured alike) and
``transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney")`` (written to reproduce librosa.filters.mel) --
and against analytic properties (perfect reconstruction istft(stft(x)) = x, Parseval).  The reference ships no audio
fixtures.  Only ``tests/`` and bench legs may import this file.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

# AdaIN-VC's preprocessing constants (the reference reads them from an external config.yaml, data_utils.py:214-220);
# n_mels follows BASELINE.json's 80-bin mels
DEFAULT = dict(sample_rate=24000, preemph=0.97, n_fft=2048, hop_length=300, win_length=1200, n_mels=80, ref_db=20.0, max_db=100.0)


def hann_padded(win_length: int, n_fft: int) -> np.ndarray:
    """scipy.signal.get_window('hann', win_length, fftbins=True) centred in n_fft samples (librosa.util.pad_center)."""
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)
    lpad = (n_fft - win_length) // 2
    out = np.zeros(n_fft)
    out[lpad:lpad + win_length] = w
    return out


def stft(y: np.ndarray, n_fft: int, hop_length: int, win_length: int) -> np.ndarray:
    """librosa.stft(y, n_fft, hop_length, win_length): center=True (reflect padding of n_fft // 2), -> [1 + n_fft/2, n_frames]."""
    w = hann_padded(win_length, n_fft)
    yp = np.pad(np.asarray(y, dtype=np.float64), n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(n_frames)[:, None]
    return np.fft.rfft(yp[idx] * w[None, :], axis=1).T


def istft(S: np.ndarray, hop_length: int, win_length: int) -> np.ndarray:
    """librosa.istft(S, hop_length, win_length, window="hann"): windowed overlap-add, divided by the window sum-square where
    it exceeds tiny(float32), the n_fft // 2 centre padding removed."""
    n_fft = 2 * (S.shape[0] - 1)
    w = hann_padded(win_length, n_fft)
    n_frames = S.shape[1]
    frames = np.fft.irfft(S.T, n=n_fft, axis=1) * w[None, :]
    n = n_fft + hop_length * (n_frames - 1)
    y, wss = np.zeros(n), np.zeros(n)
    for i in range(n_frames):
        y[i * hop_length:i * hop_length + n_fft] += frames[i]
        wss[i * hop_length:i * hop_length + n_fft] += w * w
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2: n - n_fft // 2]


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    mel = f / (200.0 / 3)
    log = f >= 1000.0
    return np.where(log, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), mel)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), (200.0 / 3) * m)


def mel_basis(sample_rate: int, n_fft: int, n_mels: int) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels): fmin 0, fmax sr/2, Slaney scale, triangles normalised to unit area -> [n_mels, 1 + n_fft/2]."""
    fft_f = np.linspace(0.0, sample_rate / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sample_rate / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    w = np.zeros((n_mels, len(fft_f)))
    for i in range(n_mels):
        w[i] = np.maximum(0.0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)          # librosa returns float32


def inv_mel_matrix(sample_rate: int, n_fft: int, n_mels: int) -> np.ndarray:
    """data_utils.py:16-31."""
    m = mel_basis(sample_rate, n_fft, n_mels)
    p = np.matmul(m, m.T)
    d = [1.0 / x if np.abs(x) > 1e-8 else x for x in np.sum(p, axis=0)]
    return np.matmul(m.T, np.diag(d))


def wav2mel(wav: np.ndarray, sample_rate=24000, preemph=0.97, n_fft=2048, hop_length=300, win_length=1200, n_mels=80,
            ref_db=20.0, max_db=100.0) -> np.ndarray:
    """data_utils.file2mel :99-117 from the trimmed waveform on -> [n_frames, n_mels] float32."""
    wav = np.append(wav[0], wav[1:] - preemph * wav[:-1])                                  # :99
    mag = np.abs(stft(wav, n_fft, hop_length, win_length))                                 # :102-105
    mel = np.dot(mel_basis(sample_rate, n_fft, n_mels), mag)                               # :108-109
    mel = 20 * np.log10(np.maximum(1e-5, mel))                                             # :112
    mel = np.clip((mel - ref_db + max_db) / max_db, 1e-8, 1)                               # :113
    return mel.T.astype(np.float32)                                                        # :114


def griffin_lim(spect: np.ndarray, hop_length: int, win_length: int, n_fft: int, n_iter: Optional[int] = 100) -> np.ndarray:
    """data_utils.py:168-197."""
    X_best = spect.astype(np.complex128)
    for _ in range(n_iter):
        X_t = istft(X_best, hop_length, win_length)
        est = stft(X_t, n_fft, hop_length, win_length)
        phase = est / np.maximum(1e-8, np.abs(est))
        X_best = spect * phase
    return np.real(istft(X_best, hop_length, win_length))


def lfilter_deemph(wav: np.ndarray, preemph: float) -> np.ndarray:
    """scipy.signal.lfilter([1], [1, -preemph], wav): y[n] = x[n] + preemph * y[n-1]."""
    y = np.empty(len(wav))
    acc = 0.0
    for i, v in enumerate(wav):
        acc = v + preemph * acc
        y[i] = acc
    return y


def mel2wav(mel: np.ndarray, sample_rate=24000, preemph=0.97, n_fft=2048, hop_length=300, win_length=1200, n_mels=80,
            ref_db=20.0, max_db=100.0, n_iter: int = 100) -> np.ndarray:
    """data_utils.mel2wav :149-165; mel [n_frames, n_mels] -> waveform float32."""
    mel = mel.T                                                                            # :150
    mel = (np.clip(mel, 0, 1) * max_db) - max_db + ref_db                                  # :151
    mel = np.power(10.0, mel * 0.05)                                                       # :152
    mag = np.dot(inv_mel_matrix(sample_rate, n_fft, n_mels), mel)                          # :155-156
    wav = griffin_lim(mag, hop_length, win_length, n_fft, n_iter)                          # :159
    return lfilter_deemph(wav, preemph).astype(np.float32)                                 # :162-164
