"""oracle/audio_oracle.py (numpy restatement of the librosa calls of data_utils.py:65-197; librosa is not installed) pinned
against two independent installed implementations of the same published algorithms -- scipy.signal.stft / istft / lfilter and
transformers.audio_utils.mel_filter_bank -- and against analytic identities."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def A():
    from oracle import audio_oracle
    return audio_oracle


@pytest.mark.parametrize("n_fft,hop,win", [(2048, 300, 1200), (1024, 256, 1024), (512, 128, 400)])
def test_stft_matches_scipy(A, n_fft, hop, win):
    import scipy.signal as ss
    rng = np.random.default_rng(n_fft)
    y = rng.standard_normal(hop * 40 + 17)
    S = A.stft(y, n_fft, hop, win)
    # scipy: same frames when the signal is reflect-padded by hand, the window is the padded periodic Hann and scaling is off
    yp = np.pad(y, n_fft // 2, mode="reflect")
    _, _, Z = ss.stft(yp, window=A.hann_padded(win, n_fft), nperseg=n_fft, noverlap=n_fft - hop, nfft=n_fft, boundary=None,
                      padded=False, return_onesided=True, scaling="spectrum")
    Z = Z * A.hann_padded(win, n_fft).sum()          # undo scipy's 1 / sum(window) spectrum scaling
    assert S.shape == Z.shape == (1 + n_fft // 2, 1 + (len(yp) - n_fft) // hop)
    assert np.abs(S - Z).max() <= 1e-9 * np.abs(Z).max()


@pytest.mark.parametrize("n_fft,hop,win", [(2048, 300, 1200), (512, 128, 400)])
def test_istft_inverts_stft(A, n_fft, hop, win):
    rng = np.random.default_rng(3)
    y = rng.standard_normal(hop * 30)
    back = A.istft(A.stft(y, n_fft, hop, win), hop, win)
    n = min(len(y), len(back))
    assert len(back) == hop * (A.stft(y, n_fft, hop, win).shape[1] - 1)
    assert np.abs(back[:n] - y[:n]).max() < 1e-9     # windowed overlap-add / window sum-square is an exact inverse (NOLA holds)


def test_mel_basis_matches_transformers(A):
    from transformers.audio_utils import mel_filter_bank
    for sr, n_fft, n_mels in ((24000, 2048, 80), (16000, 1024, 80), (24000, 2048, 512)):
        want = mel_filter_bank(num_frequency_bins=1 + n_fft // 2, num_mel_filters=n_mels, min_frequency=0.0, max_frequency=sr / 2.0,
                               sampling_rate=sr, norm="slaney", mel_scale="slaney").T
        got = A.mel_basis(sr, n_fft, n_mels)
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.abs(got - want).max() <= 2e-7 * np.abs(want).max() + 1e-12


def test_deemphasis_matches_scipy_lfilter(A):
    from scipy.signal import lfilter
    x = np.random.default_rng(5).standard_normal(5000)
    assert np.allclose(A.lfilter_deemph(x, 0.97), lfilter([1], [1, -0.97], x), rtol=0, atol=1e-12)
    pre = np.append(x[0], x[1:] - 0.97 * x[:-1])                     # data_utils.py:99 is its exact inverse
    assert np.allclose(A.lfilter_deemph(pre, 0.97), x, atol=1e-9)


def test_wav2mel_shape_range_and_griffin_lim_consistency(A):
    rng = np.random.default_rng(7)
    t = np.arange(24000 // 2) / 24000.0
    wav = 0.3 * np.sin(2 * np.pi * 220 * t) + 0.1 * np.sin(2 * np.pi * 1333 * t) + 0.01 * rng.standard_normal(len(t))
    mel = A.wav2mel(wav)
    assert mel.shape == (1 + len(wav) // 300, 80) and mel.dtype == np.float32
    assert mel.min() >= 1e-8 and mel.max() <= 1.0
    # Griffin-Lim lowers the spectral inconsistency || |stft(istft(X))| - |X| || monotonically in practice: a few iterations suffice here
    mag = np.abs(A.stft(wav, 2048, 300, 1200))
    def err(n):
        y = A.griffin_lim(mag, 300, 1200, 2048, n)
        return np.linalg.norm(np.abs(A.stft(y, 2048, 300, 1200)) - mag) / np.linalg.norm(mag)
    e0, e5 = err(0), err(5)
    assert e5 < 0.6 * e0
    assert A.inv_mel_matrix(24000, 2048, 80).shape == (1025, 80)


def test_ragged_list_helpers_group_equal_lengths_and_keep_order():
    """AudioEngine.wav2mel_list / mel2wav_list (host logic only, no GPU): utterances of one length go into ONE batched call,
    results come back in input order.  The engine is a stub whose batched transforms are recognisable functions of the input."""
    import torch
    from attack_vc_b200.audio import AudioEngine, group_by_length
    assert group_by_length([5, 3, 5, 7, 3, 5]) == [(5, [0, 2, 5]), (3, [1, 4]), (7, [3])]
    assert group_by_length([]) == []
    eng = AudioEngine.__new__(AudioEngine)
    eng.device = torch.device("cpu")
    eng.n_mels = 4
    eng._h = None
    calls = []

    def fake_wav2mel(wav):
        calls.append(("w", tuple(wav.shape)))
        return wav.sum(dim=1)[:, None, None].expand(-1, 2, 4).clone()

    def fake_mel2wav(mel, n_iter=100):
        calls.append(("m", tuple(mel.shape), n_iter))
        return mel.sum(dim=(1, 2))[:, None].expand(-1, 3).clone()
    eng.wav2mel, eng.mel2wav = fake_wav2mel, fake_mel2wav
    wavs = [torch.full((n,), float(i + 1)) for i, n in enumerate([5, 3, 5, 7, 3])]
    mels = eng.wav2mel_list(wavs)
    assert [c for c in calls if c[0] == "w"] == [("w", (2, 5)), ("w", (2, 3)), ("w", (1, 7))]
    for i, w in enumerate(wavs):
        assert mels[i].shape == (2, 4) and float(mels[i][0, 0]) == float(w.sum())
    back = eng.mel2wav_list([torch.full((f, 4), float(i + 1)) for i, f in enumerate([6, 2, 6])], n_iter=7)
    assert [c for c in calls if c[0] == "m"] == [("m", (2, 6, 4), 7), ("m", (1, 2, 4), 7)]
    assert [float(b[0]) for b in back] == [24.0, 16.0, 72.0]
