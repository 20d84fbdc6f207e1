"""Mel front-end / Griffin-Lim back-end (reference data_utils.py:65-197) on the B200 path against
oracle/audio_oracle.py (numpy; pinned against scipy.signal / transformers.audio_utils in tests/test_audio_oracle.py --
librosa itself is not installed: parity with the reference's third-party arithmetic is unpinned)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def tone(n, seed, sr=24000):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    w = 0.3 * np.sin(2 * np.pi * 220 * t) + 0.15 * np.sin(2 * np.pi * 1333 * t + 0.3) + 0.05 * np.sin(2 * np.pi * 5100 * t) + 0.02 * rng.standard_normal(n)
    return (w * np.hanning(n) ** 0.25).astype(np.float32)


@pytest.fixture(scope="module")
def audio():
    from attack_vc_b200.audio import AudioEngine
    return AudioEngine()


@pytest.mark.parametrize("n", [24000, 7001, 300 * 255 + 123])
def test_wav2mel(audio, n):
    from oracle import audio_oracle as A
    wav = tone(n, n)
    ref = A.wav2mel(wav.astype(np.float64))
    mel = audio.wav2mel(torch.from_numpy(wav).cuda())
    assert mel.shape == ref.shape == (1 + n // 300, 80)
    # values are (dB + 80) / 100 in [1e-8, 1]; the dB of a near-silent bin is ill-conditioned in its magnitude (fp32 on both
    # sides of the log), so: 1e-4 absolute (= 0.01 dB) on every bin, 1e-5 on average
    err = (mel.cpu().numpy().astype(np.float64) - ref).__abs__()
    assert err.max() < 1e-4 and err.mean() < 1e-5, (err.max(), err.mean())
    assert float(mel.min()) >= float(np.float32(1e-8)) and float(mel.max()) <= 1.0


@pytest.mark.parametrize("n_fft,hop,win,n_mels,sr", [(1024, 256, 1024, 80, 16000), (512, 128, 400, 40, 16000)])
def test_wav2mel_other_parameters(n_fft, hop, win, n_mels, sr):
    from attack_vc_b200.audio import AudioEngine
    from oracle import audio_oracle as A
    eng = AudioEngine(sample_rate=sr, n_fft=n_fft, hop_length=hop, win_length=win, n_mels=n_mels)
    wav = tone(sr // 2, 3, sr)
    ref = A.wav2mel(wav.astype(np.float64), sample_rate=sr, n_fft=n_fft, hop_length=hop, win_length=win, n_mels=n_mels)
    mel = eng.wav2mel(torch.from_numpy(wav).cuda())
    assert mel.shape == ref.shape
    assert np.abs(mel.cpu().numpy() - ref).max() < 1e-4
    eng.close()


@pytest.mark.parametrize("n_iter", [0, 3])
def test_mel2wav_few_iterations(audio, n_iter):
    """mel -> waveform with 0 / 3 Griffin-Lim iterations against the oracle, sample by sample (relative to the peak)."""
    from oracle import audio_oracle as A
    mel = A.wav2mel(tone(300 * 120, 9).astype(np.float64))
    ref = A.mel2wav(mel, n_iter=n_iter).astype(np.float64)
    wav = audio.mel2wav(torch.from_numpy(mel).cuda(), n_iter=n_iter).cpu().numpy().astype(np.float64)
    assert wav.shape == ref.shape == (300 * (mel.shape[0] - 1),)
    # n_iter = 0 is linear in the mel: tight.  Every Griffin-Lim iteration divides by |est|: a bin whose estimate is near zero
    # gets an ill-defined phase, so fp32 (here) and fp64 (oracle) drift apart by a percent of the peak within three iterations
    tol_max, tol_l2 = (2e-3, 5e-4) if n_iter == 0 else (5e-2, 2e-2)
    assert np.abs(wav - ref).max() < tol_max * np.abs(ref).max(), np.abs(wav - ref).max() / np.abs(ref).max()
    assert np.linalg.norm(wav - ref) < tol_l2 * np.linalg.norm(ref)


def test_griffin_lim_100_iterations_converges_like_the_oracle(audio):
    """The reference's 100 iterations (data_utils.py:173): phase retrieval amplifies rounding differences between two correct
    implementations, so the waveforms are compared through what the algorithm minimises -- the spectral inconsistency
    || |STFT(y)| - mag || / || mag || of the result -- and through their correlation."""
    from oracle import audio_oracle as A
    mel = A.wav2mel(tone(300 * 100, 21).astype(np.float64))
    ref = A.mel2wav(mel, n_iter=100).astype(np.float64)
    wav = audio.mel2wav(torch.from_numpy(mel).cuda(), n_iter=100).cpu().numpy().astype(np.float64)
    lin = np.power(10.0, ((np.clip(mel.T, 0, 1) * 100.0) - 100.0 + 20.0) * 0.05)
    mag = A.inv_mel_matrix(24000, 2048, 80) @ lin
    pre = lambda y: np.append(y[0], y[1:] - 0.97 * y[:-1])          # undo the de-emphasis to get back to Griffin-Lim's output
    def inconsistency(y):
        return np.linalg.norm(np.abs(A.stft(pre(y), 2048, 300, 1200)) - mag) / np.linalg.norm(mag)
    e_ref, e_gpu = inconsistency(ref), inconsistency(wav)
    assert e_gpu < 1.05 * e_ref + 1e-3, (e_gpu, e_ref)
    corr = float(np.dot(wav, ref) / (np.linalg.norm(wav) * np.linalg.norm(ref)))
    assert corr > 0.98, corr


def test_round_trip_and_errors(audio):
    from attack_vc_b200 import AvcError
    wav = torch.from_numpy(tone(24000, 1)).cuda()
    mel = audio.wav2mel(wav)
    back = audio.mel2wav(mel, n_iter=8)
    assert back.shape == (300 * (mel.shape[0] - 1),) and bool(torch.isfinite(back).all())
    mel2 = audio.wav2mel(back)
    n = min(mel.shape[0], mel2.shape[0])
    assert float((mel2[:n] - mel[:n]).abs().mean()) < 0.05        # the re-analysed mel stays close (80 bins lose information; GL is approximate)
    with pytest.raises(AvcError):
        audio.wav2mel(torch.zeros(24000))                          # CPU tensor: no fallback
    with pytest.raises(ValueError):
        audio.wav2mel(torch.zeros(500, device="cuda"))             # shorter than the reflect padding
    with pytest.raises(ValueError):
        audio.mel2wav(torch.zeros(10, 40, device="cuda"))          # wrong number of mel bins


def test_batch_equals_single_calls(audio):
    """B equal-length utterances as the rows of one GEMM per transform == B single calls (bit for bit for wav2mel: the same
    per-row arithmetic; mel2wav within the K-split / tile-shape differences of the GEMM)."""
    wavs = torch.stack([torch.from_numpy(tone(300 * 60 + 11, s)) for s in (1, 2, 3)]).cuda()
    mel_b = audio.wav2mel(wavs)
    assert mel_b.shape == (3, 61, 80)
    for b in range(3):
        assert float((mel_b[b] - audio.wav2mel(wavs[b])).abs().max()) < 2e-6
    back_b = audio.mel2wav(mel_b, n_iter=2)
    assert back_b.shape == (3, 300 * 60)
    for b in range(3):
        one = audio.mel2wav(mel_b[b].contiguous(), n_iter=2)
        assert float((back_b[b] - one).abs().max()) < 1e-3 * float(one.abs().max())
