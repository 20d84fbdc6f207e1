"""Host-side mirror of the reference's mel front-end / Griffin-Lim back-end (``data_utils.py:65-197``) on top of the C-ABI
(``avc_audio_*`` in include/avc_b200.h).  ``librosa.load`` / ``librosa.effects.trim`` (file I/O and silence trimming,
data_utils.py:93-96) stay with the caller; everything from the trimmed waveform to the normalised mel and back runs on the
GPU.  PyTorch supplies device memory and the stream only; there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from .engine import AvcError


def group_by_length(lengths: Sequence[int]) -> List[Tuple[int, List[int]]]:
    """Indices of equal-length items, in order of first appearance: [(length, [i, j, ...]), ...].  Utterances of one length
    form ONE batched call (the frames of the whole group are the rows of one GEMM per transform); ragged inputs therefore cost
    one call per distinct length, not one per utterance."""
    groups: Dict[int, List[int]] = {}
    for i, n in enumerate(lengths):
        groups.setdefault(int(n), []).append(i)
    return list(groups.items())


class AudioEngine:
    """One ``avc_audio_handle``: the preprocessing constants of ``config["preprocess"]`` (data_utils.py:214-220; the
    reference passes them as ``**config["preprocess"]`` to file2mel / mel2wav)."""

    def __init__(self, sample_rate: int = 24000, preemph: float = 0.97, n_fft: int = 2048, hop_length: int = 300,
                 win_length: int = 1200, n_mels: int = 80, ref_db: float = 20.0, max_db: float = 100.0,
                 top_db: Optional[float] = None, device: Optional[torch.device] = None):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise AvcError("attack_vc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        device = torch.device("cuda" if device is None else device)
        if device.type != "cuda":
            raise AvcError("attack_vc_b200 runs on CUDA only; there is no CPU fallback")
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        self.n_mels, self.hop_length, self.n_fft = int(n_mels), int(hop_length), int(n_fft)
        self.top_db = top_db          # used by librosa.effects.trim on the host, kept for signature compatibility
        d = _lib.AudioDesc(int(sample_rate), int(n_fft), int(hop_length), int(win_length), int(n_mels), float(preemph), float(ref_db), float(max_db))
        h = C.c_void_p()
        rc = self._lib.avc_audio_create(C.byref(h), C.byref(d), self.device.index)
        if rc != 0:
            msg = self._lib.avc_audio_last_error(None).decode()
            raise (ValueError if rc == -1 else AvcError)(f"avc_audio_create failed ({rc}): {msg}")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.avc_audio_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.avc_audio_last_error(self._h).decode()
            raise (ValueError if rc == -1 else AvcError)(f"libavc_b200: {msg}")

    def _vec(self, x: Tensor, name: str, dim: int) -> Tensor:
        if not isinstance(x, Tensor) or x.device != self.device:
            raise AvcError(f"{name} must be a tensor on {self.device} (no CPU fallback)")
        if x.dtype != torch.float32 or x.dim() != dim:
            raise ValueError(f"{name} must be float32 with {dim} dimension(s) (got {x.dtype} {tuple(x.shape)})")
        return x.contiguous()

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.avc_audio_kernel_launches(self._h))

    def wav2mel(self, wav: Tensor) -> Tensor:
        """file2mel from the trimmed waveform on (data_utils.py:99-114): wav [n] -> mel [n_frames, n_mels];
        a batch of equal-length waveforms [B, n] -> [B, n_frames, n_mels] (one GEMM over all frames)."""
        batched = isinstance(wav, Tensor) and wav.dim() == 2
        wav = self._vec(wav, "wav", 2 if batched else 1)
        B, n = (int(wav.shape[0]), int(wav.shape[1])) if batched else (1, int(wav.numel()))
        F = int(self._lib.avc_audio_frames(self._h, n))
        with torch.cuda.device(self.device):
            mel = torch.empty(B, max(F, 0), self.n_mels, device=self.device, dtype=torch.float32)
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self._check(self._lib.avc_audio_wav2mel_batch(self._h, wav.data_ptr(), B, n, mel.data_ptr(), st))
        return mel if batched else mel[0]

    def mel2wav(self, mel: Tensor, n_iter: int = 100) -> Tensor:
        """mel2wav (data_utils.py:120-165): mel [n_frames, n_mels] -> waveform [hop_length * (n_frames - 1)]."""
        batched = isinstance(mel, Tensor) and mel.dim() == 3          # [B, n_frames, n_mels]: equal-length utterances
        mel = self._vec(mel, "mel", 3 if batched else 2)
        if mel.shape[-1] != self.n_mels:
            raise ValueError(f"mel must have {self.n_mels} bins (got {mel.shape[-1]})")
        B, F = (int(mel.shape[0]), int(mel.shape[1])) if batched else (1, int(mel.shape[0]))
        n = int(self._lib.avc_audio_samples(self._h, F))
        with torch.cuda.device(self.device):
            wav = torch.empty(B, max(n, 0), device=self.device, dtype=torch.float32)
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self._check(self._lib.avc_audio_mel2wav_batch(self._h, mel.data_ptr(), B, F, int(n_iter), wav.data_ptr(), st))
        return wav if batched else wav[0]

    def wav2mel_list(self, wavs: Sequence[Tensor]) -> List[Tensor]:
        """file2mel over utterances of ANY lengths (the reference loops over files, data_utils.py:65-117): equal-length
        waveforms are batched into one call each, results come back in input order."""
        out: List[Optional[Tensor]] = [None] * len(wavs)
        for _, idx in group_by_length([int(w.numel()) for w in wavs]):
            mel = self.wav2mel(torch.stack([self._vec(wavs[i], "wav", 1) for i in idx]))
            for k, i in enumerate(idx):
                out[i] = mel[k]
        return out  # type: ignore[return-value]

    def mel2wav_list(self, mels: Sequence[Tensor], n_iter: int = 100) -> List[Tensor]:
        """mel2wav over mels of ANY frame counts: one batched Griffin-Lim per distinct length, input order kept."""
        out: List[Optional[Tensor]] = [None] * len(mels)
        for _, idx in group_by_length([int(m.shape[0]) for m in mels]):
            wav = self.mel2wav(torch.stack([self._vec(mels[i], "mel", 2) for i in idx]), n_iter=n_iter)
            for k, i in enumerate(idx):
                out[i] = wav[k]
        return out  # type: ignore[return-value]
