"""Constant-Q front end: ``CQT``, ``PhaseDifference``, ``PreprocessingModule``.

Same constructor arguments, attributes and state_dict keys as the reference
(constant_q_transform.py:94-172, 268-286; scalogram_model.py:34-102); the forward passes run the fused
sm_100a kernel behind ``cpc_cqt_fwd`` instead of 9 conv1d + ~20 elementwise ATen launches.

The filterbank itself (a one-off at construction time) is built on the host in float64 exactly as
``librosa.filters.constant_q`` (librosa <= 0.7, which the reference calls at :108-112) defines it.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops

pi = np.pi


# ---- host-side filterbank construction (init time only) -------------------------------------------

def cqt_frequencies(n_bins, fmin, bins_per_octave=12):
    return float(fmin) * 2.0 ** (np.arange(n_bins, dtype=float) / bins_per_octave)


def constant_q_filterbank(sr, fmin, n_bins, bins_per_octave, filter_scale):
    """Hann-windowed, L1-normalised complex exponentials, zero-padded (centred) to a power of two.
    Operation order follows librosa's published formula so the fp32 weights are reproducible bit for bit."""
    q = float(filter_scale) / (2.0 ** (1.0 / bins_per_octave) - 1.0)
    lengths = q * sr / cqt_frequencies(n_bins, fmin, bins_per_octave)
    freqs = q * sr / lengths                    # librosa converts the lengths back to frequencies
    if freqs[-1] * (1.0 + 0.5 * 1.50018310546875 / q) > sr / 2.0:
        raise ValueError("highest CQT filter exceeds Nyquist")
    width = int(2.0 ** np.ceil(np.log2(lengths.max())))
    bank = np.zeros((n_bins, width), dtype=np.complex128)
    for row in range(n_bins):
        length = lengths[row]
        tone = np.exp(np.arange(-length // 2, length // 2, dtype=float) * 1j * 2 * np.pi * freqs[row] / sr)
        count = len(tone)
        tone = tone * (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(count) / count))    # periodic Hann
        tone = tone / np.sum(np.abs(tone))
        start = (width - count) // 2
        bank[row, start:start + count] = tone
    return bank, lengths


def octave_groups(lengths):
    """One conv per power-of-two kernel size: a new group starts when the size drops."""
    sizes, bounds = [], []
    for idx, length in enumerate(lengths):
        size = 1 << int(math.ceil(np.log2(length)))
        if sizes and size >= sizes[-1]:
            continue
        if sizes:
            bounds.append(idx)
        sizes.append(size)
    edges = [0] + bounds + [len(lengths)]
    return sizes, [range(edges[i], edges[i + 1]) for i in range(len(sizes))]


# ---- elementwise helpers kept for API compatibility (constant_q_transform.py:36-72) ----------------

def abs(z, complex_dim=None):
    d = z.dim() - 1 if complex_dim is None else complex_dim
    re, im = z.select(d, 0), z.select(d, 1)
    return torch.sqrt(re ** 2 + im ** 2)


def angle(z, complex_dim=None):
    d = z.dim() - 1 if complex_dim is None else complex_dim
    return torch.atan2(z.select(d, 1), z.select(d, 0))


def unwrap(x):
    x = torch.where(x > pi, x - 2 * pi, x)
    return torch.where(x < -pi, x + 2 * pi, x)


def to_complex(real, imag, complex_dim=None):
    return torch.stack([real, imag], dim=real.dim() if complex_dim is None else complex_dim)


def polar_to_complex(magnitude, phase, complex_dim=None):
    return to_complex(magnitude * torch.cos(phase), magnitude * torch.sin(phase), complex_dim)


class CQT(nn.Module):
    """Drop-in for the reference ``CQT`` (constant_q_transform.py:94-172)."""

    def __init__(self, sr=16000, fmin=30, n_bins=256, bins_per_octave=32, filter_scale=1., hop_length=128,
                 trainable=False):
        super().__init__()
        self.sr, self.fmin, self.n_bins = sr, fmin, n_bins
        self.bins_per_octave, self.filter_scale, self.hop_length = bins_per_octave, filter_scale, hop_length
        bank, lengths = constant_q_filterbank(sr, fmin, n_bins, bins_per_octave, filter_scale)
        self.cqt_filter_lengths = lengths
        self.conv_kernel_sizes, self.conv_index_ranges = octave_groups(lengths)
        width = bank.shape[-1]
        self.conv_modules = nn.ModuleList()
        for size, rng in zip(self.conv_kernel_sizes, self.conv_index_ranges):
            crop = (width - size) // 2
            part = bank[rng.start:rng.stop, crop:width - crop]
            weight = torch.from_numpy(np.concatenate([part.real, part.imag], axis=0)).float().unsqueeze(1)
            conv = nn.Conv1d(1, weight.shape[0], size, stride=hop_length, bias=False)
            conv.weight = nn.Parameter(weight, requires_grad=False)
            self.conv_modules.append(conv)
        self._trainable = False
        self.trainable = trainable
        self._packed = None
        self._packed_key = None
        self._tensor_filters = None

    @property
    def trainable(self):
        return self._trainable

    @trainable.setter
    def trainable(self, value):
        for p in self.parameters():
            p.requires_grad = value
        self._trainable = value

    def __getstate__(self):
        """Whole-model pickles (the reference's snapshot format) and deepcopies carry the filterbank once, as the conv
        weights: the packed device-side copies are caches."""
        state = dict(self.__dict__)
        state["_packed"], state["_packed_key"], state["_tensor_filters"] = None, None, None
        return state

    def kernel_plan(self):
        offsets, total = [], 0
        for conv in self.conv_modules:
            offsets.append(total)
            total += conv.weight.numel()
        return {"kernel_sizes": list(self.conv_kernel_sizes),
                "ranges": [(r.start, r.stop) for r in self.conv_index_ranges],
                "weight_offsets": offsets, "hop": self.hop_length, "n_bins": self.n_bins}

    def packed_weights(self):
        """All group weights in one device buffer (rebuilt when a conv weight is replaced or modified)."""
        key = tuple((c.weight.data_ptr(), c.weight._version, c.weight.device) for c in self.conv_modules)
        if self._packed is None or key != self._packed_key:
            with torch.no_grad():
                self._packed = torch.cat([c.weight.detach().reshape(-1) for c in self.conv_modules]).contiguous()
            self._packed_key = key
            self._tensor_filters = None
        return self._packed

    def tensor_core_filters(self):
        """The filterbank as the tensor-core kernel reads it (``ops.cqt_pack_filters``), cached with the weights."""
        weights = self.packed_weights()
        if getattr(self, "_tensor_filters", None) is None:
            self._tensor_filters = (ops.cqt_pack_filters(weights, self.kernel_plan()),)
        return self._tensor_filters[0]

    def needs_autograd(self, x):
        """True when a gradient must flow through the transform: a trainable filterbank (``trainable_cqt``) or an
        input that requires grad (dreaming: audio-input optimisation, dreaming/dreaming_calculation.py:169,265)."""
        return torch.is_grad_enabled() and (self._trainable or x.requires_grad)

    def forward_differentiable(self, x):
        """constant_q_transform.py:161-172 literally, every group through the conv kernels (forward, dgrad towards
        the audio, wgrad towards the filters; second derivatives by composition).  The fused single-launch kernel is
        forward-only and serves the frozen transform."""
        real, imag = [], []
        for size, conv in zip(self.conv_kernel_sizes, self.conv_modules):
            offset = (self.conv_kernel_sizes[0] - size) // 2
            y = ops.conv1d(x[:, :, offset:x.shape[2] - (offset + 1)], conv.weight, None, stride=self.hop_length)
            r, i = torch.chunk(y, 2, dim=1)
            real.append(r)
            imag.append(i)
        return torch.stack([torch.cat(real, dim=1), torch.cat(imag, dim=1)], dim=3)

    def forward(self, x):
        if self.needs_autograd(x):
            return self.forward_differentiable(x)
        return ops.cqt_frontend(x, self.packed_weights(), self.kernel_plan(), _lib.CQT_COMPLEX,
                                packed_filters=self.tensor_core_filters())


class InverseCQT(nn.Module):
    """constant_q_transform.py:180-260: complex CQT coefficients (N, n_bins, T, 2) -> complex signal (N, L, 2), one
    ConvTranspose1d per octave group on the conv data-gradient kernels.  Same attributes and state_dict keys
    (``conv_modules.{i}.weight`` of shape (n_g, 2, K_g)).  (The reference's forward applies ``.view`` to a permuted
    tensor at :253, which raises in current PyTorch; the intended reshape is what runs here.)"""

    def __init__(self, sr=16000, fmin=30, n_bins=256, bins_per_octave=32, filter_scale=1., hop_length=128,
                 trainable=False):
        super().__init__()
        self.sr, self.fmin, self.n_bins = sr, fmin, n_bins
        self.bins_per_octave, self.filter_scale, self.hop_length = bins_per_octave, filter_scale, hop_length
        bank, lengths = constant_q_filterbank(sr, fmin, n_bins, bins_per_octave, filter_scale)
        self.cqt_filter_lengths = lengths
        self.conv_kernel_sizes, self.conv_index_ranges = octave_groups(lengths)
        width = bank.shape[-1]
        self.conv_modules = nn.ModuleList()
        for size, rng in zip(self.conv_kernel_sizes, self.conv_index_ranges):
            crop = (width - size) // 2
            part = bank[rng.start:rng.stop, crop:width - crop]
            weight = torch.stack([torch.from_numpy(part.real.copy()), torch.from_numpy(part.imag.copy())], dim=1).float()
            conv = nn.ConvTranspose1d(weight.shape[0], 2, size, bias=False, stride=hop_length, padding=size // 2)
            conv.weight = nn.Parameter(weight, requires_grad=False)
            self.conv_modules.append(conv)
        self._trainable = False
        self.trainable = trainable

    @property
    def trainable(self):
        return self._trainable

    @trainable.setter
    def trainable(self, value):
        for p in self.parameters():
            p.requires_grad = value
        self._trainable = value

    def forward(self, x):
        result = 0
        n = x.shape[0]
        for size, rng, conv in zip(self.conv_kernel_sizes, self.conv_index_ranges, self.conv_modules):
            band = x[:, rng.start:rng.stop]                                      # (N, n_g, T, 2)
            p, t = band.shape[1], band.shape[2]
            band = band.permute(3, 0, 1, 2).reshape(2 * n, p, t)                 # real parts first, then imaginary
            result = result + ops.conv_transpose1d(band, conv.weight, stride=self.hop_length, padding=size // 2)
        result = result.view(2, n, 2, -1)
        return torch.stack([result[0, :, 0] - result[1, :, 0], result[0, :, 1] + result[1, :, 1]], dim=2)


class PhaseAccumulation(nn.Module):
    """constant_q_transform.py:294-313: inverse of PhaseDifference (cumulative phase from scaled differences).
    ``start_phase`` is (1, n_bins, 1) as in the reference, so it concatenates with batch-1 inputs only (:311)."""

    def __init__(self, sr=16000, fmin=30, n_bins=256, bins_per_octave=32, hop_length=128):
        super().__init__()
        freqs = cqt_frequencies(n_bins, fmin, bins_per_octave)
        fixed = (((1.0 * freqs * hop_length / sr) + 0.5) % 1 - 0.5) * 2 * np.pi
        # a plain attribute in the reference (not in its state_dict): non-persistent buffer here so .to() moves it
        self.register_buffer("fixed_phase_diff", torch.from_numpy(fixed).float().view(1, -1, 1), persistent=False)
        scaling = torch.from_numpy(1 / np.log(freqs)).float().view(1, -1, 1)
        self.scaling = nn.Parameter(scaling, requires_grad=False)
        self.start_phase = nn.Parameter(torch.zeros_like(scaling), requires_grad=False)

    def forward(self, x):
        x = (x / self.scaling) - self.fixed_phase_diff
        x = torch.cat([self.start_phase, x], dim=2)
        x = torch.cumsum(x, dim=2)
        return x % (2 * np.pi) - np.pi


class PhaseDifference(nn.Module):
    """constant_q_transform.py:268-286; its arithmetic is fused into the front-end kernel, the module keeps
    the two per-bin constant vectors (same parameter names as the reference)."""

    def __init__(self, sr=16000, fmin=30, n_bins=256, bins_per_octave=32, hop_length=128):
        super().__init__()
        freqs = cqt_frequencies(n_bins, fmin, bins_per_octave)
        fixed = (((1.0 * freqs * hop_length / sr) + 0.5) % 1 - 0.5) * 2 * np.pi
        self.fixed_phase_diff = nn.Parameter(torch.from_numpy(fixed).float().view(1, -1, 1), requires_grad=False)
        self.scaling = nn.Parameter(torch.from_numpy(1 / np.log(freqs)).float().view(1, -1, 1), requires_grad=False)

    def forward(self, x):
        return unwrap(x[:, :, 1:] - x[:, :, :-1] + self.fixed_phase_diff) * self.scaling


class PreprocessingModule(nn.Module):
    """Drop-in for scalogram_model.py:34-102: audio (B,1,L) -> scalogram (B,C,n_bins,T')."""

    def __init__(self, cqt_dict=None, phase=False, output_requires_grad=False, offset_zero=False, output_power=1.,
                 pooling=None, scaling=1.):
        super().__init__()
        self.downsampling_factor = 1
        self.receptive_field = 1
        self.cqt = None
        if cqt_dict is not None:
            self.cqt = CQT(sr=cqt_dict['sample_rate'], fmin=cqt_dict['fmin'], n_bins=cqt_dict['n_bins'],
                           bins_per_octave=cqt_dict['bins_per_octave'], filter_scale=cqt_dict['filter_scale'],
                           hop_length=cqt_dict['hop_length'], trainable=cqt_dict['trainable_cqt'])
            self.downsampling_factor = cqt_dict['hop_length']
            self.receptive_field = self.cqt.conv_kernel_sizes[0]
        self.phase_diff = None
        if phase:
            self.phase_diff = PhaseDifference(sr=cqt_dict['sample_rate'], fmin=cqt_dict['fmin'],
                                              n_bins=cqt_dict['n_bins'], bins_per_octave=cqt_dict['bins_per_octave'],
                                              hop_length=cqt_dict['hop_length'])
        self.output_power = output_power
        if offset_zero:
            self.offset = 1e-9
            self.log_offset = -math.log(self.offset)
            self.normalization_factor = scaling / self.log_offset
        else:
            self.offset = 0
            self.log_offset = 0
            self.normalization_factor = scaling
        self.pooling = pooling
        if pooling is not None:
            if list(pooling) not in ([1, 1], [1, 2]):
                raise NotImplementedError("scalogram pooling %s: only [1, 2] (time pairs) is used by the configs "
                                          "and implemented in the fused kernel" % (pooling,))
            self.downsampling_factor *= pooling[1]
        self.output = None

    def forward(self, x):
        if self.cqt is None:
            return x
        if self.cqt.needs_autograd(x):
            return self._forward_differentiable(x)
        pool_t = 1 if self.pooling is None else int(self.pooling[1])
        if self.phase_diff is not None:
            y = ops.cqt_frontend(x, self.cqt.packed_weights(), self.cqt.kernel_plan(), _lib.CQT_LOGPOW_PHASE,
                                 phase_fixed=self.phase_diff.fixed_phase_diff.reshape(-1),
                                 phase_scale=self.phase_diff.scaling.reshape(-1), pool_t=pool_t, eps=self.offset,
                                 log_offset=self.log_offset, norm=self.normalization_factor, power=self.output_power,
                                 packed_filters=self.cqt.tensor_core_filters())
        else:
            y = ops.cqt_frontend(x, self.cqt.packed_weights(), self.cqt.kernel_plan(), _lib.CQT_LOGPOW, pool_t=pool_t,
                                 eps=self.offset, log_offset=self.log_offset, norm=self.normalization_factor,
                                 power=self.output_power, packed_filters=self.cqt.tensor_core_filters())
        self.output = y
        return y

    def _forward_differentiable(self, x):
        """scalogram_model.py:75-102 term by term on top of the differentiable CQT (autograd needs the complex
        coefficients); same values as the fused kernel."""
        z = self.cqt.forward_differentiable(x)
        if self.phase_diff is not None:
            amp = torch.log(torch.pow(abs(z[:, :, 1:]), 2) + self.offset) + self.log_offset
            y = torch.stack([amp, self.phase_diff(angle(z))], dim=1)
        else:
            y = torch.log(torch.pow(abs(z), 2) + self.offset).unsqueeze(1) + self.log_offset
        if self.pooling is not None:
            y = F.max_pool2d(y, self.pooling)
        y = (y * self.normalization_factor) ** self.output_power
        self.output = y
        return y
