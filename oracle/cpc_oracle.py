"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A CPU restatement (numpy + torch-CPU fp32/fp64) of the reference's CPC training hot path
(SURVEY.md section 8a).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this file; the
product package (``cpc_b200``) never does and fails loudly without its CUDA library.

Every function cites the reference lines it restates (paths relative to the reference
root).  The restatement is pinned against the *unmodified* reference executed on CPU in
the build container: ``oracle/make_golden.py`` freezes the reference's outputs into
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this file against them.

PARITY-UNPINNED PART: ``constant_q_filters`` / ``cqt_frequencies`` restate
``librosa.filters.constant_q`` / ``librosa.time_frequency.cqt_frequencies`` (librosa <= 0.7.x,
not vendored in the reference, version unpinned, no golden vectors anywhere in the
reference).  They follow the library's published algorithm and are anchored only on the
structural facts the reference exposes (9 octave groups of sizes 16384..64, receptive
field 16384).
"""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# CQT filterbank (third-party arithmetic: librosa <= 0.7 filters.constant_q)
# --------------------------------------------------------------------------------------

HANN_BANDWIDTH = 1.50018310546875  # librosa.filters.WINDOW_BANDWIDTHS['hann']


def cqt_frequencies(n_bins, fmin, bins_per_octave=12):
    """librosa.time_frequency.cqt_frequencies (called at constant_q_transform.py:272-274)."""
    return float(fmin) * 2.0 ** (np.arange(n_bins, dtype=float) / bins_per_octave)


def constant_q_filters(sr, fmin, n_bins, bins_per_octave, filter_scale):
    """librosa.filters.constant_q(sr, fmin, n_bins, bins_per_octave, filter_scale) with the
    defaults the reference relies on (window='hann', pad_fft=True, norm=1, tuning=0);
    call site constant_q_transform.py:108-112.  Returns (filters complex128[n_bins, M],
    lengths float64[n_bins])."""
    q = float(filter_scale) / (2.0 ** (1.0 / bins_per_octave) - 1.0)
    freqs = cqt_frequencies(n_bins, fmin, bins_per_octave)
    if freqs[-1] * (1 + 0.5 * HANN_BANDWIDTH / q) > sr / 2.0:
        raise ValueError("Filter pass-band lies beyond Nyquist")
    lengths = q * sr / freqs
    freqs = q * sr / lengths                 # librosa 0.7 converts the lengths back to frequencies
    max_len = int(2.0 ** (np.ceil(np.log2(lengths.max()))))
    bank = np.zeros((n_bins, max_len), dtype=np.complex128)
    for k in range(n_bins):
        ilen = lengths[k]
        n = np.arange(-ilen // 2, ilen // 2, dtype=float)
        sig = np.exp(n * 1j * 2 * np.pi * freqs[k] / sr)
        m = len(sig)
        # periodic ("fftbins") Hann window of m samples
        win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(m) / m)
        sig = sig * win
        sig = sig / np.sum(np.abs(sig))          # util.normalize(norm=1)
        lpad = (max_len - m) // 2                # util.pad_center
        bank[k, lpad:lpad + m] = sig
    return bank, lengths


def cqt_group_plan(lengths):
    """Octave grouping of the filterbank, constant_q_transform.py:116-130: one strided conv per
    power-of-two kernel size.  Returns (kernel_sizes, [(lo, hi), ...])."""
    sizes, ranges = [], []
    current, last = None, 0
    for i, l in enumerate(lengths):
        ks = 2 ** math.ceil(np.log2(l))
        if current is not None and ks >= current:
            continue
        sizes.append(ks)
        current = ks
        if i != 0:
            ranges.append((last, i))
        last = i
    ranges.append((last, len(lengths)))
    return sizes, ranges


def cqt_group_weights(bank, sizes, ranges):
    """Per-group conv weights, constant_q_transform.py:132-146: centre-crop to K_g, stack
    [real; imag] on the channel axis, cast to fp32.  Returns list of (2 n_g, K_g) fp32."""
    m = bank.shape[-1]
    out = []
    for size, (lo, hi) in zip(sizes, ranges):
        off = (m - size) // 2
        filt = bank[lo:hi, off:m - off] if off > 0 else bank[lo:hi, :]
        out.append(np.concatenate([np.real(filt), np.imag(filt)], axis=0).astype(np.float32))
    return out


class CqtPlan:
    """Everything CQT.__init__ derives from its arguments (constant_q_transform.py:95-149)."""

    def __init__(self, sr=16000, fmin=30, n_bins=256, bins_per_octave=32, filter_scale=1.0, hop_length=128):
        self.sr, self.fmin, self.n_bins = sr, fmin, n_bins
        self.bins_per_octave, self.filter_scale, self.hop = bins_per_octave, filter_scale, hop_length
        self.bank, self.lengths = constant_q_filters(sr, fmin, n_bins, bins_per_octave, filter_scale)
        self.kernel_sizes, self.ranges = cqt_group_plan(self.lengths)
        self.weights = cqt_group_weights(self.bank, self.kernel_sizes, self.ranges)
        self.freqs = cqt_frequencies(n_bins, fmin, bins_per_octave)

    def n_frames(self, n_samples):
        return (n_samples - 1 - self.kernel_sizes[0]) // self.hop + 1


def cqt_forward(x, plan, dtype=torch.float32):
    """CQT.forward, constant_q_transform.py:161-172.  x (B,1,L) -> (B, n_bins, T, 2).
    Note the crop ``x[:, :, off:-(off+1)]`` drops the last sample for every group."""
    x = torch.as_tensor(x).to(dtype)
    k0 = plan.kernel_sizes[0]
    real, imag = [], []
    for size, w in zip(plan.kernel_sizes, plan.weights):
        off = (k0 - size) // 2
        wt = torch.from_numpy(w).to(dtype).unsqueeze(1)
        y = F.conv1d(x[:, :, off:x.shape[2] - (off + 1)], wt, stride=plan.hop)
        n = w.shape[0] // 2
        real.append(y[:, :n])
        imag.append(y[:, n:])
    return torch.stack([torch.cat(real, 1), torch.cat(imag, 1)], dim=3)


def phase_constants(plan):
    """PhaseDifference.__init__, constant_q_transform.py:272-280 -> (fixed_phase_diff, scaling),
    both fp32 vectors of n_bins (the reference casts them to FloatTensor)."""
    fixed = (((1.0 * plan.freqs * plan.hop / plan.sr) + 0.5) % 1 - 0.5) * 2 * np.pi
    scaling = 1.0 / np.log(plan.freqs)
    return fixed.astype(np.float32), scaling.astype(np.float32)


def phase_difference(phi, plan):
    """PhaseDifference.forward + unwrap, constant_q_transform.py:69-72, 282-286.
    phi (B,F,T) -> (B,F,T-1)."""
    fixed, scaling = phase_constants(plan)
    fixed = torch.from_numpy(fixed).to(phi.dtype).view(1, -1, 1)
    scaling = torch.from_numpy(scaling).to(phi.dtype).view(1, -1, 1)
    pd = phi[:, :, 1:] - phi[:, :, :-1] + fixed
    pd = torch.where(pd > math.pi, pd - 2 * math.pi, pd)
    pd = torch.where(pd < -math.pi, pd + 2 * math.pi, pd)
    return pd * scaling


def preprocess(x, plan, phase=False, offset_zero=False, output_power=1.0, pooling=None, scaling=1.0,
               dtype=torch.float32):
    """PreprocessingModule.forward, scalogram_model.py:75-102 (ctor constants :62-69).
    x (B,1,L) -> (B,1,F,T) or (B,2,F,T-1) [time halves with pooling=[1,2]]."""
    z = cqt_forward(x, plan, dtype)
    if offset_zero:
        eps = 1e-9
        log_offset = -math.log(eps)
        norm = scaling / log_offset
    else:
        eps, log_offset, norm = 0.0, 0.0, scaling
    re, im = z[..., 0], z[..., 1]
    if phase:
        amp = torch.sqrt(re[:, :, 1:] ** 2 + im[:, :, 1:] ** 2) ** 2
        amp = torch.log(amp + eps) + log_offset
        phi = phase_difference(torch.atan2(im, re), plan)
        y = torch.stack([amp, phi], dim=1)
    else:
        y = torch.sqrt(re ** 2 + im ** 2) ** 2
        y = torch.log(y + eps).unsqueeze(1) + log_offset
    if pooling is not None:
        y = F.max_pool2d(y, pooling)
    y = y * norm
    y = y ** output_power
    return y


# --------------------------------------------------------------------------------------
# Encoders (functional restatements; parameters are passed in explicitly)
# --------------------------------------------------------------------------------------

def audio_encoder_geometry(kernel_sizes, strides):
    """AudioEncoder.__init__, audio_model.py:19-25 -> (receptive_field, downsampling_factor)."""
    rf, s = kernel_sizes[0], 1
    for i in range(1, len(strides)):
        s *= strides[i - 1]
        rf += (kernel_sizes[i] - 1) * s
    return rf, int(np.prod(strides))


def audio_encoder_forward(x, weights, biases, strides):
    """AudioEncoder.forward, audio_model.py:36-44: strided conv1d, ReLU after all but the last."""
    for l, (w, b, s) in enumerate(zip(weights, biases, strides)):
        x = F.conv1d(x, w, b, stride=s)
        if l < len(weights) - 1:
            x = F.relu(x)
    return x


def encoder_block_forward(x, cfg, p, training=True, bn_eps=1e-5):
    """ScalogramEncoderBlock.forward, scalogram_model.py:387-479.
    ``cfg`` is the block dict (configs/scalogram_resnet_configs.py:3-20); ``p`` maps
    'conv_a.weight', 'conv_a.bias', 'bn_a.weight', 'bn_a.bias', 'conv_b.*', 'bn_b.*',
    'res.weight' to tensors.  BatchNorm uses batch statistics when ``training``."""

    def bn(t, name):
        if not cfg['batch_norm']:
            return t
        return F.batch_norm(t, p.get(name + '.running_mean'), p.get(name + '.running_var'),
                            p[name + '.weight'], p[name + '.bias'], training or p.get(name + '.running_mean') is None,
                            0.1, bn_eps)

    def half(t, which, conv):
        if cfg['top_padding_' + which] is not None:
            t = F.pad(t, (0, 0, cfg['top_padding_' + which], 0))
        t = F.conv2d(t, p[conv + '.weight'], p.get(conv + '.bias'), stride=cfg['stride_' + which],
                     padding=cfg['padding_' + which])
        t = bn(t, 'bn_' + conv[-1])
        if cfg['pooling_' + which] > 1:
            t = F.max_pool2d(t, cfg['pooling_' + which], ceil_mode=cfg['ceil_pooling'])
        return F.relu(t)

    main = half(half(x, '1', 'conv_a'), '2', 'conv_b')
    if cfg['residual']:
        res = x
        sp = cfg['stride_1'] * cfg['stride_2'] * cfg['pooling_1'] * cfg['pooling_2']
        if sp > 1:
            res = F.max_pool2d(res, sp, ceil_mode=True)
        if cfg['in_channels'] != cfg['out_channels']:
            res = F.conv2d(res, p['res.weight'], None, padding=cfg['padding_1'] + cfg['padding_2'])
        o_h = (res.shape[2] - main.shape[2] + 1) / 2
        o_w = (res.shape[3] - main.shape[3] + 1) / 2
        if int(o_h) > 0:
            res = res[:, :, -int(o_h + main.shape[2]):-int(o_h), :]
        if int(o_w) > 0:
            res = res[:, :, :, -int(o_w + main.shape[3]):-int(o_w)]
        main = main + res
    return main


def residual_encoder_forward(x, block_cfgs, block_params, training=True):
    """ScalogramResidualEncoder.forward, scalogram_model.py:520-529."""
    if x.dim() == 3:
        x = x.unsqueeze(2)
    for i, (cfg, p) in enumerate(zip(block_cfgs, block_params)):
        x = encoder_block_forward(x, cfg, p, training)
        if i < len(block_cfgs) - 1:
            x = F.relu(x)
    return x[:, :, 0, :]


# --------------------------------------------------------------------------------------
# Model composition + InfoNCE
# --------------------------------------------------------------------------------------

def predictive_split(z, visible_steps, prediction_steps):
    """AudioPredictiveCodingModel.forward slicing, audio_model.py:197-198."""
    targets = z[:, :, -prediction_steps:]
    vis = z[:, :, -(visible_steps + prediction_steps):-prediction_steps]
    return targets, vis


def scores_full(pred, targets, kind='linear'):
    """linear/softplus score function, contrastive_estimation_training.py:12-22.
    pred (B,K,E), targets (B,E,K) -> (B,K,B,K) = [data_batch, data_step, target_batch, target_step]."""
    s = torch.tensordot(pred, targets, dims=([2], [1]))
    if kind == 'softplus':
        s = F.softplus(s)
    elif kind != 'linear':
        raise ValueError(kind)
    return s


def infonce_loss(pred, targets, all_steps, kind='linear', regularization=0.0):
    """Training loss block, contrastive_estimation_training.py:106-122,141 restated literally
    (including the per-step layout-scrambling ``.view``, :116-118).
    Returns (loss, max_score) -- max_score is what the trainer logs at :166, i.e. the max of
    the *reduced* score tensor in per-step mode and of the full tensor in all-steps mode."""
    b, k, _ = pred.shape
    scores = scores_full(pred, targets, kind)
    if all_steps:
        noise = torch.logsumexp(scores.reshape(-1, b, k), dim=0)
        valid = torch.diagonal(torch.diagonal(scores, dim1=0, dim2=2), dim1=0, dim2=1)
    else:
        scores = torch.diagonal(scores, dim1=1, dim2=3).permute(0, 2, 1).contiguous()  # (d, k, t)
        noise = torch.logsumexp(scores.view(-1, b, k), dim=0)
        valid = torch.diagonal(scores, dim1=0, dim2=2).permute(1, 0)
    loss = torch.mean(-torch.mean(valid - noise, dim=1))
    loss = loss + regularization * torch.mean(torch.mean(scores, dim=1) ** 2)
    return loss, scores.max()


def gradient_penalty(pred, targets, scalogram, all_steps, kind='linear', factor=10.0):
    """Wasserstein gradient penalty, contrastive_estimation_training.py:144-155: sum of the (in per-step mode already
    diagonal-reduced, :116) scores differentiated w.r.t. the scalogram with create_graph, then
    factor * mean((||grad||_2 over the channel axis - 1)^2).  ``scalogram`` must require grad and be the tensor the
    encoder consumed."""
    scores = scores_full(pred, targets, kind)
    if not all_steps:
        scores = torch.diagonal(scores, dim1=1, dim2=3).permute(0, 2, 1).contiguous()
    grad, = torch.autograd.grad(outputs=scores.sum(), inputs=scalogram, create_graph=True, retain_graph=True,
                                only_inputs=True)
    return ((grad.norm(2, dim=1) - 1) ** 2).mean() * factor


def infonce_loss_clean(pred, targets, all_steps, kind='linear', regularization=0.0):
    """The same loss written as the cross-entropy it is: softmax over the
    predictions (d[,k]) for each target (t,k').  Equal to ``infonce_loss`` because the
    reference's scrambling view is a bijection under the mean."""
    b, k, e = pred.shape
    if all_steps:
        s = scores_full(pred, targets, kind).reshape(b * k, b * k)      # rows (d,k), cols (t,k')
        lse = torch.logsumexp(s, dim=0)
        loss = (lse - torch.diagonal(s)).mean()
        reg = (scores_full(pred, targets, kind).mean(dim=1) ** 2).mean()
    else:
        u = torch.einsum('dke,tek->dkt', pred, targets)
        s = F.softplus(u) if kind == 'softplus' else u
        lse = torch.logsumexp(s, dim=0)                                  # (k, t)
        valid = torch.diagonal(s, dim1=0, dim2=2)                        # (k, b)
        loss = (lse - valid).mean()
        reg = (s.mean(dim=1) ** 2).mean()
    return loss + regularization * reg


def infonce_with_grads(pred, targets, all_steps, kind='linear', regularization=0.0, dtype=torch.float64):
    """Loss, logged max score and d loss/d pred, d loss/d targets via CPU autograd."""
    p = pred.detach().to(dtype).requires_grad_(True)
    z = targets.detach().to(dtype).requires_grad_(True)
    loss, mx = infonce_loss(p, z, all_steps, kind, regularization)
    loss.backward()
    return loss.detach(), mx.detach(), p.grad, z.grad


def validation_metrics(pred, targets, all_steps, kind='linear'):
    """validate() per-batch block, contrastive_estimation_training.py:224-247: per-step losses
    (mean over dim 0, with the scrambled view in per-step mode), accuracy (argmax over the
    target axis), mean score."""
    b, k, _ = pred.shape
    scores = scores_full(pred, targets, kind)
    n = b * k if all_steps else b
    if all_steps:
        noise = torch.logsumexp(scores.reshape(-1, b, k), dim=0)
        valid = torch.diagonal(torch.diagonal(scores, dim1=0, dim2=2), dim1=0, dim2=1)
        template = torch.arange(n).view(b, k)
    else:
        scores = torch.diagonal(scores, dim1=1, dim2=3).permute(0, 2, 1).contiguous()
        noise = torch.logsumexp(scores.view(-1, b, k), dim=0)
        valid = torch.diagonal(scores, dim1=0, dim2=2).permute(1, 0)
        template = torch.arange(n).unsqueeze(1).repeat(1, k)
    losses = -torch.mean(valid - noise, dim=0)
    best = torch.argmax(scores.reshape(b, k, -1), dim=2)
    acc = torch.sum(torch.eq(template, best), dim=0).to(pred.dtype) / n
    return losses, acc, scores.mean()


# --------------------------------------------------------------------------------------
# Sampler (integer path -- must be bit exact)
# --------------------------------------------------------------------------------------

def file_batch_sampler(index_count_per_file, batch_size, file_batch_size=1, drop_last=True, seed=None,
                       rng=random):
    """FileBatchSampler.__iter__, audio_dataset.py:235-260 -> list of index lists.
    ``rng`` is the Python ``random`` module (global Mersenne Twister state when seed is None)."""
    if drop_last:
        per_file = [n // file_batch_size for n in index_count_per_file]
    else:
        per_file = [-(-n // file_batch_size) for n in index_count_per_file]
    total = int(sum(per_file))                                          # __len__, :262-263

    def chunks(l, n):
        out = []
        for i in range(0, len(l), n):
            if drop_last and i + n > len(l):
                break
            out.append(l[i:i + n])
        return out

    if file_batch_size == 1:
        order = list(range(total))
        if seed is not None:
            rng.seed(seed)
        rng.shuffle(order)
        return chunks(order, batch_size)
    files, s = [], 0
    for n in index_count_per_file:
        files.append(list(range(s, s + n)))
        s += n
    for i, f in enumerate(files):
        if seed is not None:
            rng.seed(seed + i)
        rng.shuffle(f)
    batches = []
    for f in files:
        batches.extend(chunks(f, file_batch_size))
    if seed is not None:
        rng.seed(seed)
    rng.shuffle(batches)
    per_batch = batch_size // file_batch_size
    if per_batch > 1:
        merged = []
        for i in range(0, len(batches), per_batch):
            if drop_last and i + per_batch > len(batches):
                break
            merged.append([j for c in batches[i:i + per_batch] for j in c])
        batches = merged
    return batches
