"""CPU ORACLE (whole training step) -- TEST / BASELINE INFRASTRUCTURE ONLY.

A plain-PyTorch CPU port of the reference's training step for the benchmark configurations, built from
the functional restatements in ``cpc_oracle.py``: PreprocessingModule -> ScalogramResidualEncoder ->
ConvolutionalArModel -> W_k -> InfoNCE -> backward -> Adam (contrastive_estimation_training.py:97-162,
scalogram_model.py:372-529, audio_model.py:80-213).  ``bench.py`` times it on the host cores as the
``cpu_baseline`` / ``--impl reference`` arm (kind "port": /root/reference does not exist on the GPU box)
and ``__graft_entry__.smoke()`` uses it as the checker.  Never imported by the product.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import cpc_oracle as O


def block_7(i, o, k1=(3, 3), k2=(3, 3), **kw):
    d = {'in_channels': i, 'hidden_channels': None, 'out_channels': o, 'kernel_size_1': k1, 'kernel_size_2': k2,
         'top_padding_1': None, 'top_padding_2': None, 'padding_1': 0, 'padding_2': 0, 'stride_1': 1, 'stride_2': 1,
         'pooling_1': 1, 'pooling_2': 1, 'bias': True, 'separable': False, 'residual': True, 'batch_norm': False,
         'ceil_pooling': False}
    d.update(kw)
    return d


def arch7_blocks(in_channels=2):
    """scalogram_resnet_architecture_7 as imported (configs/scalogram_resnet_configs.py:215-257)."""
    return [block_7(in_channels, 32, k2=(64, 1), top_padding_2=63, stride_1=2, batch_norm=True),
            block_7(32, 128, k2=(30, 1), stride_1=2, batch_norm=True),
            block_7(128, 256, k2=(15, 1), stride_1=2, batch_norm=True),
            block_7(256, 512, k1=(2, 2), k2=(1, 1))]


class OracleBlock(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        hid = cfg['hidden_channels'] or cfg['out_channels']
        self.conv_a = nn.Conv2d(cfg['in_channels'], hid, cfg['kernel_size_1'], bias=cfg['bias'])
        self.conv_b = nn.Conv2d(hid, cfg['out_channels'], cfg['kernel_size_2'], bias=cfg['bias'])
        if cfg['batch_norm']:
            self.bn_a, self.bn_b = nn.BatchNorm2d(hid), nn.BatchNorm2d(cfg['out_channels'])
        if cfg['residual'] and cfg['in_channels'] != cfg['out_channels']:
            self.res = nn.Conv2d(cfg['in_channels'], cfg['out_channels'], 1, bias=False)

    def params(self):
        p = {'conv_a.weight': self.conv_a.weight, 'conv_a.bias': self.conv_a.bias,
             'conv_b.weight': self.conv_b.weight, 'conv_b.bias': self.conv_b.bias}
        if self.cfg['batch_norm']:
            for n in ('bn_a', 'bn_b'):
                m = getattr(self, n)
                p[n + '.weight'], p[n + '.bias'] = m.weight, m.bias
        if hasattr(self, 'res'):
            p['res.weight'] = self.res.weight
        return p

    def forward(self, x):
        return O.encoder_block_forward(x, self.cfg, self.params(), training=True)


class OracleConvAr(nn.Module):
    """ConvolutionalArModel (audio_model.py:80-161) with the residual add written out of place."""

    def __init__(self, kernel_sizes, channels, pooling, batch_norm=True, residual=True):
        super().__init__()
        self.pooling, self.residual = pooling, residual
        self.convs = nn.ModuleList(nn.Conv1d(channels[i], channels[i + 1], k) for i, k in enumerate(kernel_sizes))
        self.bns = nn.ModuleList(nn.BatchNorm1d(channels[i + 1]) if batch_norm else nn.Identity()
                                 for i in range(len(kernel_sizes)))
        self.skips = nn.ModuleList(nn.Conv1d(channels[i], channels[i + 1], 1) if channels[i] != channels[i + 1]
                                   else nn.Identity() for i in range(len(kernel_sizes)))

    def forward(self, x):
        for conv, bn, skip, pool in zip(self.convs, self.bns, self.skips, self.pooling):
            main = F.max_pool1d(x, pool, ceil_mode=True) if pool > 1 else x
            main = F.relu(bn(conv(main)))
            if self.residual:
                r = F.max_pool1d(x, pool, ceil_mode=True) if pool > 1 else x
                main = main + skip(r)[:, :, -main.shape[2]:]
            x = main
        return x[:, :, -1]


class OracleE24(nn.Module):
    """experiments['e24']: CQT(+phase) -> arch 7 -> conv AR arch 3 -> Linear(256 -> 16*512)."""

    def __init__(self, visible_steps=60, prediction_steps=16):
        super().__init__()
        self.plan = O.CqtPlan(16000, 30, 256, 32, 0.5, 128)
        self.blocks = nn.ModuleList(OracleBlock(c) for c in arch7_blocks(2))
        self.ar = OracleConvAr([5] * 6, [512, 512, 512, 256, 256, 256, 256], [1, 1, 2, 1, 2, 1])
        self.v, self.k, self.e = visible_steps, prediction_steps, 512
        self.predict = nn.Linear(256, self.k * self.e, bias=False)
        self.item_length = 19200 + (visible_steps + prediction_steps) * 1024

    def forward(self, audio):
        with torch.no_grad():
            x = O.preprocess(audio.unsqueeze(1), self.plan, phase=True)
        x = x.requires_grad_(True)
        for i, b in enumerate(self.blocks):
            x = b(x)
            if i < len(self.blocks) - 1:
                x = F.relu(x)
        z = x[:, :, 0, :]
        self.last_z = z.detach()
        targets, vis = O.predictive_split(z, self.v, self.k)
        pred = self.predict(self.ar(vis)).view(-1, self.k, self.e)
        return pred, targets


class OracleRawWave(nn.Module):
    """BASELINE configs[0]: AudioEncoder (audio_model.py:14-44, default dict: 512 channels, strides 5,4,2,2,2) ->
    AudioGRUModel(512, 256) (audio_model.py:47-77: a GRUCell unrolled over the visible steps) -> Linear(256 -> K*512)."""

    def __init__(self, visible_steps=100, prediction_steps=12, channels=512, ar_size=256):
        super().__init__()
        self.kernel_sizes, self.strides = [10, 8, 4, 4, 4], [5, 4, 2, 2, 2]
        widths = [1] + (list(channels) if isinstance(channels, (list, tuple)) else [channels] * 5)
        channels = widths[-1]
        self.convs = nn.ModuleList(nn.Conv1d(widths[i], widths[i + 1], k, stride=s)
                                   for i, (k, s) in enumerate(zip(self.kernel_sizes, self.strides)))
        self.gru = nn.GRUCell(channels, ar_size)
        self.v, self.k, self.e = visible_steps, prediction_steps, channels
        self.predict = nn.Linear(ar_size, self.k * self.e, bias=False)
        rf, ds = O.audio_encoder_geometry(self.kernel_sizes, self.strides)
        self.item_length = rf + (visible_steps + prediction_steps) * ds

    def forward(self, audio):
        z = O.audio_encoder_forward(audio.unsqueeze(1), [c.weight for c in self.convs], [c.bias for c in self.convs],
                                    self.strides)
        targets, vis = O.predictive_split(z, self.v, self.k)
        h = None
        for t in range(vis.shape[2]):
            h = self.gru(vis[:, :, t], h)
        pred = self.predict(h).view(-1, self.k, self.e)
        return pred, targets


def train_steps(model, audio_batches, lr=1e-4, all_steps=True, kind='linear', regularization=0.0):
    """Runs one optimiser step per batch; returns the list of loss values."""
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses = []
    model.train()
    for audio in audio_batches:
        pred, targets = model(audio)
        loss, _ = O.infonce_loss(pred, targets, all_steps, kind, regularization)
        model.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return losses


# ---------------------------------------------------------------------------------------------------
# Shared fixtures for the full-size e24 golden (tests/golden/e24_step.npz, oracle/make_golden.py::golden_e24)
# ---------------------------------------------------------------------------------------------------

def _name_seed(name, salt=0):
    import zlib
    return (zlib.crc32(name.encode()) + salt) % (2 ** 31)


def seeded_values(shapes):
    """{name: shape} -> {name: tensor}: deterministic parameter values keyed by the parameter NAME, so that the
    reference (make_golden.py), this oracle and the CUDA product can be given identical weights without a 37 MB
    fixture.  Conv / linear weights and their biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's default
    bound); batch-norm scales ~ U(0.8, 1.2), shifts ~ U(-0.1, 0.1) (non-trivial on purpose)."""
    out = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        g = torch.Generator().manual_seed(_name_seed(name))
        u = torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1
        numel = 1
        for d in shape:
            numel *= d
        if len(shape) >= 2:
            val = u * (1.0 / (numel / shape[0]) ** 0.5)
            if name.startswith("prediction_model."):
                val = val * 0.05                                         # keeps the initial scores O(1): softmax not saturated
        else:
            w = shapes.get(name[:-len("bias")] + "weight") if name.endswith("bias") else None
            if w is not None and len(w) >= 2:                            # conv / linear bias
                wn = 1
                for d in w:
                    wn *= d
                val = u * (1.0 / (wn / w[0]) ** 0.5)
            elif name.endswith("weight"):                                # batch-norm scale
                val = 1.0 + 0.2 * u
            else:                                                        # batch-norm shift
                val = 0.1 * u
        out[name] = val
    return out


def reseed_parameters(named_parameters, name_map=None):
    """Overwrite parameters in place with ``seeded_values``.  ``name_map`` (reference key -> own key) lets a model
    with different parameter names (OracleE24) take the values of the reference's keys."""
    named = dict(named_parameters)
    if name_map is None:
        name_map = {n: n for n in named}
    assert set(name_map.values()) == set(named), set(name_map.values()) ^ set(named)
    values = seeded_values({ref: tuple(named[own].shape) for ref, own in name_map.items()})
    with torch.no_grad():
        for ref, own in name_map.items():
            named[own].copy_(values[ref].to(named[own].device))


def subsample_index(name, numel, count=8192):
    """Sorted, seeded subset of flat indices (all of them when the tensor is small) used to store gradients /
    parameters of the full-size model compactly."""
    if numel <= count:
        return torch.arange(numel)
    g = torch.Generator().manual_seed(_name_seed(name, 1))
    return torch.sort(torch.randperm(numel, generator=g)[:count]).values


def e24_audio(batch, length=97024, seed=1234):
    """The BASELINE synthetic input: 0.1 * randn, seeded."""
    return 0.1 * torch.randn(batch, length, generator=torch.Generator().manual_seed(seed))


def oracle_e24_name_map():
    """reference state_dict key -> OracleE24 parameter name (e24: arch 7 as imported + ar_conv_architecture_3)."""
    m = {}
    main = {0: {0: 'conv_a', 1: 'bn_a', 5: 'conv_b', 6: 'bn_b'}, 1: {0: 'conv_a', 1: 'bn_a', 4: 'conv_b', 5: 'bn_b'},
            2: {0: 'conv_a', 1: 'bn_a', 4: 'conv_b', 5: 'bn_b'}, 3: {0: 'conv_a', 3: 'conv_b'}}
    res_idx = {0: 1, 1: 1, 2: 1, 3: 0}
    for b, mods in main.items():
        for idx, ours in mods.items():
            for leaf in ('weight', 'bias'):
                m['encoder.blocks.%d.main_modules.%d.%s' % (b, idx, leaf)] = 'blocks.%d.%s.%s' % (b, ours, leaf)
        m['encoder.blocks.%d.residual_modules.%d.weight' % (b, res_idx[b])] = 'blocks.%d.res.weight' % b
    pooling = [1, 1, 2, 1, 2, 1]
    chans = [512, 512, 512, 256, 256, 256, 256]
    for l, pool in enumerate(pooling):
        c = 1 if pool > 1 else 0
        for leaf in ('weight', 'bias'):
            m['autoregressive_model.module_list.%d.main_modules.%d.%s' % (l, c, leaf)] = 'ar.convs.%d.%s' % (l, leaf)
            m['autoregressive_model.module_list.%d.main_modules.%d.%s' % (l, c + 1, leaf)] = 'ar.bns.%d.%s' % (l, leaf)
            if chans[l] != chans[l + 1]:
                m['autoregressive_model.module_list.%d.residual_modules.%d.%s' % (l, c, leaf)] = 'ar.skips.%d.%s' % (l, leaf)
    m['prediction_model.weight'] = 'predict.weight'
    return m
