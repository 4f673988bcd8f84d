"""Autoregressive (context) models.  The north star keeps these *unchanged*: they are caller-side modules
plugged into ``AudioPredictiveCodingModel`` and run on stock PyTorch.  They are restated here only so that
the package is usable without the reference checkout (audio_model.py:47-161, attention_model.py:9-82).

Three deliberate deviations, all numerically identical in the forward pass (the third also in the backward pass):
  * ``ConvolutionalArBlock`` adds the residual out of place (the reference's in-place ``+=`` on a ReLU
    output raises an autograd error on torch >= 1.5);
  * ``PositionalEncoder`` scales out of place (the reference scales a view of ``z`` in place);
  * a ``ConvolutionalArBlock`` whose two branches start with the same pooling of ``x`` computes it once.
On a B200 the conv1d and pooling layers of the convolutional AR model (``ArConv1d``, ``ArMaxPool1d``) run on the package's
kernels in fp32 arithmetic; the modules, parameters and state_dict keys are the stock ones.
"""
import math
import warnings

import torch
import torch.nn as nn

from . import ops
from .model import ActivationWriter


class ArConv1d(nn.Conv1d):
    """``nn.Conv1d`` of the AR models (same parameters / state_dict keys).  On a B200 in the fp32 parity mode its
    forward / backward run on the same fp32-faithful tcgen05 implicit-GEMM kernels as the encoder instead of cuDNN --
    cuDNN's fp32 path is either TF32 (not fp32: 10-bit mantissa) or CUDA-core FMA (2.8 ms of a 14 ms e24 step).  Any
    other input (CPU tensors, other dtypes, dilation / groups) takes the stock torch path: the AR model is caller-side
    code and stays usable everywhere."""

    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and self.groups == 1 and self.dilation == (1,)
                and self.padding_mode == 'zeros' and not isinstance(self.padding, str) and self.in_channels >= 16
                and not ops.second_order_enabled()):
            return ops.conv1d(x, self.weight, self.bias, self.stride[0], self.padding[0])
        return super().forward(x)


class ArMaxPool1d(nn.MaxPool1d):
    """``nn.MaxPool1d`` of the AR blocks (audio_model.py:95-97, 110-112: ``MaxPool1d(pooling, ceil_mode=True)``).  On a
    B200 its non-overlapping, unpadded case runs on the package's pooling kernels, as 1 x k windows of the (B, C, 1, T)
    view (a k x k window clipped to the single row): ATen's max_pool backward takes 22 us per call on these
    (64, 512, <= 60) tensors.  Anything else takes the stock path."""

    def runs_on_kernels(self, x):
        k = self.kernel_size if isinstance(self.kernel_size, int) else None
        stride = self.stride if isinstance(self.stride, int) else None
        return (k is not None and stride == k and self.padding == 0 and self.dilation == 1 and not self.return_indices
                and x.is_cuda and x.dim() == 3 and x.dtype == torch.float32 and (self.ceil_mode or x.shape[2] % k == 0)
                and not ops.second_order_enabled())

    def same_pooling(self, other):
        return (isinstance(other, nn.MaxPool1d) and other.kernel_size == self.kernel_size and other.stride == self.stride
                and other.padding == self.padding and other.dilation == self.dilation
                and other.ceil_mode == self.ceil_mode and not other.return_indices and not self.return_indices)

    def forward(self, x):
        if self.runs_on_kernels(x):
            return ops.max_pool2d(x.unsqueeze(2), self.kernel_size, True).squeeze(2)
        return super().forward(x)


class AudioGRUModel(nn.Module):
    """GRUCell unrolled over the visible steps; input (batch, input_size, steps) -> last hidden state
    (audio_model.py:47-77).  The parameters are the stock ``nn.GRUCell``'s (state_dict keys ``gruCell.*``).

    On a CUDA device the recurrence runs as ONE fused library call over the whole sequence (``torch._VF.gru`` = cuDNN's GRU,
    the same gate equations as ``nn.GRUCell``: r, z, n with ``n = tanh(W_in x + b_in + r * (W_hn h + b_hn))``) instead of
    ``steps`` cell launches forward and ~8 x ``steps`` backward: the raw-wave configuration (BASELINE configs[0], 100
    steps) was bound by exactly those launches.  The unrolled loop remains for CPU tensors and for second-order autograd
    (cuDNN's RNN backward is not differentiable again); ``fused`` = True / False forces one path (tests)."""

    def __init__(self, input_size, hidden_size, bias=True, reset_hidden=True):
        super().__init__()
        self.gruCell = nn.GRUCell(input_size=input_size, hidden_size=hidden_size, bias=bias)
        self.hidden = None
        self.reset_hidden = reset_hidden
        self.fused = None                                        # None: fused on CUDA outside second-order mode

    def _use_fused(self, input):
        if self.fused is not None:
            return bool(self.fused)
        return input.is_cuda and not ops.second_order_enabled()

    def forward(self, input):
        state = None if self.reset_hidden else self.hidden
        if self._use_fused(input):
            cell = self.gruCell
            weights = [cell.weight_ih, cell.weight_hh] + ([cell.bias_ih, cell.bias_hh] if cell.bias else [])
            h0 = state if state is not None else input.new_zeros(input.shape[0], cell.hidden_size)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")                  # "weights are not one contiguous chunk": 0.6 M floats, copied
                _, h_n = torch._VF.gru(input.permute(2, 0, 1), h0.unsqueeze(0), weights, bool(cell.bias), 1, 0.0,
                                       self.training, False, False)
            state = h_n[0]
        else:
            for frame in input.unbind(dim=2):
                state = self.gruCell(frame, state)
        self.hidden = None if self.reset_hidden else state
        return state


class ConvolutionalArBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, pooling=1, stride=1, bias=True, residual=False,
                 batch_norm=False, name='ar_block', activation_register=None):
        super().__init__()
        self.name = name
        self.main_modules = nn.ModuleList()
        if pooling > 1:
            self.main_modules.append(ArMaxPool1d(pooling, ceil_mode=True))
        self.main_modules.append(ArConv1d(in_channels, out_channels, kernel_size, stride=stride, bias=bias))
        if batch_norm:
            self.main_modules.append(nn.BatchNorm1d(out_channels))
        self.main_modules.append(nn.ReLU())
        self.residual = residual
        self.residual_modules = None
        if residual:
            self.residual_modules = nn.ModuleList()
            if pooling * stride > 1:
                self.residual_modules.append(ArMaxPool1d(pooling * stride, ceil_mode=True))
            if in_channels != out_channels:
                self.residual_modules.append(ArConv1d(in_channels, out_channels, kernel_size=1))
        self.output_activation_writer = ActivationWriter(register=activation_register, name=self.name)

    def forward(self, x):
        main, pooled = x, None
        for i, m in enumerate(self.main_modules):
            main = m(main)
            if i == 0 and isinstance(m, ArMaxPool1d):
                pooled = main
        if self.residual:
            skip = x
            for i, m in enumerate(self.residual_modules):
                # with stride 1 both branches start with the same pooling of x: computed once
                if i == 0 and pooled is not None and self.main_modules[0].same_pooling(m):
                    skip = pooled
                else:
                    skip = m(skip)
            main = main + skip[:, :, -main.shape[2]:]
        self.output_activation_writer(main)
        return main


class ConvolutionalArModel(nn.Module):
    def __init__(self, args_dict):
        super().__init__()
        self.module_list = nn.ModuleList()
        for l, kernel in enumerate(args_dict['kernel_sizes']):
            self.module_list.append(ConvolutionalArBlock(in_channels=args_dict['channel_count'][l],
                                                         out_channels=args_dict['channel_count'][l + 1],
                                                         kernel_size=kernel, stride=args_dict['stride'][l],
                                                         pooling=args_dict['pooling'][l], bias=args_dict['bias'],
                                                         batch_norm=args_dict['batch_norm'],
                                                         residual=args_dict['residual'], name='ar_block_' + str(l),
                                                         activation_register=args_dict.get('activation_register')))
        self.encoding_size = args_dict['channel_count'][0]
        self.ar_size = args_dict['channel_count'][-1]

    def forward(self, x):
        for m in self.module_list:
            x = m(x)
        return x[:, :, -1]


class PositionalEncoder(nn.Module):
    def __init__(self, code_size, max_seq_len=128, max_wavelength=10000):
        super().__init__()
        self.code_size = code_size
        pos = torch.arange(max_seq_len, dtype=torch.float64).unsqueeze(1)
        i = torch.arange(0, code_size, 2, dtype=torch.float64).unsqueeze(0)
        arg = math.pi * pos / (max_wavelength ** ((2 * i) / code_size))
        pe = torch.zeros(max_seq_len, code_size, dtype=torch.float64)
        pe[:, 0::2] = torch.sin(arg)
        pe[:, 1::2] = torch.cos(arg)
        self.register_buffer('pe', pe.float().unsqueeze(1))

    def forward(self, x):
        return x * math.sqrt(self.code_size) + self.pe[:x.size(0)]


class AttentionModel(nn.Module):
    """Causally masked transformer encoder, mean over time, linear head (attention_model.py:38-82)."""

    def __init__(self, args_dict):
        super().__init__()
        channels = args_dict['channels']
        self.num_layers = args_dict['num_layers']
        self.positional_encoder = PositionalEncoder(channels, args_dict['sequence_length'])
        layer = nn.TransformerEncoderLayer(channels, args_dict['num_heads'], args_dict['feedforward_size'],
                                           args_dict['dropout'])
        self.encoder = nn.TransformerEncoder(layer, args_dict['num_layers'], nn.LayerNorm(channels),
                                             enable_nested_tensor=False)
        self.end_layer = nn.Linear(channels, args_dict['output_size'])

    def forward(self, x):
        x = self.positional_encoder(x.permute(2, 0, 1))          # sequence, batch, channels
        s = x.shape[0]
        mask = torch.triu(torch.full((s, s), float('-inf'), device=x.device), diagonal=1)
        x = self.encoder(x, mask=mask, is_causal=True)          # the hint skips torch's mask inspection (a host sync)
        return self.end_layer(x.sum(dim=0) / s)
