"""Encoders: ``AudioEncoder`` (raw wave) and the scalogram encoders.

Constructor dicts, attribute names, module-list ordering and therefore state_dict keys follow the
reference (audio_model.py:14-44; scalogram_model.py:129-227, 372-529) so that ``configs/*.py`` load
unchanged; every convolution runs through the sm_100a kernels behind ``cpc_conv_*``.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from . import frontend
from .frontend import CQT, PhaseDifference
from .model import ActivationWriter

encoder_default_dict = {'strides': [5, 4, 2, 2, 2],
                        'kernel_sizes': [10, 8, 4, 4, 4],
                        'channel_count': [512, 512, 512, 512, 512],
                        'bias': True}

cqt_default_dict = {'sample_rate': 16000, 'fmin': 30, 'n_bins': 256, 'bins_per_octave': 32,
                    'filter_scale': 0.5, 'hop_length': 128, 'trainable_cqt': False}

scalogram_encoder_default_dict = {'kernel_sizes': [(127, 1), (5, 5), (63, 1), (5, 5), (26, 1), (5, 5)],
                                  'top_padding': [126, 0, 0, 0, 0, 0],
                                  'channel_count': [1, 32, 32, 64, 128, 256, 512],
                                  'pooling': [1, 2, 1, 2, 1, 1],
                                  'stride': [1, 1, 1, 1, 1, 1],
                                  'bias': True, 'batch_norm': False, 'phase': False, 'separable': False,
                                  'lowpass_init': 0., 'instance_norm': False, 'dropout': 0.}


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class Conv1d(nn.Conv1d):
    """nn.Conv1d parameters, B200 kernel forward/backward."""

    fuse_relu = False

    def forward(self, x):
        if self.groups != 1 or self.dilation != (1,) or self.padding_mode != 'zeros' or isinstance(self.padding, str):
            raise NotImplementedError("cpc_b200.Conv1d supports groups=1, dilation=1, zero padding")
        return ops.conv1d(x, self.weight, self.bias, self.stride[0], self.padding[0], relu=self.fuse_relu)


class Conv2d(nn.Conv2d):
    """nn.Conv2d parameters, B200 kernel forward/backward; ``extra_top`` folds a preceding ZeroPad2d."""

    def supported(self):
        return not (self.groups != 1 or self.dilation != (1, 1) or self.padding_mode != 'zeros'
                    or isinstance(self.padding, str))

    def forward(self, x, extra_top=0, relu=False):
        if not self.supported():
            raise NotImplementedError("cpc_b200.Conv2d supports groups=1, dilation=1, zero padding")
        return ops.conv2d(x, self.weight, self.bias, self.stride, self.padding, extra_top, relu)


class MaxPool2d(nn.MaxPool2d):
    """nn.MaxPool2d whose non-overlapping, unpadded square case (all the reference uses) runs on the B200 kernels."""

    def runs_on_kernels(self, x):
        k = self.kernel_size if isinstance(self.kernel_size, int) else None
        stride = self.stride if isinstance(self.stride, int) else None
        return (k is not None and stride == k and self.padding == 0 and self.dilation == 1 and not self.return_indices
                and x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and not ops.second_order_enabled())

    def forward(self, x):
        if self.runs_on_kernels(x):
            return ops.max_pool2d(x, self.kernel_size, self.ceil_mode)
        return super().forward(x)


def _run_modules(modules, x):
    """Run a reference-ordered module list, folding ZeroPad2d(top) into the conv that follows it."""
    pending_top = 0
    for m in modules:
        if isinstance(m, nn.ZeroPad2d):
            left, right, top, bottom = m.padding
            if left or right or bottom:
                x = m(x)
            else:
                pending_top += top
            continue
        if isinstance(m, (Conv2d, Conv2dSeparable)):
            x = m(x, extra_top=pending_top)
            pending_top = 0
            continue
        if pending_top:
            x = F.pad(x, (0, 0, pending_top, 0))
            pending_top = 0
        x = m(x)
    if pending_top:
        x = F.pad(x, (0, 0, pending_top, 0))
    return x


class AudioEncoder(nn.Module):
    """audio_model.py:14-44: strided conv1d stack, ReLU after all but the last layer."""

    def __init__(self, args_dict=encoder_default_dict):
        super().__init__()
        self.num_layers = len(args_dict['strides'])
        self.downsampling_factor = np.prod(args_dict['strides'])
        field, jump = args_dict['kernel_sizes'][0], 1
        for k, s in zip(args_dict['kernel_sizes'][1:], args_dict['strides'][:-1]):
            jump *= s
            field += (k - 1) * jump
        self.receptive_field = field
        self.layers = nn.ModuleList()
        widths = [1] + list(args_dict['channel_count'])
        for l in range(self.num_layers):
            conv = Conv1d(widths[l], widths[l + 1], args_dict['kernel_sizes'][l], stride=args_dict['strides'][l],
                          bias=args_dict['bias'])
            conv.fuse_relu = l < self.num_layers - 1          # ReLU runs in the conv epilogue
            self.layers.append(conv)

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class Conv2dSeparable(nn.Module):
    """scalogram_model.py:532-544: depthwise conv (groups = in_channels, no bias) followed by a 1x1 conv.  Same
    sub-module names (``conv``, ``conv_1x1``) and therefore state_dict keys; the depthwise half runs on
    ``cpc_dwconv_*`` (memory-bound CUDA-core kernels), the 1x1 half on the implicit-GEMM kernels."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=False, groups=in_channels)
        self.conv_1x1 = Conv2d(in_channels, out_channels, 1, bias=bias)

    @property
    def weight(self):
        return self.conv.weight

    def forward(self, x, extra_top=0):
        dw = self.conv
        if dw.dilation != (1, 1) or dw.padding_mode != 'zeros' or isinstance(dw.padding, str):
            raise NotImplementedError("cpc_b200.Conv2dSeparable supports dilation=1, zero padding")
        return self.conv_1x1(ops.depthwise_conv2d(x, dw.weight, dw.stride, dw.padding, extra_top))


class ScalogramEncoder(nn.Module):
    """scalogram_model.py:129-227: own CQT + log power (+phase) + conv/pool/ReLU stack."""

    def __init__(self, args_dict=scalogram_encoder_default_dict):
        super().__init__()
        self.num_layers = len(args_dict['kernel_sizes'])
        self.cqt = CQT(sr=args_dict['sample_rate'], fmin=args_dict['fmin'], n_bins=args_dict['n_bins'],
                       bins_per_octave=args_dict['bins_per_octave'], filter_scale=args_dict['filter_scale'],
                       hop_length=args_dict['hop_length'], trainable=args_dict['trainable_cqt'])
        self.phase = args_dict['phase']
        if self.phase:
            args_dict['channel_count'][0] = 2
            self.phase_diff = PhaseDifference(sr=args_dict['sample_rate'], fmin=args_dict['fmin'],
                                              n_bins=args_dict['n_bins'],
                                              bins_per_octave=args_dict['bins_per_octave'],
                                              hop_length=args_dict['hop_length'])
        else:
            args_dict['channel_count'][0] = 1
        if args_dict['lowpass_init'] > 0:
            raise NotImplementedError("lowpass_init relies on torch.rfft, removed from torch; not supported")
        self.module_list = nn.ModuleList()
        for l in range(self.num_layers):
            ks = _pair(args_dict['kernel_sizes'][l])
            if args_dict['top_padding'][l] > 0:
                self.module_list.add_module('pad_' + str(l), nn.ZeroPad2d((0, 0, args_dict['top_padding'][l], 0)))
            if l > 0 and args_dict['separable']:
                conv = Conv2dSeparable(args_dict['channel_count'][l], args_dict['channel_count'][l + 1], ks,
                                       bias=args_dict['bias'], stride=args_dict['stride'][l])
            else:
                # bias only when the kernel is wider than one frame (scalogram_model.py:168-170)
                conv = Conv2d(args_dict['channel_count'][l], args_dict['channel_count'][l + 1], ks,
                              bias=args_dict['bias'] if ks[1] > 1 else False, stride=args_dict['stride'][l])
            self.module_list.add_module('conv_' + str(l), conv)
            if args_dict['pooling'][l] > 1:
                self.module_list.add_module('pooling_' + str(l), MaxPool2d(kernel_size=args_dict['pooling'][l]))
            if l < self.num_layers - 1:
                self.module_list.add_module('relu_' + str(l), nn.ReLU())
                if args_dict['dropout'] > 0.:
                    self.module_list.add_module('dropout_' + str(l), nn.Dropout2d(args_dict['dropout']))
                if args_dict['batch_norm']:
                    self.module_list.add_module('batch_norm_' + str(l),
                                                nn.BatchNorm2d(num_features=args_dict['channel_count'][l + 1]))
                if args_dict['instance_norm']:
                    self.module_list.add_module('instance_norm_' + str(l),
                                                nn.InstanceNorm2d(num_features=args_dict['channel_count'][l + 1],
                                                                  affine=True, track_running_stats=True))
        self.receptive_field = self.cqt.conv_kernel_sizes[0]
        s = args_dict['hop_length']
        for i in range(self.num_layers):
            self.receptive_field += (_pair(args_dict['kernel_sizes'][i])[1] - 1) * s
            s *= args_dict['pooling'][i] * args_dict['stride'][i]
        self.downsampling_factor = args_dict['hop_length'] * np.prod(args_dict['pooling']) * np.prod(args_dict['stride'])

    def forward(self, x):
        if self.cqt.needs_autograd(x):
            # trainable filterbank / differentiable audio: scalogram_model.py:212-222 term by term on the conv kernels
            z = self.cqt.forward_differentiable(x)
            if self.phase:
                amp = torch.log(torch.pow(frontend.abs(z[:, :, 1:]), 2) + 1e-9)
                x = torch.stack([amp, self.phase_diff(frontend.angle(z))], dim=1)
            else:
                x = torch.log(torch.pow(frontend.abs(z), 2) + 1e-9).unsqueeze(1)
        elif self.phase:
            x = ops.cqt_frontend(x, self.cqt.packed_weights(), self.cqt.kernel_plan(), _lib.CQT_LOGPOW_PHASE,
                                 phase_fixed=self.phase_diff.fixed_phase_diff.reshape(-1),
                                 phase_scale=self.phase_diff.scaling.reshape(-1), eps=1e-9,
                                 packed_filters=self.cqt.tensor_core_filters())
        else:
            x = ops.cqt_frontend(x, self.cqt.packed_weights(), self.cqt.kernel_plan(), _lib.CQT_LOGPOW, eps=1e-9,
                                 packed_filters=self.cqt.tensor_core_filters())
        x = _run_modules(self.module_list, x)
        return x.squeeze(2)


default_encoder_block_dict = {'in_channels': 64, 'hidden_channels': None, 'out_channels': 64,
                              'kernel_size_1': (3, 3), 'kernel_size_2': (3, 3),
                              'top_padding_1': None, 'top_padding_2': None,
                              'padding_1': 0, 'padding_2': 0, 'stride_1': 1, 'stride_2': 1,
                              'pooling_1': 1, 'pooling_2': 1, 'bias': True, 'separable': False,
                              'residual': True, 'batch_norm': False, 'ceil_pooling': False}


class ScalogramEncoderBlock(nn.Module):
    """scalogram_model.py:372-479.

        +--------------- pooling -- conv_1x1 ------------------+
        |                                                      |
      --+-- [pad] conv_a [bn] [pool] relu [pad] conv_b [bn] [pool] relu --+--
    """

    def __init__(self, args_dict=default_encoder_block_dict, name='scalogram_block', activation_register=None):
        super().__init__()
        self.name = name
        if args_dict['hidden_channels'] is None:
            args_dict['hidden_channels'] = args_dict['out_channels']
        conv_module = Conv2dSeparable if args_dict['separable'] else Conv2d
        ceil_pooling = args_dict.get('ceil_pooling', False)
        self.main_modules = nn.ModuleList()
        stages = (('1', args_dict['in_channels'], args_dict['hidden_channels']),
                  ('2', args_dict['hidden_channels'], args_dict['out_channels']))
        for tag, c_in, c_out in stages:
            if args_dict['top_padding_' + tag] is not None:
                self.main_modules.append(nn.ZeroPad2d((0, 0, args_dict['top_padding_' + tag], 0)))
            self.main_modules.append(conv_module(in_channels=c_in, out_channels=c_out,
                                                 kernel_size=args_dict['kernel_size_' + tag], bias=args_dict['bias'],
                                                 padding=args_dict['padding_' + tag],
                                                 stride=args_dict['stride_' + tag]))
            if args_dict['batch_norm']:
                self.main_modules.append(nn.BatchNorm2d(c_out))
            if args_dict['pooling_' + tag] > 1:
                self.main_modules.append(MaxPool2d(kernel_size=args_dict['pooling_' + tag], ceil_mode=ceil_pooling))
            self.main_modules.append(nn.ReLU())
            self.main_modules.append(ActivationWriter(register=activation_register,
                                                      name=self.name + '_main_conv_' + tag))
        self.residual = args_dict['residual']
        if self.residual:
            self.residual_modules = nn.ModuleList()
            stride_pool = args_dict['stride_1'] * args_dict['stride_2'] * args_dict['pooling_1'] * args_dict['pooling_2']
            if stride_pool > 1:
                self.residual_modules.append(MaxPool2d(kernel_size=stride_pool, ceil_mode=True))
            if args_dict['in_channels'] != args_dict['out_channels']:
                self.residual_modules.append(Conv2d(args_dict['in_channels'], args_dict['out_channels'], 1,
                                                    padding=args_dict['padding_1'] + args_dict['padding_2'],
                                                    bias=False))
        self.output_activation_writer = ActivationWriter(register=activation_register,
                                                         name=self.name + '_main_conv_2')

    def _taps_attached(self):
        return any(isinstance(m, ActivationWriter) and m.register is not None for m in self.main_modules) or \
            self.output_activation_writer.register is not None

    def _residual_branch(self, x, main_shape, pooled=None, crop_early=False):
        """Residual branch output and the crop origin that centre-aligns it with ``main`` (:456-470).  ``pooled``: output
        of the branch's leading MaxPool2d when the caller already has it (ops.conv2d_with_pool)."""
        mods = list(self.residual_modules)
        if pooled is not None:
            x, mods = pooled, mods[1:]
        m_h, m_w = main_shape
        last = mods[-1] if mods else None
        if (crop_early and not ops.switch("CPC_NO_EARLY_CROP") and type(last) is Conv2d and last.supported()
                and tuple(last.kernel_size) == (1, 1) and tuple(last.stride) == (1, 1) and tuple(last.padding) == (0, 0)):
            # A pointwise conv commutes with the centre crop: crop its input, so that it (and its backward) only works on
            # the part of the branch the block output reads (34 of 64 rows in block 1 of architecture 7, 2 of 17 in block 2)
            pre = _run_modules(mods[:-1], x)
            off_h, off_w = self._crop_origin(pre.shape, main_shape)
            pre = pre[:, :, off_h:off_h + m_h, off_w:off_w + m_w]
            return last(pre.contiguous()), (0, 0)
        res = _run_modules(mods, x)
        off_h, off_w = self._crop_origin(res.shape, main_shape)
        return res, (off_h, off_w)

    @staticmethod
    def _crop_origin(res_shape, main_shape):
        """Origin of the centre crop of the residual branch, with the reference's rounding (:456-470)."""
        r_h, r_w = res_shape[2], res_shape[3]
        m_h, m_w = main_shape
        o_h = int((r_h - m_h + 1) / 2)
        o_w = int((r_w - m_w + 1) / 2)
        off_h = r_h - (o_h + m_h) if o_h > 0 else 0
        off_w = r_w - (o_w + m_w) if o_w > 0 else 0
        if off_h < 0 or off_w < 0 or off_h + m_h > r_h or off_w + m_w > r_w:
            raise ValueError("residual branch %s cannot be cropped to %s" % (tuple(res_shape), (m_h, m_w)))
        return off_h, off_w

    def _forward_fused(self, x, outer_relu):
        """conv -> fused(BN, ReLU) -> conv -> fused(BN, ReLU, + residual crop, [ReLU]); used when every stage is
        ``[pad] conv bn relu`` and no activation tap is attached."""
        mods = [m for m in self.main_modules if not isinstance(m, ActivationWriter)]
        stages, i = [], 0
        while i < len(mods):
            top = 0
            if isinstance(mods[i], nn.ZeroPad2d):
                left, right, top, bottom = mods[i].padding
                if left or right or bottom:
                    return None
                i += 1
            if i + 3 > len(mods):
                return None
            conv, bn, act = mods[i], mods[i + 1], mods[i + 2]
            if not (isinstance(conv, Conv2d) and isinstance(bn, nn.BatchNorm2d) and isinstance(act, nn.ReLU)):
                return None
            stages.append((top, conv, bn))
            i += 3
        if len(stages) != 2:
            return None
        (top_a, conv_a, bn_a), (top_b, conv_b, bn_b) = stages
        pooled = None
        pool = self.residual_modules[0] if self.residual and len(self.residual_modules) else None
        if (isinstance(pool, MaxPool2d) and pool.runs_on_kernels(x) and type(conv_a) is Conv2d and conv_a.supported()
                and x.requires_grad and torch.is_grad_enabled() and not ops.switch("CPC_NO_CONV_POOL_NODE")):
            # x feeds conv_a and the residual pooling: one node, so that backward adds the two gradients of x in place
            y_a, pooled = ops.conv2d_with_pool(x, conv_a.weight, conv_a.bias, conv_a.stride, conv_a.padding, top_a,
                                               pool.kernel_size, pool.ceil_mode)
        else:
            y_a = conv_a(x, extra_top=top_a)
        if ops.block_tail_eligible(y_a, bn_a, conv_b, top_b, bn_b):
            # bn_a + ReLU + conv_b + bn_b (+ residual) + ReLU as one autograd node with packed intermediates
            oh = y_a.shape[2] + top_b + 2 * conv_b.padding[0] - conv_b.weight.shape[2] + 1
            res, off = (self._residual_branch(x, (oh, y_a.shape[3]), pooled, crop_early=True) if self.residual
                        else (None, (0, 0)))
            return ops.block_tail(y_a, bn_a, conv_b, top_b, bn_b, residual=res, res_off=off, outer_relu=outer_relu)
        h = x
        for idx, (top, conv, bn) in enumerate(stages):
            h = y_a if idx == 0 else conv(h, extra_top=top)
            if idx == 0 or not self.residual:
                h = ops.bn_relu(h, bn, relu=True, outer_relu=False)
                if idx == 1 and outer_relu:
                    h = F.relu(h)
            else:
                res, off = self._residual_branch(x, (h.shape[2], h.shape[3]), pooled, crop_early=True)
                h = ops.bn_relu(h, bn, residual=res, res_off=off, relu=True, outer_relu=outer_relu)
        return h

    def forward(self, x, outer_relu=False):
        """``outer_relu``: also apply the ReLU the encoder puts between blocks (scalogram_model.py:523-527)."""
        if not self._taps_attached() and not ops.second_order_enabled():
            fused = self._forward_fused(x, outer_relu)
            if fused is not None:
                return fused
        main = _run_modules(self.main_modules, x)
        if self.residual:
            res, (off_h, off_w) = self._residual_branch(x, (main.shape[2], main.shape[3]))
            main = main + res[:, :, off_h:off_h + main.shape[2], off_w:off_w + main.shape[3]]
        self.output_activation_writer(main)
        return F.relu(main) if outer_relu else main


class ScalogramResidualEncoder(nn.Module):
    """scalogram_model.py:488-529: stack of blocks with ReLU in between; returns x[:, :, 0, :]."""

    def __init__(self, args_dict=None, preprocessing_module=None, verbose=0):
        super().__init__()
        self.verbose = verbose
        self.phase = args_dict['phase']
        if self.phase:
            args_dict['blocks'][0]['in_channels'] = 2
        if preprocessing_module is None:
            self.receptive_field = 1
            self.downsampling_factor = 1
        else:
            self.receptive_field = preprocessing_module.receptive_field
            self.downsampling_factor = preprocessing_module.downsampling_factor
        self.blocks = nn.ModuleList()
        for i, block_dict in enumerate(args_dict['blocks']):
            self.blocks.append(ScalogramEncoderBlock(block_dict, name='scalogram_block_' + str(i),
                                                     activation_register=args_dict.get('activation_register')))
            self.receptive_field += (block_dict['kernel_size_1'][1] - 1) * self.downsampling_factor
            self.downsampling_factor *= block_dict['pooling_1'] * block_dict['stride_1']
            self.receptive_field += (block_dict['kernel_size_2'][1] - 1) * self.downsampling_factor
            self.downsampling_factor *= block_dict['pooling_2'] * block_dict['stride_2']
            if self.verbose > 0:
                print("receptive field after block", i, ":", self.receptive_field)

    def forward(self, x):
        if x.dim() == 3:
            x = x.unsqueeze(2)
        last = len(self.blocks) - 1
        for i, block in enumerate(self.blocks):
            x = block(x, outer_relu=i < last)
            if self.verbose > 1:
                print("activation shape after block", i, ":", x.shape)
        return x[:, :, 0, :]
