"""Model composition: ``AudioPredictiveCodingModel`` and the activation taps.

Same public surface as audio_model.py:164-284 of the reference.  The composition itself is plain Python
(slicing, one Linear = the K stacked predictors W_k); the encoder inside runs on the B200 kernels and the
autoregressive model is whatever ``nn.Module`` the caller plugs in.
"""
import contextlib
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import ops


class ActivationRegister:
    """Collects named intermediate activations for visualisation (audio_model.py:222-272)."""

    def __init__(self, writing_condition=None, clone_activations=False, batch_filter=None, move_to_cpu=False,
                 devices=None):
        self.devices = devices
        self.activations = OrderedDict() if devices is None else {dev: OrderedDict() for dev in devices}
        self.active = True
        self.writing_condition = writing_condition
        self.clone_activations = clone_activations
        self.batch_filter = batch_filter
        self.move_to_cpu = move_to_cpu

    def write_activation(self, name, value):
        if not self.active:
            return
        if self.writing_condition is not None and not self.writing_condition(value):
            return
        if self.batch_filter is not None:
            value = value[self.batch_filter]
        dev = value.device.index if self.devices is not None else None
        if self.move_to_cpu:
            value = value.cpu()
        if self.clone_activations:
            value = value.clone()
        if self.devices is not None:
            self.activations[dev][name] = value
        else:
            self.activations[name] = value

    def get_activations(self):
        if self.devices is None:
            return self.activations
        first = self.devices[0]
        return {key: torch.cat([self.activations[dev][key].to(self.activations[first][key].device)
                                for dev in self.devices], dim=0)
                for key in self.activations[first].keys()}


class ActivationWriter(nn.Module):
    def __init__(self, register, name):
        super().__init__()
        self.register = register
        self.name = name

    def forward(self, x):
        if self.register is not None:
            self.register.write_activation(self.name, x)
        return x


def cuda0_writing_condition(x):
    return x.device.index == 0 if x.device.type == 'cuda' else True


def num_parameters(model):
    return sum(int(np.prod(p.shape)) for p in model.parameters())


def load_to_cpu(path):
    model = torch.load(path, map_location=lambda storage, loc: storage, weights_only=False)
    model.cpu()
    return model


class AudioPredictiveCodingModel(nn.Module):
    """encoder -> (targets, visible z) -> autoregressive model -> K linear predictors.

    forward(x) -> (predicted_z (B,K,E), targets (B,E,K) [a strided, non-detached view of the encoder
    output], z (B,E,V), c (B,A)); audio_model.py:193-213."""

    def __init__(self, encoder, autoregressive_model, enc_size, ar_size, visible_steps=100, prediction_steps=12,
                 activation_register=None):
        super().__init__()
        self.enc_size = enc_size
        self.ar_size = ar_size
        self.visible_steps = visible_steps
        self.prediction_steps = prediction_steps
        self.encoder = encoder
        self.autoregressive_model = autoregressive_model
        self.prediction_model = nn.Linear(in_features=ar_size, out_features=enc_size * prediction_steps, bias=False)
        self.activation_register = activation_register
        self.input_activation_writer = ActivationWriter(activation_register, 'scalogram')
        self.z_activation_writer = ActivationWriter(activation_register, 'z_code')
        self.c_activation_writer = ActivationWriter(activation_register, 'c_code')
        self.prediction_activation_writer = ActivationWriter(activation_register, 'prediction')

    @property
    def item_length(self):
        steps = self.visible_steps + self.prediction_steps
        return self.encoder.receptive_field + steps * self.encoder.downsampling_factor

    def forward(self, x):
        x = self.input_activation_writer(x)
        code = self.encoder(x)
        k, v = self.prediction_steps, self.visible_steps
        targets = code[:, :, -k:]
        z = self.z_activation_writer(code[:, :, -(v + k):-k])
        # bf16 operand mode (ops.set_default_precision("bf16"), BASELINE configs[2]): the stock-PyTorch parts -- the AR
        # model and the predictors W_k -- run under bf16 autocast (library GEMMs with bf16 operands and fp32 accumulation,
        # softmax / layer norm in fp32), like the conv kernels of that mode; results return to fp32 for the scoring.
        bf16 = z.is_cuda and ops.get_default_precision() == "bf16"
        with (torch.autocast("cuda", dtype=torch.bfloat16) if bf16 else contextlib.nullcontext()):
            c = self.autoregressive_model(z)
            if c.dim() == 3:
                c = c[:, :, 0]
            predicted_z = self.prediction_model(c)
        if bf16:
            c, predicted_z = c.float(), predicted_z.float()
        c = self.c_activation_writer(c)
        predicted_z = self.prediction_activation_writer(predicted_z.view(-1, k, self.enc_size))
        return predicted_z, targets, z, c

    def parameter_count(self):
        return num_parameters(self)
