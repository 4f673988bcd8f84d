"""The experiment configurations the benchmark and tests run, as plain dict literals with the reference's
keys (configs/cqt_configs.py, scalogram_resnet_configs.py, autoregressive_model_configs.py,
contrastive_estimation_configs.py), plus ``setup_model`` (setup_functions.py:68-117).

Values are the ones the reference ends up with *as imported* (its configs alias and mutate shared dicts;
oracle/make_golden.py dumps them from the reference into tests/golden/configs.json and
tests/test_host_logic.py compares).  The reference's own ``configs/`` package loads unchanged against this
package through the module shims in ``compat/``: tests/test_compat_configs.py builds every experiment
(e0 ... e32) that way and compares item lengths and state_dict keys with the reference's own setup_model.
"""
import copy

import torch

from .ar_models import AttentionModel, ConvolutionalArModel
from .encoders import ScalogramResidualEncoder
from .frontend import PreprocessingModule
from .model import AudioPredictiveCodingModel
from .trainer import linear_score_function, softplus_score_function

cqt_default_dict = {'sample_rate': 16000, 'fmin': 30, 'n_bins': 256, 'bins_per_octave': 32, 'filter_scale': 0.5,
                    'hop_length': 128, 'trainable_cqt': False}
cqt_high_res_dict = dict(cqt_default_dict, sample_rate=44100, n_bins=292, hop_length=256)


def _block(in_channels, out_channels, kernel_size_1=(3, 3), kernel_size_2=(3, 3), **overrides):
    block = {'in_channels': in_channels, 'hidden_channels': None, 'out_channels': out_channels,
             'kernel_size_1': kernel_size_1, 'kernel_size_2': kernel_size_2,
             'top_padding_1': None, 'top_padding_2': None, 'padding_1': 0, 'padding_2': 0,
             'stride_1': 1, 'stride_2': 1, 'pooling_1': 1, 'pooling_2': 1, 'bias': True, 'separable': False,
             'residual': True, 'batch_norm': False, 'ceil_pooling': False}
    block.update(overrides)
    return block


def _resnet(blocks, **overrides):
    cfg = {'model': ScalogramResidualEncoder, 'phase': True, 'scalogram_offset_zero': False,
           'scalogram_output_power': 1., 'scalogram_scaling': 1., 'scalogram_pooling': None,
           'blocks': blocks, 'activation_register': None}
    cfg.update(overrides)
    return cfg


def scalogram_resnet_architecture_7():
    """Strided 3x3 convs alternating with tall pitch convs (64x1 / 30x1 / 15x1); 512-d code, hop 1024.
    (2,256,629) -> (32,127,314) -> (128,34,156) -> (256,2,77) -> (512,1,76).  Block 3 has no batch norm in
    the as-imported reference (a later config mutates the shared dict)."""
    return _resnet([
        _block(1, 32, kernel_size_2=(64, 1), top_padding_2=63, stride_1=2, batch_norm=True),
        _block(32, 128, kernel_size_2=(30, 1), stride_1=2, batch_norm=True),
        _block(128, 256, kernel_size_2=(15, 1), stride_1=2, batch_norm=True),
        _block(256, 512, kernel_size_1=(2, 2), kernel_size_2=(1, 1), batch_norm=False),
    ])


def ar_conv_architecture_3():
    return {'model': ConvolutionalArModel, 'kernel_sizes': [5] * 6,
            'channel_count': [512, 512, 512, 256, 256, 256, 256], 'stride': [1] * 6, 'pooling': [1, 1, 2, 1, 2, 1],
            'bias': True, 'batch_norm': True, 'residual': True, 'encoding_size': 512, 'ar_code_size': 256,
            'activation_register': None, 'self_attention': [False] * 6}


def attention_architecture_1():
    return {'model': AttentionModel, 'channels': 512, 'output_size': 256, 'num_layers': 3, 'num_heads': 8,
            'feedforward_size': 512, 'sequence_length': 60, 'dropout': 0.1, 'encoding_size': 512,
            'ar_code_size': 256}


def attention_architecture_2():
    return dict(attention_architecture_1(), num_layers=6, feedforward_size=2048)


def contrastive_estimation_default():
    return {'regularization': 0.01, 'prediction_noise': 0., 'optimizer': torch.optim.Adam, 'file_batch_size': 1,
            'score_over_all_timesteps': True, 'log_interval': 20, 'validation_interval': 1000,
            'snapshot_interval': 5000, 'train_batch_size': 64, 'validate_batch_size': 64,
            'max_validation_steps': 300, 'learning_rate': 1e-4, 'max_epochs': 100, 'visible_steps': 60,
            'prediction_steps': 16, 'score_function': softplus_score_function,
            'wasserstein_gradient_penalty': False, 'gradient_penalty_factor': 10., 'trace_model': False,
            'use_all_GPUs': True}


def _linear_training(**overrides):
    return dict(contrastive_estimation_default(), score_function=linear_score_function, regularization=0.,
                file_batch_size=8, **overrides)


def experiment(name):
    """'e24' (BASELINE config 2: arch 7 + conv AR, B=64), 'e25' (per-step variant, B=32),
    'e20' (arch 7 + attention AR; without the gradient penalty the B200 path does not implement yet)."""
    if name == 'e24':
        return {'cqt_config': dict(cqt_default_dict), 'encoder_config': scalogram_resnet_architecture_7(),
                'ar_model_config': ar_conv_architecture_3(), 'training_config': _linear_training()}
    if name == 'e25':
        return {'cqt_config': dict(cqt_default_dict), 'encoder_config': scalogram_resnet_architecture_7(),
                'ar_model_config': ar_conv_architecture_3(),
                'training_config': _linear_training(train_batch_size=32, score_over_all_timesteps=False)}
    if name == 'e20':
        return {'cqt_config': dict(cqt_default_dict), 'encoder_config': scalogram_resnet_architecture_7(),
                'ar_model_config': attention_architecture_1(),
                'training_config': _linear_training(train_batch_size=32)}
    raise KeyError(name)


def experiment_from_plain(plain):
    """An experiment dict whose classes / functions are given by NAME (tests/golden/configs.json: the reference's
    experiment dicts as imported, dumped by oracle/make_golden.py) -> the same dict with cpc_b200 objects."""
    from . import encoders, ar_models, trainer
    names = {"ScalogramResidualEncoder": encoders.ScalogramResidualEncoder, "ScalogramEncoder": encoders.ScalogramEncoder,
             "AudioEncoder": encoders.AudioEncoder, "ConvolutionalArModel": ar_models.ConvolutionalArModel,
             "AttentionModel": ar_models.AttentionModel, "AudioGRUModel": ar_models.AudioGRUModel,
             "Adam": torch.optim.Adam, "SGD": torch.optim.SGD,
             "linear_score_function": trainer.linear_score_function,
             "softplus_score_function": trainer.softplus_score_function,
             "difference_score_function": trainer.difference_score_function}

    def convert(key, value):
        if isinstance(value, dict):
            return {k: convert(k, v) for k, v in value.items()}
        if isinstance(value, list):
            value = [convert(key, v) for v in value]
            return tuple(value) if key.startswith("kernel_size") else value
        if key in ("model", "optimizer", "score_function") and isinstance(value, str):
            return names[value]
        return value

    return {k: convert(k, v) for k, v in plain.items()}


def setup_model(cqt_params=None, encoder_params=None, ar_params=None, trainer_args=None, device=None,
                visible_steps=60, prediction_steps=16, trace_model=False, use_all_GPUs=True,
                activation_register=None):
    """setup_functions.py:68-117 -> (pc_model, preprocessing_module, untraced_model).  Multi-GPU is one
    process per GPU here (see ddp.py), so no DataParallel wrapping happens; tracing is not supported."""
    cqt_params = dict(cqt_default_dict) if cqt_params is None else cqt_params
    encoder_params = scalogram_resnet_architecture_7() if encoder_params is None else encoder_params
    ar_params = ar_conv_architecture_3() if ar_params is None else ar_params
    trainer_args = contrastive_estimation_default() if trainer_args is None else trainer_args
    encoder_params['activation_register'] = activation_register
    ar_params['activation_register'] = activation_register
    preprocessing_module = PreprocessingModule(cqt_dict=cqt_params, phase=encoder_params['phase'],
                                               offset_zero=encoder_params['scalogram_offset_zero'],
                                               output_power=encoder_params['scalogram_output_power'],
                                               pooling=encoder_params['scalogram_pooling'],
                                               scaling=encoder_params['scalogram_scaling'])
    encoder = encoder_params['model'](args_dict=encoder_params, preprocessing_module=preprocessing_module)
    ar_model = ar_params['model'](args_dict=ar_params)
    pc_model = AudioPredictiveCodingModel(encoder=encoder, autoregressive_model=ar_model,
                                          enc_size=ar_params['encoding_size'], ar_size=ar_params['ar_code_size'],
                                          visible_steps=trainer_args['visible_steps'],
                                          prediction_steps=trainer_args['prediction_steps'],
                                          activation_register=activation_register)
    if trainer_args.get('trace_model'):
        raise NotImplementedError("torch.jit.trace of the custom-kernel model is not supported")
    if device is not None:
        pc_model = pc_model.to(device)
        preprocessing_module = preprocessing_module.to(device)
    return pc_model, preprocessing_module, pc_model
