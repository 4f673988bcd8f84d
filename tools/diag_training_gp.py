"""Diagnostic: first-step gradient error of the gradient-penalty training replay vs the reference golden,
per CQT kernel path (CPC_NO_TENSOR_CQT) and per conv kernel path (CPC_FORCE_CUDA_CORE_CONV)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "constrastive-predictive-coding-audio_b200")]
import torch
import conftest
import test_gpu_parity as T
import cpc_b200 as cpc

full = conftest.load_golden("trainer_gp.npz")
for tag, all_steps in (("a", True), ("p", False)):
    g = {k[len(tag) + 1:]: v for k, v in full.items() if k.startswith(tag + ".")}
    for cqt_flag, conv_flag in (("0", "0"), ("1", "0"), ("1", "1")):
        os.environ["CPC_NO_TENSOR_CQT"] = cqt_flag
        os.environ["CPC_FORCE_CUDA_CORE_CONV"] = conv_flag
        cfg = T.small_resnet_cfg()
        cfg['blocks'][2] = dict(cfg['blocks'][2], kernel_size_1=(30, 2), pooling_1=1, ceil_pooling=False)
        cfg['blocks'][1] = dict(cfg['blocks'][1], kernel_size_2=(35, 1))
        pre = cpc.PreprocessingModule(dict(cpc.cqt_default_dict), phase=True)
        enc = cpc.ScalogramResidualEncoder(cfg, preprocessing_module=pre)
        ar = cpc.ConvolutionalArModel({'kernel_sizes': [3, 3], 'channel_count': [24, 16, 16], 'stride': [1, 1],
                                       'pooling': [1, 2], 'bias': True, 'batch_norm': True, 'residual': False,
                                       'activation_register': None})
        model = cpc.AudioPredictiveCodingModel(enc, ar, enc_size=24, ar_size=16, visible_steps=10, prediction_steps=3)
        fn = cpc.linear_score_function if all_steps else cpc.softplus_score_function
        log, snaps, lr = T._replay_trainer(cpc, g, model, pre, seed=5, steps=2, regularization=0.25,
                                           score_over_all_timesteps=all_steps, score_function=fn, prediction_steps=3,
                                           wasserstein_gradient_penalty=True, gradient_penalty_factor=10.)
        print("[%s] NO_TENSOR_CQT=%s CUDA_CORE_CONV=%s losses ours %s ref %s" % (tag, cqt_flag, conv_flag, log.l, list(g["losses"])))
        noise_only = conftest.bn_shadowed_biases(snaps[0].keys())
        errs = []
        for k, after in snaps[0].items():
            if not after.dtype.is_floating_point or k.endswith("running_mean") or k.endswith("running_var") or k in noise_only:
                continue
            before = torch.from_numpy(g["s0." + k])
            errs.append((conftest.grad_err((before - after) / lr, (before - torch.from_numpy(g["s1." + k])) / lr), k))
        errs.sort(reverse=True)
        print("   first-step grad err, worst 6:", [(round(e, 5), k) for e, k in errs[:6]])
