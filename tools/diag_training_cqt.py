"""Diagnostic: parameter drift of the CQT+resnet training replay vs the reference golden, per CQT kernel path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "constrastive-predictive-coding-audio_b200")]
import torch
import conftest
import test_gpu_parity as T
import cpc_b200 as cpc

for flag in ("1", "0"):
    os.environ["CPC_NO_TENSOR_CQT"] = flag
    g = conftest.load_golden("trainer_cqt.npz")
    cfg = T.small_resnet_cfg()
    cfg['blocks'][2] = dict(cfg['blocks'][2], kernel_size_1=(30, 2), pooling_1=1, ceil_pooling=False)
    cfg['blocks'][1] = dict(cfg['blocks'][1], kernel_size_2=(35, 1))
    pre = cpc.PreprocessingModule(dict(cpc.cqt_default_dict), phase=True)
    enc = cpc.ScalogramResidualEncoder(cfg, preprocessing_module=pre)
    ar = cpc.ConvolutionalArModel({'kernel_sizes': [3, 3], 'channel_count': [24, 16, 16], 'stride': [1, 1],
                                   'pooling': [1, 2], 'bias': True, 'batch_norm': True, 'residual': False,
                                   'activation_register': None})
    model = cpc.AudioPredictiveCodingModel(enc, ar, enc_size=24, ar_size=16, visible_steps=10, prediction_steps=3)
    log, snaps, lr = T._replay_trainer(cpc, g, model, pre, seed=3, steps=3, regularization=0.25, score_over_all_timesteps=True,
                                       score_function=cpc.linear_score_function, prediction_steps=3)
    print("NO_TENSOR_CQT=%s losses ours %s ref %s" % (flag, log.l, list(g["losses"])))
    worst = []
    for k, after in snaps[-1].items():
        if after.dtype.is_floating_point:
            worst.append((conftest.rel_err(after, g["s%d.%s" % (len(snaps), k)]), k))
    worst.sort(reverse=True)
    print("  final-snapshot rel err, worst 8:", [(round(e, 5), k) for e, k in worst[:8]])
    upd = []
    for k, after in snaps[-1].items():
        if after.dtype.is_floating_point and not k.endswith("running_mean") and not k.endswith("running_var"):
            s0 = torch.from_numpy(g["s0." + k]); want = torch.from_numpy(g["s%d.%s" % (len(snaps), k)])
            upd.append((conftest.rel_err(after - s0, want - s0), k))
    upd.sort(reverse=True)
    print("  accumulated-update rel err, worst 8:", [(round(e, 5), k) for e, k in upd[:8]])
    # scalogram difference vs oracle for the first batch
    x = torch.from_numpy(g["items"][:4]).unsqueeze(1).to("cuda:0")
    y = pre.to("cuda:0")(x)
    print("  scalogram mean/abs", float(y[:, 0].mean()), float(y[:, 1].abs().mean()))
    torch.save(y.cpu(), "/tmp/scal_%s.pt" % flag)
a, b = torch.load("/tmp/scal_1.pt"), torch.load("/tmp/scal_0.pt")
print("amp rel diff", float((a[:, 0] - b[:, 0]).norm() / a[:, 0].norm()), "phase max abs diff", float((a[:, 1] - b[:, 1]).abs().max()),
      "phase mean abs diff", float((a[:, 1] - b[:, 1]).abs().mean()))
