"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):   python oracle/make_golden.py
The reference is imported through oracle/ref_shim.py (in-memory stand-ins for librosa / mutagen /
ml_utilities / matplotlib).  Everything saved here is an output of the reference's own code:

  cqt.npz          CQT.forward and three PreprocessingModule variants on seeded noise (+ a silent segment)
  audio_encoder.npz AudioEncoder forward + parameter/input gradients of sum(out * g)
  resnet_encoder.npz a small ScalogramResidualEncoder (BN, top padding, strides, residual crop) fwd + grads
  trainer_raw.npz  ContrastiveEstimationTrainer.train() itself, 2 SGD steps, raw-wave model (config-1 shaped,
                   shrunk): per-step loss / max score from the logger, and parameters before/after, so
                   (before - after) / lr is the reference's own gradient of its own loss
  trainer_cqt.npz  same through PreprocessingModule + ScalogramResidualEncoder + ConvolutionalArModel,
                   all-steps linear scoring with regulariser
  infonce.npz      score functions + the trainer's loss on random (pred, targets): obtained by running
                   train() on an identity "model" that returns the stored tensors (so it is still the
                   reference's code computing loss and gradients)
  sampler.npz      FileBatchSampler index streams for several (counts, batch, file_batch, seed) settings
  configs.json     as-imported experiment dicts e24 / e25 / e20 (classes replaced by their names)
  e24_step.npz     BASELINE configs[1] at full item length: setup_model(experiments['e24']) trained for two Adam steps by
                   the reference's train() (losses, encoder output, gradient / parameter subsamples)
"""
import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class ListDataset(torch.utils.data.Dataset):
    def __init__(self, items, files=1):
        self.items, self.files = items, files

    def __len__(self):
        return self.items.shape[0]

    def __getitem__(self, i):
        return self.items[i]

    def get_example_count_per_file(self):
        base, extra = divmod(len(self), self.files)
        return [base + (1 if i < extra else 0) for i in range(self.files)]


class CaptureLogger:
    def __init__(self):
        self.losses, self.scores = [], []
        outer = self

        class M:
            def __init__(self, sink):
                self.sink = sink

            def update(self, v, n=1):
                self.sink.append(float(v))

        self.loss_meter, self.score_meter = M(outer.losses), M(outer.scores)

    def log(self, step):
        pass


def sd_np(module, prefix=""):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def golden_cqt(ref):
    sm = ref["scalogram_model"]
    cq = ref["constant_q_transform"]
    g = torch.Generator().manual_seed(1234)
    x = 0.1 * torch.randn(2, 1, 16384 + 1 + 128 * 9 + 37, generator=g)
    x[1, 0, 3000:9000] = 0.0                        # silent stretch: exercises the eps floor in high bins
    cqt = cq.CQT(sr=16000, fmin=30, n_bins=256, bins_per_octave=32, filter_scale=0.5, hop_length=128)
    out = {"x": x.numpy(), "complex": cqt(x).numpy()}
    d = dict(sm.cqt_default_dict)
    out["logpow"] = sm.PreprocessingModule(d, phase=False)(x).numpy()
    out["logpow_phase"] = sm.PreprocessingModule(d, phase=True)(x).numpy()
    out["offset_pool_power"] = sm.PreprocessingModule(d, phase=False, offset_zero=True, output_power=2.,
                                                      pooling=[1, 2], scaling=10.)(x).numpy()
    out["phase_offset_pool"] = sm.PreprocessingModule(d, phase=True, offset_zero=True, pooling=[1, 2])(x).numpy()
    # a second, smaller filterbank (different grouping): filter_scale 1, 24 bins/octave, hop 64
    cqt2 = cq.CQT(sr=8000, fmin=55, n_bins=120, bins_per_octave=24, filter_scale=1., hop_length=64)
    x2 = 0.1 * torch.randn(3, 1, cqt2.conv_kernel_sizes[0] + 1 + 64 * 5, generator=g)
    out["x2"] = x2.numpy()
    out["complex2"] = cqt2(x2).numpy()
    out["kernel_sizes2"] = np.array(cqt2.conv_kernel_sizes)
    np.savez_compressed(os.path.join(OUT, "cqt.npz"), **out)


def golden_cqt_high_res(ref):
    """The high-resolution filterbank of experiments e27 ... e32 (configs/cqt_configs.py:9-12: 44.1 kHz, 292 bins, hop 256;
    11 octave groups from 65 536 taps down to 64): CQT.forward and the PreprocessingModule variants those experiments use
    (phase scalogram; offset_zero + pooling [1, 2] of scalogram_resnet_architecture_8 / 9), on a few frames."""
    import importlib
    sm = ref["scalogram_model"]
    cq = ref["constant_q_transform"]
    cc = importlib.import_module("configs.cqt_configs")
    d = dict(cc.cqt_high_res_dict)
    g = torch.Generator().manual_seed(2024)
    cqt = cq.CQT(sr=d['sample_rate'], fmin=d['fmin'], n_bins=d['n_bins'], bins_per_octave=d['bins_per_octave'],
                 filter_scale=d['filter_scale'], hop_length=d['hop_length'])
    x = 0.1 * torch.randn(2, 1, cqt.conv_kernel_sizes[0] + 1 + d['hop_length'] * 6 + 17, generator=g)
    out = {"x": x.numpy(), "kernel_sizes": np.array(cqt.conv_kernel_sizes),
           "ranges": np.array([[r.start, r.stop] for r in cqt.conv_index_ranges]),
           "cfg": np.array(json.dumps({k: v for k, v in d.items()})), "complex": cqt(x).numpy()}
    out["phase"] = sm.PreprocessingModule(d, phase=True)(x).numpy()
    out["offset_pool"] = sm.PreprocessingModule(d, phase=False, offset_zero=True, pooling=[1, 2])(x).numpy()
    np.savez_compressed(os.path.join(OUT, "cqt_high_res.npz"), **out)
    print("high-res CQT: kernel sizes", cqt.conv_kernel_sizes, "frames", out["complex"].shape)


def golden_cqt_grad(ref):
    """Gradients THROUGH the front end: d/d(audio) and d/d(filterbank) of the trainable CQT and
    of the phase scalogram, plus InverseCQT and PhaseAccumulation outputs (constant_q_transform.py:155-260, 294-313).
    A small filterbank (8 kHz, 96 bins, hop 64) keeps the fixture small."""
    sm = ref["scalogram_model"]
    cq = ref["constant_q_transform"]
    g = torch.Generator().manual_seed(4321)
    kw = dict(sr=8000, fmin=55, n_bins=96, bins_per_octave=24, filter_scale=0.5, hop_length=64)
    cqt = cq.CQT(trainable=True, **kw)
    x = (0.1 * torch.randn(2, 1, cqt.conv_kernel_sizes[0] + 1 + 64 * 6 + 11, generator=g)).requires_grad_(True)
    z = cqt(x)
    gz = torch.randn(z.shape, generator=g)
    (z * gz).sum().backward()
    out = {"x": x.detach().numpy(), "z": z.detach().numpy(), "gz": gz.numpy(), "gx": x.grad.numpy(),
           "kernel_sizes": np.array(cqt.conv_kernel_sizes)}
    for i, conv in enumerate(cqt.conv_modules):
        out["gw%d" % i] = conv.weight.grad.numpy()
    # phase scalogram of a trainable front end: loss = <y, gy>, gradient w.r.t. the audio and the filters
    d = {'sample_rate': 8000, 'fmin': 55, 'n_bins': 96, 'bins_per_octave': 24, 'filter_scale': 0.5, 'hop_length': 64,
         'trainable_cqt': True}
    pre = sm.PreprocessingModule(d, phase=True, offset_zero=True, output_power=1., pooling=[1, 2], scaling=3.)
    x2 = (0.1 * torch.randn(2, 1, x.shape[2], generator=g)).requires_grad_(True)
    y = pre(x2)
    gy = torch.randn(y.shape, generator=g)
    (y * gy).sum().backward()
    out.update({"x2": x2.detach().numpy(), "y2": y.detach().numpy(), "gy2": gy.numpy(), "gx2": x2.grad.numpy()})
    for i, conv in enumerate(pre.cqt.conv_modules):
        out["g2w%d" % i] = conv.weight.grad.numpy()
    # InverseCQT.forward itself raises on this torch (view of a permuted tensor, constant_q_transform.py:253), so the
    # fixture applies the reference module's OWN ConvTranspose1d layers with that line's view replaced by reshape
    icqt = cq.InverseCQT(**kw)
    zi = torch.randn(2, 96, 5, 2, generator=g)
    result = 0
    for i, conv in enumerate(icqt.conv_modules):
        band = zi[:, icqt.conv_index_ranges[i]]
        n, p_, t, c = band.shape
        result = result + conv(band.permute(3, 0, 1, 2).reshape(c * n, p_, t))
    result = result.view(2, n, 2, -1)
    out["icqt_in"] = zi.numpy()
    out["icqt_out"] = torch.stack([result[0, :, 0] - result[1, :, 0], result[0, :, 1] + result[1, :, 1]],
                                  dim=2).detach().numpy()
    acc = cq.PhaseAccumulation(sr=8000, fmin=55, n_bins=96, bins_per_octave=24, hop_length=64)
    ph = torch.randn(1, 96, 7, generator=g)                     # start_phase is (1, F, 1): batch 1 only (:311)
    out["acc_in"] = ph.numpy()
    out["acc_out"] = acc(ph).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "cqt_grad.npz"), **out)


def golden_audio_encoder(ref):
    am = ref["audio_model"]
    torch.manual_seed(0)
    cfg = {'strides': [5, 4, 2, 2, 2], 'kernel_sizes': [10, 8, 4, 4, 4], 'channel_count': [24, 32, 40, 32, 48],
           'bias': True}
    enc = am.AudioEncoder(cfg)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 1, 465 + 160 * 6 + 11, generator=g).requires_grad_(True)
    y = enc(x)
    gy = torch.randn(y.shape, generator=g)
    (y * gy).sum().backward()
    out = {"x": x.detach().numpy(), "y": y.detach().numpy(), "gy": gy.numpy(), "gx": x.grad.numpy()}
    out.update(sd_np(enc, "p."))
    out.update({"g." + n: p.grad.numpy() for n, p in enc.named_parameters()})
    # known-answer facts from tests/test_audioEncoder.py:19-48
    enc2 = am.AudioEncoder({'strides': [5, 4, 2, 2, 2], 'kernel_sizes': [10, 8, 4, 4, 4],
                            'channel_count': [32, 32, 32, 32, 32], 'bias': True})
    out["kat_shape"] = np.array(enc2(torch.zeros(7, 1, 4800)).shape)
    out["kat_rf_ds"] = np.array([enc2.receptive_field, enc2.downsampling_factor])
    np.savez_compressed(os.path.join(OUT, "audio_encoder.npz"), **out)


def small_resnet_cfg(ref):
    sm = ref["scalogram_model"]
    base = dict(sm.default_encoder_block_dict)
    base['ceil_pooling'] = False
    b0 = dict(base, in_channels=1, out_channels=8, kernel_size_2=(9, 1), top_padding_2=8, stride_1=2, batch_norm=True)
    b1 = dict(base, in_channels=8, out_channels=16, kernel_size_2=(6, 1), stride_1=2, batch_norm=True, padding_1=1)
    b2 = dict(base, in_channels=16, out_channels=24, kernel_size_1=(2, 2), kernel_size_2=(1, 1), pooling_1=2,
              ceil_pooling=True)
    return {'model': sm.ScalogramResidualEncoder, 'phase': True, 'scalogram_offset_zero': False,
            'scalogram_output_power': 1., 'scalogram_scaling': 1., 'scalogram_pooling': None,
            'blocks': [b0, b1, b2], 'activation_register': None}


def golden_resnet_encoder(ref):
    sm = ref["scalogram_model"]
    torch.manual_seed(1)
    cfg = small_resnet_cfg(ref)
    enc = sm.ScalogramResidualEncoder(cfg)
    enc.train()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 2, 40, 37, generator=g).requires_grad_(True)
    y = enc(x)
    gy = torch.randn(y.shape, generator=g)
    (y * gy).sum().backward()
    out = {"x": x.detach().numpy(), "y": y.detach().numpy(), "gy": gy.numpy(), "gx": x.grad.numpy()}
    out.update(sd_np(enc, "p."))           # includes BN running stats AFTER one training forward
    out.update({"g." + n: p.grad.numpy() for n, p in enc.named_parameters()})
    np.savez_compressed(os.path.join(OUT, "resnet_encoder.npz"), **out)


def run_reference_trainer(ref, model, dataset, preprocessing, steps, batch_size, lr, seed, **kw):
    cet = ref["contrastive_estimation_training"]
    logger = CaptureLogger()
    trainer = cet.ContrastiveEstimationTrainer(model=model, dataset=dataset, logger=logger, device=None,
                                               optimizer=torch.optim.SGD, preprocessing=preprocessing, **kw)
    random.seed(seed)
    snaps = [sd_np(model)]
    for s in range(steps):
        trainer.train(batch_size=batch_size, epochs=1, lr=lr, continue_training_at_step=s, num_workers=0,
                      max_steps=s + 1)
        snaps.append(sd_np(model))
    return logger, snaps


def golden_trainer_raw(ref):
    am = ref["audio_model"]
    torch.manual_seed(0)
    enc = am.AudioEncoder({'strides': [5, 4, 2, 2, 2], 'kernel_sizes': [10, 8, 4, 4, 4],
                           'channel_count': [16, 24, 24, 24, 32], 'bias': True})
    ar = am.AudioGRUModel(input_size=32, hidden_size=16)
    model = am.AudioPredictiveCodingModel(enc, ar, enc_size=32, ar_size=16, visible_steps=9, prediction_steps=4)
    g = torch.Generator().manual_seed(1234)
    items = 0.1 * torch.randn(16, model.item_length, generator=g)
    # NB: each train() call builds a fresh sampler and reshuffles; seed once, record the stream by replay
    logger, snaps = run_reference_trainer(ref, model, ListDataset(items), None, steps=2, batch_size=8, lr=0.5, seed=0,
                                          regularization=1., score_over_all_timesteps=False,
                                          score_function=ref["contrastive_estimation_training"].softplus_score_function,
                                          prediction_steps=4)
    out = {"items": items.numpy(), "losses": np.array(logger.losses), "max_scores": np.array(logger.scores),
           "lr": np.array(0.5), "batch_size": np.array(8)}
    for i, s in enumerate(snaps):
        out.update({"s%d.%s" % (i, k): v for k, v in s.items()})
    np.savez_compressed(os.path.join(OUT, "trainer_raw.npz"), **out)


def golden_trainer_cqt(ref):
    sm, am = ref["scalogram_model"], ref["audio_model"]
    cet = ref["contrastive_estimation_training"]
    torch.manual_seed(2)
    cfg = small_resnet_cfg(ref)
    # make the pitch axis collapse to 1: 256 bins -> s2 -> 127 -> (9x1, pad 8) 127 -> s2,p1 -> 64 -> (6x1) 59 ...
    cfg['blocks'][2] = dict(cfg['blocks'][2], kernel_size_1=(30, 2), pooling_1=1, ceil_pooling=False)
    cfg['blocks'][1] = dict(cfg['blocks'][1], kernel_size_2=(35, 1))
    pre = sm.PreprocessingModule(dict(sm.cqt_default_dict), phase=True)
    enc = sm.ScalogramResidualEncoder(cfg, preprocessing_module=pre)
    ar = am.ConvolutionalArModel({'kernel_sizes': [3, 3], 'channel_count': [24, 16, 16], 'stride': [1, 1],
                                  'pooling': [1, 2], 'bias': True, 'batch_norm': True, 'residual': False,
                                  'activation_register': None})
    model = am.AudioPredictiveCodingModel(enc, ar, enc_size=24, ar_size=16, visible_steps=10, prediction_steps=3)
    g = torch.Generator().manual_seed(99)
    items = 0.1 * torch.randn(8, model.item_length, generator=g)
    # lr small enough that the SGD trajectory is stable (at lr = 0.1 the loss jumps 3.3 -> 34 in one step and the
    # replay amplifies 1e-7 input differences by ~1e4, which tests chaos, not kernels)
    logger, snaps = run_reference_trainer(ref, model, ListDataset(items), pre, steps=3, batch_size=4, lr=0.003, seed=3,
                                          regularization=0.25, score_over_all_timesteps=True,
                                          score_function=cet.linear_score_function, prediction_steps=3)
    out = {"items": items.numpy(), "losses": np.array(logger.losses), "max_scores": np.array(logger.scores),
           "lr": np.array(0.003), "batch_size": np.array(4), "item_length": np.array(model.item_length)}
    for i, s in enumerate(snaps):
        out.update({"s%d.%s" % (i, k): v for k, v in s.items()})
    np.savez_compressed(os.path.join(OUT, "trainer_cqt.npz"), **out)


def golden_trainer_gp(ref):
    """Reference training steps with wasserstein_gradient_penalty=True (contrastive_estimation_training.py:144-155):
    the penalty differentiates d(sum scores)/d(scalogram) a second time through encoder and AR model."""
    sm, am = ref["scalogram_model"], ref["audio_model"]
    cet = ref["contrastive_estimation_training"]
    out = {}
    for tag, all_steps, fn in (("a", True, cet.linear_score_function), ("p", False, cet.softplus_score_function)):
        torch.manual_seed(4)
        cfg = small_resnet_cfg(ref)
        cfg['blocks'][2] = dict(cfg['blocks'][2], kernel_size_1=(30, 2), pooling_1=1, ceil_pooling=False)
        cfg['blocks'][1] = dict(cfg['blocks'][1], kernel_size_2=(35, 1))
        pre = sm.PreprocessingModule(dict(sm.cqt_default_dict), phase=True)
        enc = sm.ScalogramResidualEncoder(cfg, preprocessing_module=pre)
        ar = am.ConvolutionalArModel({'kernel_sizes': [3, 3], 'channel_count': [24, 16, 16], 'stride': [1, 1],
                                      'pooling': [1, 2], 'bias': True, 'batch_norm': True, 'residual': False,
                                      'activation_register': None})
        model = am.AudioPredictiveCodingModel(enc, ar, enc_size=24, ar_size=16, visible_steps=10, prediction_steps=3)
        g = torch.Generator().manual_seed(77)
        items = 0.1 * torch.randn(8, model.item_length, generator=g)
        lr = 1e-4
        logger, snaps = run_reference_trainer(ref, model, ListDataset(items), pre, steps=2, batch_size=4, lr=lr, seed=5,
                                              regularization=0.25, score_over_all_timesteps=all_steps,
                                              score_function=fn, prediction_steps=3,
                                              wasserstein_gradient_penalty=True, gradient_penalty_factor=10.)
        out.update({tag + ".items": items.numpy(), tag + ".losses": np.array(logger.losses),
                    tag + ".max_scores": np.array(logger.scores), tag + ".lr": np.array(lr),
                    tag + ".batch_size": np.array(4), tag + ".item_length": np.array(model.item_length)})
        for i, s in enumerate(snaps):
            out.update({"%s.s%d.%s" % (tag, i, k): v for k, v in s.items()})
    np.savez_compressed(os.path.join(OUT, "trainer_gp.npz"), **out)


def golden_scalogram_encoder(ref):
    """ScalogramEncoder (scalogram_model.py:129-227): own CQT -> log power (+ phase difference) ->
    pad / conv / pool / ReLU / BatchNorm stack.  Small stacks, both input variants; forward, parameter gradients."""
    sm = ref["scalogram_model"]
    out = {}
    for tag, phase, separable in (("m", False, False), ("p", True, False), ("s", False, True)):
        torch.manual_seed(8)
        cfg = dict(sm.cqt_default_dict)
        cfg.update({'kernel_sizes': [(9, 1), (5, 5), (5, 1), (3, 3)], 'top_padding': [8, 0, 0, 0],
                    'channel_count': [1, 8, 8, 16, 24], 'pooling': [1, 2, 1, 2], 'stride': [1, 1, 1, 1], 'bias': True,
                    'batch_norm': True, 'phase': phase, 'separable': separable, 'lowpass_init': 0., 'instance_norm': False,
                    'dropout': 0.})
        enc = sm.ScalogramEncoder(cfg)
        enc.train()
        g = torch.Generator().manual_seed(31)
        x = 0.1 * torch.randn(2, 1, 16384 + 1 + 128 * 40, generator=g)
        y = enc(x)
        gy = torch.randn(y.shape, generator=g)
        (y * gy).sum().backward()
        out.update({tag + ".x": x.numpy(), tag + ".y": y.detach().numpy(), tag + ".gy": gy.numpy(),
                    tag + ".rf": np.array(enc.receptive_field), tag + ".ds": np.array(enc.downsampling_factor)})
        out.update({tag + ".p." + k: v for k, v in sd_np(enc).items() if not k.startswith("cqt.")})
        out.update({tag + ".g." + n: p.grad.numpy() for n, p in enc.named_parameters() if p.grad is not None})
        # conditioning probe: the reference's own gradients when its CQT output moves by 1e-6 relative (the error level of
        # any independent fp32 evaluation of a 16 384-tap filter).  log / atan2 of near-silent cells and the ReLU gates behind
        # them amplify it; tests hold each gradient to max(1e-3, 3 x this self-noise)
        base = {n: p.grad.detach().clone() for n, p in enc.named_parameters() if p.grad is not None}
        noise_gen = torch.Generator().manual_seed(3)
        hook = enc.cqt.register_forward_hook(lambda m, i, o: o * (1 + 1e-6 * torch.randn(o.shape, generator=noise_gen)))
        enc.zero_grad()
        (enc(x) * gy).sum().backward()
        hook.remove()
        for n, p in enc.named_parameters():
            if p.grad is not None:
                out[tag + ".sn." + n] = np.array(float((p.grad - base[n]).norm() / base[n].norm().clamp_min(1e-30)))
    np.savez_compressed(os.path.join(OUT, "scalogram_encoder.npz"), **out)


def golden_snapshot(ref):
    """A whole-model snapshot pickle written the way the reference writes them (torch.save(model), SnapshotManager /
    audio_model.load_to_cpu): a small CQT + residual-encoder + conv-AR model, plus its state_dict for comparison."""
    sm, am = ref["scalogram_model"], ref["audio_model"]
    torch.manual_seed(6)
    cfg = small_resnet_cfg(ref)
    cfg['blocks'][2] = dict(cfg['blocks'][2], kernel_size_1=(30, 2), pooling_1=1, ceil_pooling=False)
    cfg['blocks'][1] = dict(cfg['blocks'][1], kernel_size_2=(35, 1))
    pre = sm.PreprocessingModule(dict(sm.cqt_default_dict), phase=True)
    enc = sm.ScalogramResidualEncoder(cfg, preprocessing_module=pre)
    ar = am.ConvolutionalArModel({'kernel_sizes': [3, 3], 'channel_count': [24, 16, 16], 'stride': [1, 1],
                                  'pooling': [1, 2], 'bias': True, 'batch_norm': True, 'residual': False,
                                  'activation_register': None})
    model = am.AudioPredictiveCodingModel(enc, ar, enc_size=24, ar_size=16, visible_steps=10, prediction_steps=3)
    torch.save(model, os.path.join(OUT, "ref_snapshot_1200"))
    np.savez_compressed(os.path.join(OUT, "ref_snapshot_state.npz"), **sd_np(model))


class StoredPairModel(torch.nn.Module):
    """A 'model' whose parameters ARE (pred, targets): lets the reference trainer's own loss code and
    autograd produce loss / gradients for arbitrary (pred, targets)."""

    def __init__(self, pred, targets):
        super().__init__()
        self.pred = torch.nn.Parameter(pred.clone())
        self.tgt = torch.nn.Parameter(targets.clone())

    def forward(self, batch):
        return self.pred, self.tgt, None, None


def golden_infonce(ref):
    cet = ref["contrastive_estimation_training"]
    out, cases = {}, []
    g = torch.Generator().manual_seed(5)
    idx = 0
    for (b, k, e) in [(8, 12, 64), (6, 4, 40), (16, 16, 96)]:
        for all_steps in (False, True):
            for kind in ("linear", "softplus"):
                for reg in (0.0, 0.7):
                    pred = torch.randn(b, k, e, generator=g) * (1.5 / e ** 0.5)
                    tgt = torch.randn(b, e, k, generator=g)
                    model = StoredPairModel(pred, tgt)
                    fn = cet.linear_score_function if kind == "linear" else cet.softplus_score_function
                    ds = ListDataset(torch.zeros(b, 4))
                    logger, snaps = run_reference_trainer(ref, model, ds, None, steps=1, batch_size=b, lr=1.0, seed=0,
                                                          regularization=reg, score_over_all_timesteps=all_steps,
                                                          score_function=fn, prediction_steps=k)
                    tag = "c%d" % idx
                    out[tag + ".pred"] = pred.numpy()
                    out[tag + ".tgt"] = tgt.numpy()
                    out[tag + ".loss"] = np.array(logger.losses[0])
                    out[tag + ".max"] = np.array(logger.scores[0])
                    out[tag + ".dpred"] = snaps[0]["pred"] - snaps[1]["pred"]
                    out[tag + ".dtgt"] = snaps[0]["tgt"] - snaps[1]["tgt"]
                    cases.append({"tag": tag, "b": b, "k": k, "e": e, "all_steps": all_steps, "kind": kind, "reg": reg})
                    idx += 1
    out["cases"] = np.array(json.dumps(cases))
    np.savez_compressed(os.path.join(OUT, "infonce.npz"), **out)


def golden_validate(ref):
    """The reference's own ContrastiveEstimationTrainer.validate() (contrastive_estimation_training.py:178-269) on
    stored (pred, targets) pairs -> per-step losses, accuracies, mean score, MI bound."""
    cet = ref["contrastive_estimation_training"]
    out, cases = {}, []
    g = torch.Generator().manual_seed(17)
    idx = 0
    for (b, k, e) in [(8, 12, 64), (16, 4, 40), (24, 16, 96)]:
        for all_steps in (False, True):
            for kind in ("linear", "softplus"):
                pred = torch.randn(b, k, e, generator=g) * (2.5 / e ** 0.5)
                tgt = torch.randn(b, e, k, generator=g)
                # make some predictions actually hit their own target so that accuracies are not all ~1/n
                for d in range(0, b, 2):
                    pred[d] = pred[d] + 0.6 * tgt[d].t()
                model = StoredPairModel(pred, tgt)
                fn = cet.linear_score_function if kind == "linear" else cet.softplus_score_function
                ds = ListDataset(torch.zeros(b, 4))
                trainer = cet.ContrastiveEstimationTrainer(model=model, dataset=ds, validation_set=ds, device=None,
                                                           score_over_all_timesteps=all_steps, score_function=fn,
                                                           prediction_steps=k)
                losses, acc, score, mi = trainer.validate(batch_size=b, num_workers=0, max_steps=1)
                tag = "v%d" % idx
                out[tag + ".pred"] = pred.numpy()
                out[tag + ".tgt"] = tgt.numpy()
                out[tag + ".losses"] = losses.detach().numpy()
                out[tag + ".acc"] = acc.detach().numpy()
                out[tag + ".score"] = np.array(float(score))
                out[tag + ".mi"] = mi.detach().numpy()
                cases.append({"tag": tag, "b": b, "k": k, "e": e, "all_steps": all_steps, "kind": kind})
                idx += 1
    out["cases"] = np.array(json.dumps(cases))
    np.savez_compressed(os.path.join(OUT, "validate.npz"), **out)


def _run_e24(ref, audio_seed, perturb=0.0, batch=4, layer_noise=0.0, experiment='e24', steps=2, no_dropout=False):
    """One run of the reference on experiments['e24']: setup_model, seeded weights, train() for two Adam steps.
    ``perturb`` > 0 multiplies the scalogram by (1 + perturb * randn): the conditioning probe (see golden_e24)."""
    import importlib
    import cpc_oracle_model as OM
    cfg = ref_shim.load_reference_configs()
    sf = importlib.import_module("setup_functions")
    am, cet = ref["audio_model"], ref["contrastive_estimation_training"]

    def ar_block_forward(self, x):                       # audio_model.py:124-136 with `main_x = main_x + ...`
        original_x = x
        for m in self.main_modules:
            x = m(x)
        main_x = x
        if self.residual:
            x = original_x
            for m in self.residual_modules:
                x = m(x)
            main_x = main_x + x[:, :, -main_x.shape[2]:]
        self.output_activation_writer(main_x)
        return main_x

    saved_forward = am.ConvolutionalArBlock.forward
    am.ConvolutionalArBlock.forward = ar_block_forward
    try:
        e = cfg.experiments[experiment]
        tc = e['training_config']
        if no_dropout and 'dropout' in e['ar_model_config']:
            e['ar_model_config']['dropout'] = 0.                  # the attention AR model's dropout draws from torch's RNG
        torch.manual_seed(0)
        model, pre, _ = sf.setup_model(cqt_params=e['cqt_config'], encoder_params=e['encoder_config'],
                                       ar_params=e['ar_model_config'], trainer_args=tc, device=None)
        OM.reseed_parameters(model.named_parameters())
        run = {"model": model, "tc": tc, "names": [n for n, _ in model.named_parameters()],
               "before": {n: p.detach().clone() for n, p in model.named_parameters()}}
        items = OM.e24_audio(2 * batch, model.item_length, seed=audio_seed)
        captured = run["captured"] = {}

        class RecordingAdam(torch.optim.Adam):
            def step(self, closure=None):
                if "grads" not in captured:
                    captured["grads"] = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
                return super().step(closure)

        order = run["order"] = []

        class LoggingDataset(ListDataset):
            def __getitem__(self, i):
                order.append(int(i))
                return self.items[i]

        def keep(key):
            def hook(module, inputs, output):                    # must return None: a returned value replaces the output
                if key not in captured:
                    captured[key] = output.detach().clone()
            return hook

        hooks = [model.encoder.register_forward_hook(keep("z")), pre.register_forward_hook(keep("scal"))]
        if layer_noise > 0:
            # every conv / linear OUTPUT multiplied by (1 + layer_noise * randn): rounding differences of the size an
            # independent fp32-faithful kernel has in each layer (forward only; hooks sit outside the reference code)
            layer_gen = torch.Generator().manual_seed(7)

            def jitter(module, inputs, output):
                return output * (1 + layer_noise * torch.randn(output.shape, generator=layer_gen))

            for m in model.modules():
                if isinstance(m, (torch.nn.Conv1d, torch.nn.Conv2d, torch.nn.Linear)):
                    hooks.append(m.register_forward_hook(jitter))
        preprocessing = pre
        if perturb > 0:
            noise_gen = torch.Generator().manual_seed(99)

            class Perturbed(torch.nn.Module):
                def forward(self, x):
                    y = pre(x)
                    return y * (1 + perturb * torch.randn(y.shape, generator=noise_gen))

            preprocessing = Perturbed()
        logger = run["logger"] = CaptureLogger()
        trainer = cet.ContrastiveEstimationTrainer(model=model, dataset=LoggingDataset(items), logger=logger, device=None,
                                                   optimizer=RecordingAdam, regularization=tc['regularization'],
                                                   prediction_noise=tc['prediction_noise'], file_batch_size=1,
                                                   score_over_all_timesteps=tc['score_over_all_timesteps'],
                                                   score_function=tc['score_function'],
                                                   wasserstein_gradient_penalty=tc['wasserstein_gradient_penalty'],
                                                   gradient_penalty_factor=tc['gradient_penalty_factor'],
                                                   preprocessing=preprocessing, prediction_steps=tc['prediction_steps'],
                                                   ar_size=model.ar_size)
        random.seed(0)
        trainer.train(batch_size=batch, epochs=1, lr=tc['learning_rate'], num_workers=0,
                      max_steps=1 if (perturb > 0 or layer_noise > 0) else steps)
        for h in hooks:
            h.remove()
        run["items"] = items
    finally:
        am.ConvolutionalArBlock.forward = saved_forward
    return run


def golden_e29(ref):
    """The reference's DEFAULT experiment (train_script.py:11): high-resolution CQT (44.1 kHz, 292 bins, hop 256), offset +
    time-pooled scalogram, resnet arch 9, attention AR model, linear all-steps scoring with the Wasserstein gradient penalty.
    One training step of setup_model(experiments['e29']) + train() at the full item length (367 616 samples), batch 2, with
    the attention dropout set to 0 (it draws from torch's RNG) and the seeded weights / audio of golden_e24: the logged loss
    (InfoNCE + penalty), max score and the encoder output."""
    run = _run_e24(ref, 1234, batch=2, experiment='e29', steps=1, no_dropout=True)
    model, tc, logger, captured = run["model"], run["tc"], run["logger"], run["captured"]
    out = {"item_length": np.array(model.item_length), "batch": np.array(2), "order": np.array(run["order"]),
           "losses": np.array(logger.losses), "max_scores": np.array(logger.scores), "z": captured["z"].numpy(),
           "audio_check": run["items"][:, ::4099].numpy().copy(), "scal_shape": np.array(captured["scal"].shape),
           "gradient_penalty_factor": np.array(float(tc["gradient_penalty_factor"])),
           "names": np.array(json.dumps(run["names"]))}
    # conditioning of the penalised loss: the penalty is a function of d(scores)/d(scalogram), i.e. of the ReLU gate pattern.
    # The reference's own loss under the probes of golden_e24 (every conv / linear output * (1 + eps randn)):
    for tag, eps in (("loss_snl", 2e-6), ("loss_snl5", 5e-6)):
        probe = _run_e24(ref, 1234, batch=2, experiment='e29', steps=1, no_dropout=True, layer_noise=eps)
        out[tag] = np.array(abs(probe["logger"].losses[0] - logger.losses[0]) / abs(logger.losses[0]))
        out[tag + "_z"] = np.array(float((probe["captured"]["z"] - captured["z"]).norm() / captured["z"].norm()))
        print(tag, float(out[tag]), "encoder output moves by", float(out[tag + "_z"]))
    np.savez_compressed(os.path.join(OUT, "e29_step.npz"), **out)
    print("e29 golden: loss", logger.losses, "max score", logger.scores, "z", tuple(captured["z"].shape))


def golden_e24(ref):
    """BASELINE configs[1] itself: the reference's setup_model(experiments['e24']) (CQT+phase -> resnet arch 7 ->
    ConvolutionalArModel arch 3 -> Linear 256 -> 16*512) at the full item length L = 97 024, batch 4, trained by the
    reference's own ContrastiveEstimationTrainer.train() for two Adam steps (lr 1e-4, the experiment's settings:
    linear score, all-steps softmax, regularization 0).  Stored: both logged losses / max scores, the encoder output
    and scalogram statistics of step 1, every parameter gradient of step 1 (seeded 8192-entry subsamples + full norms)
    and every parameter after step 2 (same subsamples).  Weights and audio are NOT stored: they are regenerated from
    seeds by oracle/cpc_oracle_model.py (reseed_parameters / e24_audio), which this script uses too.

    One line of the reference cannot run on torch >= 1.5: audio_model.py:133 adds the residual IN PLACE onto a ReLU
    output (autograd error in backward).  The forward of that block is replaced here by the same statements with the
    add written out of place -- numerically identical.

    Conditioning (measured here, stored as ``sn6.<name>`` / ``sn7.<name>``).  The gradients of a ReLU / max-pool network
    are discontinuous in the activations: a pre-activation within rounding distance of zero gates its whole downstream
    gradient on or off.  The log-power scalogram sits around -14 with ulp 9.5e-7 while its informative spread is ~1.3, so
    ANY independent evaluation of CQT + log (other summation order, other log) differs from the reference by ~1 ulp on
    many entries -- and the reference's OWN gradients move by 2e-3 ... 1e-2 (encoder, first AR layers) when its scalogram is
    multiplied by (1 + 1e-7 * randn) resp. (1 + 1e-6 * randn), or when only its CPU thread count changes (a gate at
    3.8e-7 in AR layer 1 flips between 1 and 8 threads: 8e-3 on every encoder gradient).  Sixteen audio seeds all behave
    the same.  A 1e-3 bound on these gradients is therefore not a property any implementation can have; the tests hold
    forward quantities (scalogram, encoder output, loss, max score) to 1e-3 and each gradient to
    max(1e-3, 3 x the reference's self-noise), the self-noise being the largest of three probes of the reference itself:
    scalogram * (1 + 1e-6 randn) (``sn6``), scalogram * (1 + 1e-7 randn) (``sn7``), and every conv / linear output
    * (1 + 2e-6 randn) resp. (1 + 5e-6 randn) (``snl``, ``snl5``).  2e-6 per layer is the forward error of the fp32-faithful
    tensor-core kernels (bf16 hi/lo operands): it moves the reference's encoder output by 2.5e-5, the CUDA path's encoder
    output differs from the oracle's by 3.0e-5 (tools/diag_e24_gates.py)."""
    import cpc_oracle_model as OM
    audio_seed = 1234
    run = _run_e24(ref, audio_seed)
    keys = set(run["model"].state_dict().keys())
    shadowed = set()
    for k in keys:                                                # conv bias directly in front of a batch norm: true gradient 0
        if k.endswith(".bias"):
            head, idx = k[:-len(".bias")].rsplit(".", 1)
            if idx.isdigit() and ("%s.%d.running_mean" % (head, int(idx) + 1)) in keys:
                shadowed.add(k)
    noises = {}
    for tag, eps in (("sn6", 1e-6), ("sn7", 1e-7), ("snl", 2e-6), ("snl5", 5e-6)):
        probe = _run_e24(ref, audio_seed, layer_noise=eps) if tag.startswith("snl") else _run_e24(ref, audio_seed, perturb=eps)
        if tag.startswith("snl"):
            zd = (run["captured"]["z"] - probe["captured"]["z"]).norm() / run["captured"]["z"].norm()
            print(tag, "encoder output moves by %.2e" % float(zd))
        noises[tag] = {}
        for n in run["names"]:
            a, b = run["captured"]["grads"][n].double(), probe["captured"]["grads"][n].double()
            noises[tag][n] = float((a - b).norm() / a.norm().clamp_min(1e-30))
        print(tag, "worst self-noise %.2e" % max(v for n, v in noises[tag].items() if n not in shadowed))
    model, tc, names, captured, logger = run["model"], run["tc"], run["names"], run["captured"], run["logger"]
    items, order, before = run["items"], run["order"], run["before"]
    batch = 4
    scal = captured["scal"].detach()
    out = {"item_length": np.array(model.item_length), "batch": np.array(batch), "lr": np.array(tc['learning_rate']),
           "audio_seed": np.array(audio_seed),
           "order": np.array(order), "losses": np.array(logger.losses), "max_scores": np.array(logger.scores),
           "audio_check": items[:, ::4099].numpy().copy(),
           "z": captured["z"].numpy(), "scal_shape": np.array(scal.shape),
           "scal_power_sub": scal[:, 0, ::7, ::11].numpy().copy(), "scal_phase_sub": scal[:, 1, ::7, ::11].numpy().copy(),
           "scal_power_norm": np.array(float(scal[:, 0].double().norm())),
           "names": np.array(json.dumps(names)), "bn_shadowed": np.array(json.dumps(sorted(shadowed))),
           "score_kind": np.array(tc['score_function'].__name__), "all_steps": np.array(bool(tc['score_over_all_timesteps'])),
           "regularization": np.array(float(tc['regularization'])),
           "prediction_steps": np.array(int(tc['prediction_steps'])), "visible_steps": np.array(int(tc['visible_steps']))}
    after = dict(model.named_parameters())
    for n in names:
        idx = OM.subsample_index(n, before[n].numel())
        g = captured["grads"][n]
        out["g." + n] = g.reshape(-1)[idx].numpy().copy()
        out["gn." + n] = np.array(float(g.double().norm()))
        out["sn6." + n] = np.array(noises["sn6"][n])
        out["sn7." + n] = np.array(noises["sn7"][n])
        out["snl." + n] = np.array(noises["snl"][n])
        out["snl5." + n] = np.array(noises["snl5"][n])
        out["p2." + n] = after[n].detach().reshape(-1)[idx].numpy().copy()
    for k, v in model.state_dict().items():                       # batch-norm running statistics after two steps
        if k.endswith("running_mean") or k.endswith("running_var"):
            out["bn." + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "e24_step.npz"), **out)
    print("e24 golden: audio seed", audio_seed, "losses", logger.losses, "max scores", logger.scores, "order", order)


AUDIO_DATASET_FILES = [("b_second.wav", 5000), ("a_first.wav", 1234), ("sub/c_third.wav", 8000), ("d_last.wav", 300)]


def golden_audio_dataset(ref):
    """The reference's AudioDataset / AudioTestingDataset index arithmetic and cross-file item assembly (audio_dataset.py:
    75-165) with only ``load_file`` replaced (it calls a torchaudio 0.2 API that no longer exists): file k holds the samples
    k*100000 + position, so an item lists exactly which (file, position) pairs the reference reads."""
    import tempfile
    ad = ref["audio_dataset"]
    lengths = {}

    def fake_loader(base):
        class Stub(base):
            def load_file(self, file, frames=-1, start=0):
                name = os.path.relpath(str(file), str(self.location)).replace(os.sep, "/")
                k = [n for n, _ in AUDIO_DATASET_FILES].index(name)
                n = lengths[name]
                stop = n if frames == -1 else min(n, int(start) + int(frames))
                return (k * 100000 + torch.arange(int(start), stop)).type(self.dtype)
        return Stub

    out = {"files": AUDIO_DATASET_FILES, "cases": []}
    with tempfile.TemporaryDirectory() as tmp:
        for name, n in AUDIO_DATASET_FILES:
            os.makedirs(os.path.dirname(os.path.join(tmp, name)), exist_ok=True)
            open(os.path.join(tmp, name), "wb").close()
            lengths[name] = n
        for cls_name, item, unique in (("AudioDataset", 1000, 400), ("AudioDataset", 700, 700), ("AudioDataset", 1500, 64),
                                       ("AudioTestingDataset", 250, 100)):
            ds = fake_loader(getattr(ad, cls_name))(tmp, item_length=item, unique_length=unique)
            case = {"class": cls_name, "item_length": item, "unique_length": unique, "len": len(ds),
                    "order": [os.path.relpath(str(f), tmp).replace(os.sep, "/") for f in ds.files],
                    "start_samples": [int(v) for v in ds.start_samples],
                    "counts": [int(v) for v in ds.get_example_count_per_file()], "items": {}}
            for idx in sorted(set([0, 1, 2, len(ds) // 3, len(ds) // 2, len(ds) - 2, len(ds) - 1] + list(range(10, 16)))):
                if 0 <= idx < len(ds):
                    item_value = ds[idx]
                    label = None
                    if cls_name == "AudioTestingDataset":
                        item_value, label = item_value[0], int(item_value[1])
                    v = item_value.long()
                    case["items"][str(idx)] = {"first": int(v[0]), "last": int(v[-1]), "n": int(v.numel()),
                                               "sum": int(v.sum()), "label": label,
                                               "breaks": [int(i) for i in (v[1:] - v[:-1] != 1).nonzero().flatten()]}
            out["cases"].append(case)
    with open(os.path.join(OUT, "audio_dataset.json"), "w") as fh:
        json.dump(out, fh)
    print("audio dataset golden:", [(c["class"], c["len"], c["counts"]) for c in out["cases"]])


def golden_sampler(ref):
    ad = ref["audio_dataset"]
    out, cases = {}, []
    settings = [([64], 8, 1, None, 0), ([37, 12, 50], 16, 1, None, 5), ([40, 24, 33], 16, 8, None, 1),
                ([40, 24, 33], 16, 8, 0, None), ([100], 32, 8, 0, None), ([17, 9], 4, 4, 3, None),
                ([256], 64, 8, None, 0)]
    for i, (counts, bs, fbs, seed, global_seed) in enumerate(settings):
        if global_seed is not None:
            random.seed(global_seed)
        s = ad.FileBatchSampler(counts, bs, fbs, drop_last=True, seed=seed)
        epochs = [[list(map(int, b)) for b in iter(s)] for _ in range(2)]
        cases.append({"counts": counts, "batch_size": bs, "file_batch_size": fbs, "seed": seed,
                      "global_seed": global_seed, "epochs": epochs, "len": int(len(s))})
    with open(os.path.join(OUT, "sampler.json"), "w") as fh:
        json.dump(cases, fh)


def golden_configs():
    cfg = ref_shim.load_reference_configs()

    def clean(v):
        if isinstance(v, dict):
            return {k: clean(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [clean(x) for x in v]
        if isinstance(v, (int, float, str, bool)) or v is None:
            return v
        return getattr(v, "__name__", str(v))

    dump = {}
    for name in ("e24", "e25", "e20", "e29", "e32"):          # e29: the reference's default experiment (train_script.py:11)
        e = cfg.experiments[name]
        dump[name] = {k: clean(e[k]) for k in ("cqt_config", "encoder_config", "ar_model_config", "training_config")}
    with open(os.path.join(OUT, "configs.json"), "w") as fh:
        json.dump(dump, fh, indent=1, sort_keys=True)


def build_all_experiments(experiments, setup_model):
    """Build every experiment of configs/experiment_configs.py in dict order through ``setup_model`` (the configs alias
    and mutate shared dicts, so order matters) -> {name: {'item_length', 'params': {key: shape}, 'buffers': [...]} |
    {'error': exception type}}.  Used for the reference here and for the compat/ shims in tests/test_compat_configs.py."""
    out = {}
    for name, e in experiments.items():
        if not isinstance(e, dict) or 'encoder_config' not in e:
            continue                                              # classification experiments (c1, c2): other workload
        try:
            tc = e['training_config']
            model, pre, _ = setup_model(cqt_params=e['cqt_config'], encoder_params=e['encoder_config'],
                                        ar_params=e['ar_model_config'], trainer_args=tc, device=None)
            sd = model.state_dict()
            params = {n for n, _ in model.named_parameters()}
            out[name] = {"item_length": int(model.item_length),
                         "params": {k: list(v.shape) for k, v in sd.items() if k in params},
                         "buffers": sorted(k for k in sd if k not in params),
                         "preprocessing": {k: list(v.shape) for k, v in pre.state_dict().items()},
                         "score_function": tc['score_function'].__name__,
                         "encoder": type(model.encoder).__name__, "ar": type(model.autoregressive_model).__name__}
        except Exception as exc:                                   # noqa: BLE001 -- the reference's own failures are data here
            out[name] = {"error": type(exc).__name__}
    return out


def golden_configs_all(ref=None):
    """Every experiment the reference's configs define (e0 ... e32, default, *_local), built by the reference's own
    setup_model: item length, parameter names / shapes, buffer names.  tests/test_compat_configs.py builds the same
    experiments from the UNCHANGED reference configs against the product through compat/ and compares."""
    import importlib
    cfg = ref_shim.load_reference_configs()
    sf = importlib.import_module("setup_functions")
    out = build_all_experiments(cfg.experiments, sf.setup_model)
    with open(os.path.join(OUT, "configs_all.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    print({k: (v.get("item_length") or v.get("error")) for k, v in out.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    ref = ref_shim.load_reference()
    golden_configs()
    golden_sampler(ref)
    golden_audio_dataset(ref)
    golden_cqt(ref)
    golden_cqt_high_res(ref)
    golden_audio_encoder(ref)
    golden_resnet_encoder(ref)
    golden_infonce(ref)
    golden_trainer_raw(ref)
    golden_trainer_cqt(ref)
    golden_validate(ref)
    golden_trainer_gp(ref)
    golden_cqt_grad(ref)
    golden_snapshot(ref)
    golden_scalogram_encoder(ref)
    golden_e24(ref)
    golden_e29(ref)
    golden_configs_all(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    if len(sys.argv) > 1:                      # regenerate selected goldens only: make_golden.py trainer_cqt ...
        torch.set_num_threads(8)
        _ref = ref_shim.load_reference()
        for _name in sys.argv[1:]:
            globals()["golden_" + _name](_ref)
    else:
        main()
