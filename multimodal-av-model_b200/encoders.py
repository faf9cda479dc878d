"""Producers that feed the hot path: VisualEncoder and AudioEncoder.

These are OUT OF SCOPE for hand-written kernels (frozen 3D-conv + ResNet-18 front-end; third-party wav2vec2)
and stay PyTorch/cuDNN/HF, exactly as SURVEY.md §2 scopes them.  They exist here only so that a full training
step (BASELINE config 4) can run end to end on the GPU box, where /root/reference is absent.  Module and
parameter names follow /root/reference/model/encoder.py:6-100 so reference checkpoints load unchanged
(`frontend3D.*`, `trunk.layer{1..4}.*`, `model.*`).
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _act(kind, channels):
    return nn.PReLU(channels) if kind == "prelu" else nn.ReLU(inplace=True)


class BasicBlock(nn.Module):
    """3x3-3x3 residual block with a per-channel PReLU (encoder.py:6-22)."""

    def __init__(self, inplanes, planes, stride=1, downsample=None, relu_type="prelu"):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = _act(relu_type, planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        skip = x if self.downsample is None else self.downsample(x)
        y = self.bn2(self.conv2(self.relu(self.bn1(self.conv1(x)))))
        return self.relu(y + skip)


class ResNet(nn.Module):
    """ResNet trunk without stem: four stages of `layers[i]` blocks, widths 64/128/256/512 (encoder.py:24-53)."""

    def __init__(self, block, layers, relu_type="prelu"):
        super().__init__()
        self.inplanes = 64
        widths, strides = (64, 128, 256, 512), (1, 2, 2, 2)
        for i, (w, s, n) in enumerate(zip(widths, strides, layers), start=1):
            setattr(self, f"layer{i}", self._stage(block, w, n, s, relu_type))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))

    def _stage(self, block, planes, blocks, stride, relu_type):
        down = None
        if stride != 1 or self.inplanes != planes:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False),
                                 nn.BatchNorm2d(planes))
        mods = [block(self.inplanes, planes, stride, down, relu_type)]
        self.inplanes = planes
        mods += [block(planes, planes, relu_type=relu_type) for _ in range(blocks - 1)]
        return nn.Sequential(*mods)

    def forward(self, x):
        for i in range(1, 5):
            x = getattr(self, f"layer{i}")(x)
        return torch.flatten(self.avgpool(x), 1)


class VisualEncoder(nn.Module):
    """[B,1,T,96,96] -> [B,T,512]: Conv3d(1->64,(5,7,7),s(1,2,2)) + BN + PReLU + MaxPool3d, then a per-frame
    ResNet-18 trunk (encoder.py:57-75)."""

    def __init__(self, relu_type="prelu"):
        super().__init__()
        self.frontend3D = nn.Sequential(
            nn.Conv3d(1, 64, kernel_size=(5, 7, 7), stride=(1, 2, 2), padding=(2, 3, 3), bias=False),
            nn.BatchNorm3d(64),
            _act(relu_type, 64),
            nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1)))
        self.trunk = ResNet(BasicBlock, [2, 2, 2, 2], relu_type=relu_type)
        self.output_dim = 512

    def _channels_last(self):
        """cuDNN's NHWC tensor-core kernels and ATen's channels-last BatchNorm are ~2x faster here than NCHW on B200
        (measured: 20.4 -> 10.8 ms per [8,1,150,96,96] call); values are unchanged up to bf16 rounding."""
        if not getattr(self, "_cl_done", False):
            self.trunk.to(memory_format=torch.channels_last)
            self._cl_done = True

    def _frontend_as_2d(self, x):
        """frontend3D (encoder.py:57-62) evaluated frame-wise with 2-D kernels, same parameters and same arithmetic:
        a (5,7,7) Conv3d over ONE input channel with stride (1,2,2) is a 7x7 Conv2d whose 5 input channels are the
        temporal taps t-2..t+2 (zero padded); BatchNorm3d over (B,T,H,W) equals BatchNorm2d over (B*T,H,W);
        MaxPool3d((1,3,3)) is MaxPool2d(3) per frame.  cuDNN has bf16 tensor-core kernels for the 2-D form (the 3-D
        form falls back to a TF32 kernel plus layout conversions) and the [B,64,T,H,W] -> [B*T,64,H,W] transpose copy
        disappears.  Returns [B*T,64,H',W'] channels-last."""
        import torch.nn.functional as F
        conv, bn, act, pool = self.frontend3D[0], self.frontend3D[1], self.frontend3D[2], self.frontend3D[3]
        b, _, t, h, w = x.shape
        kt = conv.kernel_size[0]
        pt = conv.padding[0]
        xp = F.pad(x[:, 0], (0, 0, 0, 0, pt, pt))                                   # [B, T+2pt, H, W]
        taps = xp.unfold(1, kt, 1)                                                    # [B, T, H, W, kt] (view)
        taps = taps.permute(0, 1, 4, 2, 3).reshape(b * t, kt, h, w)                   # temporal taps as channels
        taps = taps.contiguous(memory_format=torch.channels_last)
        w2 = conv.weight[:, 0]                                                        # [64, kt, 7, 7]
        y = F.conv2d(taps, w2, None, stride=conv.stride[1:], padding=conv.padding[1:])
        if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias,
                         bn.training or not bn.track_running_stats, bn.momentum if bn.momentum is not None else 0.1, bn.eps)
        y = act(y)
        return F.max_pool2d(y, pool.kernel_size[1:], pool.stride[1:], pool.padding[1:])

    def forward(self, x):
        b, t = x.shape[0], x.shape[2]
        conv = self.frontend3D[0]
        if (x.is_cuda and x.shape[1] == 1 and conv.stride[0] == 1 and conv.dilation == (1, 1, 1)
                and isinstance(self.frontend3D[1], nn.BatchNorm3d) and self.frontend3D[1].momentum is not None):
            self._channels_last()
            y = self._frontend_as_2d(x)
            return self.trunk(y).view(b, t, 512)
        y = self.frontend3D(x)                                   # [B,64,T,H',W']
        t, h, w = y.shape[2:]
        y = y.transpose(1, 2).reshape(b * t, 64, h, w)
        return self.trunk(y).view(b, t, 512)


def xlsr_large_config(**overrides):
    """Wav2Vec2 XLSR-53-large layout (what kresnik/wav2vec2-large-xlsr-korean uses), for offline random init."""
    from transformers import Wav2Vec2Config
    cfg = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
               feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=True, num_conv_pos_embeddings=128,
               num_conv_pos_embedding_groups=16)
    cfg.update(overrides)
    return Wav2Vec2Config(**cfg)


class AudioEncoder(nn.Module):
    """HF Wav2Vec2Model wrapper (encoder.py:80-100): returns (last_hidden_state, mean of hidden_states[6:10]).
    `config=` builds a randomly initialised model without touching the network (bench / tests)."""

    def __init__(self, model_name="kresnik/wav2vec2-large-xlsr-korean", freeze=True, config=None):
        super().__init__()
        from transformers import Wav2Vec2Model
        if config is not None:
            config.output_hidden_states = True
            self.model = Wav2Vec2Model(config)
        else:
            self.model = Wav2Vec2Model.from_pretrained(model_name, output_hidden_states=True)
        self.output_dim = self.model.config.hidden_size
        if freeze:
            self.model.requires_grad_(False)

    def forward(self, x, attention_mask=None):
        # HF marks the conv feature extractor's output as requiring grad in train mode (a gradient-checkpointing
        # aid) unless freeze_feature_encoder() was called; the reference only sets requires_grad=False on the
        # parameters (main.py:26-31), so autograd back-propagates through seven frozen conv layers for nothing.
        # With every parameter frozen the flag changes no gradient that is ever used.
        fe = getattr(self.model, "feature_extractor", None)
        if fe is not None and getattr(fe, "_requires_grad", False) and not any(p.requires_grad for p in fe.parameters()):
            fe._requires_grad = False
        if fe is not None and not hasattr(fe, "_avctc_cached"):
            _install_feature_cache(fe)
        if attention_mask is not None:
            attention_mask = attention_mask.long()
        out = self.model(input_values=x, attention_mask=attention_mask, return_dict=True)
        middle = torch.stack(out.hidden_states[6:10], dim=0).mean(dim=0)
        return out.last_hidden_state, middle


def _install_feature_cache(fe):
    """The trainer runs the audio encoder twice per step on the SAME waveform tensor (once per speaker mask,
    trainer.py:94-95).  The conv feature extractor is deterministic (no dropout) and frozen, so its output for an
    unchanged input tensor is reused instead of recomputed: identical values, one conv stack pass per step instead of
    two.  Installed as an instance-level forward wrapper, so module structure and state_dict keys do not change."""
    import weakref
    inner = fe.forward
    state = {"ref": None, "key": None, "out": None}

    def cached_forward(input_values):
        frozen = not any(p.requires_grad for p in fe.parameters()) and not getattr(fe, "_requires_grad", False)
        if not frozen or torch.is_grad_enabled() and input_values.requires_grad:
            state["ref"] = None
            return inner(input_values)
        # identity of the live tensor OBJECT (a weak reference) + its version counter: a later batch that happens to be
        # allocated at the same address is a different object, and in-place edits bump the version
        key = (input_values._version, tuple(input_values.shape), input_values.dtype, torch.is_autocast_enabled(),
               fe.training)
        same = state["ref"] is not None and state["ref"]() is input_values and state["key"] == key
        if not same:
            with torch.no_grad():
                state["out"] = inner(input_values)
            state["ref"], state["key"] = weakref.ref(input_values), key
        return state["out"]

    fe.forward = cached_forward
    fe._avctc_cached = True


def unfreeze_middle_layers(model):
    """main.py:26-31: only encoder.layers.6..9 of the wav2vec2 model train."""
    tags = tuple(f"encoder.layers.{i}." for i in range(6, 10))
    for name, p in model.named_parameters():
        p.requires_grad = any(t in name for t in tags)
