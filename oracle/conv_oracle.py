"""CPU restatement of the reference's conv encoder and decoder.  TEST INFRASTRUCTURE ONLY.

Plain torch-CPU functional code over a state_dict with the reference's key names, NCHW like the reference, any
dtype (float64 for tight comparisons).  Each function cites the reference lines it follows.  Pinned to the
unmodified reference by tests/golden/conv_enc_*.npz (oracle/make_golden.py:gen_conv, checked in
tests/test_conv_oracle.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import torch
from torch.nn import functional as F

EPS = torch.finfo(torch.float32).eps


def normalize(weight, lognorm):
    """`_normalize` base/common.py:426-431 with normalize_dim = 0."""
    n = torch.exp(lognorm).view(-1, 1, 1, 1)
    wn = torch.linalg.vector_norm(weight, dim=[1, 2, 3], keepdim=True)
    return n * weight / (wn + EPS)


def conv2d(x, sd, prefix, stride=1, padding=0):
    """Conv2D.forward base/common.py:313-323."""
    return F.conv2d(x, normalize(sd[prefix + ".weight"], sd[prefix + ".lognorm"]), sd[prefix + ".bias"], stride=stride,
                    padding=padding)


def se_layer(x, sd, prefix):
    """SELayer.forward base/common.py:138-142."""
    se = x.mean(dim=[2, 3])
    se = F.relu(F.linear(se, sd[prefix + ".fc.0.weight"], sd[prefix + ".fc.0.bias"]))
    se = torch.sigmoid(F.linear(se, sd[prefix + ".fc.2.weight"], sd[prefix + ".fc.2.bias"]))
    return x * se[:, :, None, None]


def cell(x, sd, prefix, stride, eps=1.0):
    """Cell.forward base/common.py:187-192 with ConvLayer.forward :233-242 (SiLU -> conv), encoder cell types."""
    if stride == 1:
        skip = x
    elif (prefix + ".skip.ops.0.weight") in sd:  # FactorizedReduce.forward base/common.py:118-126
        a = F.silu(x)
        outs = []
        for idx in range(4):
            i, j = idx // 2, idx % 2
            outs.append(conv2d(a[..., i:, j:], sd, f"{prefix}.skip.ops.{idx}", stride=2))
        skip = torch.cat(outs, dim=1)
    else:  # StrideReduce.forward base/common.py:92-95
        skip = conv2d(F.silu(x), sd, prefix + ".skip.conv", stride=2)
    h = conv2d(F.silu(x), sd, prefix + ".ops.0.conv", stride=stride, padding=1)
    h = conv2d(F.silu(h), sd, prefix + ".ops.1.conv", stride=1, padding=1)
    h = se_layer(h, sd, prefix + ".se")
    return skip + eps * h


def res_conv_pool(x, sd, prefix, eps=1.0):
    """ResConvPool.forward main/vae.py:699-706."""
    skip = x.mean(dim=[2, 3], keepdim=True)
    return skip + eps * conv2d(F.silu(x), sd, prefix + ".conv")


def res_dense(x, sd, prefix):
    """ResDenseLayer.forward main/vae.py:672-679 in eval mode (dropout off)."""
    h = F.relu(F.linear(x, sd[prefix + ".fc1.weight"], sd[prefix + ".fc1.bias"]))
    h = F.linear(h, sd[prefix + ".fc2.weight"], sd[prefix + ".fc2.bias"])
    return F.layer_norm(h + x, (x.shape[1],), sd[prefix + ".layer_norm.weight"], sd[prefix + ".layer_norm.bias"], 1e-5)


def conv_encoder(x, sd, dataset: str):
    """BaseVAE.encode conv branch main/vae.py:40-52 with `_build_conv_enc` main/vae.py:710-756: x (N,1,H,W) ->
    pre-activation of the posterior (N, K)."""
    pad = 0 if dataset.endswith("MNIST") else 1
    n_layers = 3 if dataset.endswith("MNIST") else (2 if dataset in ("vH16", "CIFAR16") else 4)
    h = conv2d(x, sd, "stem", padding=pad)
    h = cell(h, sd, "enc.0", 1)
    i = 1
    for _ in range(n_layers):
        h = cell(h, sd, f"enc.{i}", 2)
        h = cell(h, sd, f"enc.{i + 1}", 1)
        i += 2
    h = res_conv_pool(h, sd, f"enc.{i}").flatten(1)
    h = res_dense(h, sd, f"enc.{i + 2}")
    return F.linear(h, sd["fc_enc.weight"], sd.get("fc_enc.bias"))


def dec_cell(x, sd, prefix, up: bool, eps=1.0):
    """Cell.forward base/common.py:187-192 for 'normal_dec' / 'up_dec' (n_nodes = 1): ConvLayer.forward :233-242
    (SiLU -> [nearest 2x up-sampling] -> conv), skip = identity or up-sampling -> 1x1 conv (base/common.py:35-47)."""
    h = F.silu(x)
    if up:
        skip = conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), sd, prefix + ".skip.1")
        h = F.interpolate(h, scale_factor=2, mode="nearest")
    else:
        skip = x
    h = conv2d(h, sd, prefix + ".ops.0.conv", padding=1)
    h = se_layer(h, sd, prefix + ".se")
    return skip + eps * h


def conv_decoder(z, sd, dataset: str):
    """BaseVAE.decode conv branch main/vae.py:54-58 with `_build_conv_dec` main/vae.py:760-789: z (N, K) -> (N, 1, S, S)."""
    nch = sd["fc_dec.weight"].shape[0] // 4
    h = F.linear(z, sd["fc_dec.weight"], sd.get("fc_dec.bias")).view(-1, nch, 2, 2)
    n_layers = 4 if dataset.endswith("MNIST") else 3
    h = dec_cell(h, sd, "dec.0", False)
    i = 1
    for layer in range(n_layers):
        h = dec_cell(h, sd, f"dec.{i}", True)
        h = dec_cell(h, sd, f"dec.{i + 1}", False)
        i += 2
        if dataset.endswith("MNIST") and layer == 2:
            h = F.interpolate(h, size=14, mode="nearest")  # nn.Upsample(size=14)
            i += 1
    return conv2d(h, sd, f"dec.{i}")
