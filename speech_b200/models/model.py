"""Encoder base class - host-side mirror of speech/models/model.py:10-113 (reference).

Same constructor, attributes, state_dict key names (`conv.{0,2,..}.{weight,bias}`,
`rnn.weight_ih_l{k}[_reverse]` ...) and picklability as the reference `Model`; the nn.Conv2d /
nn.GRU modules are kept as PARAMETER CONTAINERS (identical construction order => identical
initialisation under the same torch seed) while the arithmetic of `encode` runs in the
hand-written sm_100a kernels (speech_b200/ops.py -> csrc/).  There is no CPU path: calling
`encode` on a CPU tensor raises.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, ops


class Model(nn.Module):

    def __init__(self, input_dim, config):
        super().__init__()
        self.input_dim = input_dim
        enc = config["encoder"]
        layers = []
        in_c = 1
        out_c = 1
        for out_c, h, w, s in enc["conv"]:
            layers.append(nn.Conv2d(in_c, out_c, (h, w), stride=(s, s), padding=0))
            layers.append(nn.ReLU())
            if config["dropout"] != 0:
                layers.append(nn.Dropout(p=config["dropout"]))
            in_c = out_c
        self.conv = nn.Sequential(*layers)
        conv_out = out_c * self.conv_out_size(input_dim, 1)
        assert conv_out > 0, "Convolutional output frequency dimension is negative."

        rnn_cfg = enc["rnn"]
        self.rnn = nn.GRU(input_size=conv_out, hidden_size=rnn_cfg["dim"],
                          num_layers=rnn_cfg["layers"], batch_first=True,
                          dropout=config["dropout"], bidirectional=rnn_cfg["bidirectional"])
        self._encoder_dim = rnn_cfg["dim"]
        self.volatile = False

    # reference: model.py:44-52 (valid convolution, ceil((n - k + 1) / s) per layer)
    def conv_out_size(self, n, dim):
        for c in self.conv.children():
            if isinstance(c, nn.Conv2d):
                k, s = c.kernel_size[dim], c.stride[dim]
                n = int(math.ceil((n - k + 1) / s))
        return n

    def forward(self, batch):
        raise NotImplementedError

    def loss(self, batch):
        raise NotImplementedError

    def infer(self, batch):
        raise NotImplementedError

    def encode(self, x):
        """x (B, T, F) float32 on the CUDA device -> (B, T', encoder_dim)   (model.py:60-79)."""
        _lib.require_cuda(x, "encode() input")
        x = ops.conv_stack(x, self.conv, self.training)          # (B, T', C*F') c-major features
        p = self.rnn.dropout if self.training else 0.0
        x = ops.gru_stack(x, self.rnn, dropout=p)                 # (B, T', ndir*H)
        if self.rnn.bidirectional:
            half = x.shape[-1] // 2
            x = x[:, :, :half] + x[:, :, half:]
        return x

    def set_eval(self):
        self.eval()
        self.volatile = True

    def set_train(self):
        self.train()
        self.volatile = False

    @property
    def is_cuda(self):
        return next(self.parameters()).is_cuda

    @property
    def encoder_dim(self):
        return self._encoder_dim

    def _grad_ctx(self):
        # the reference marks eval batches `volatile`; on modern torch that is no_grad
        return torch.no_grad() if self.volatile else torch.enable_grad()


class LinearND(nn.Module):
    """nn.Linear over the last dimension of an N-D input (reference model.py:115-133)."""

    def __init__(self, *args):
        super().__init__()
        self.fc = nn.Linear(*args)

    def forward(self, x):
        lead = x.shape[:-1]
        out = self.fc(x.reshape(-1, x.shape[-1]))
        return out.view(*lead, out.shape[-1])


def zero_pad_concat(inputs):
    """list of (T_i, F) arrays -> (B, max T, F) float32, zero padded (reference model.py:135-141)."""
    max_t = max(inp.shape[0] for inp in inputs)
    out = np.zeros((len(inputs), max_t, inputs[0].shape[1]), dtype=np.float32)
    for e, inp in enumerate(inputs):
        out[e, :inp.shape[0], :] = inp
    return out


_pinned = {}


def zero_pad_concat_pinned(inputs):
    """zero_pad_concat into a reused PINNED staging tensor so the H2D copy can be asynchronous
    (batch assembly becomes the critical path once the step is ms-scale)."""
    max_t = max(inp.shape[0] for inp in inputs)
    shape = (len(inputs), max_t, inputs[0].shape[1])
    buf = _pinned.get(shape)
    if buf is None:
        buf = torch.zeros(shape, dtype=torch.float32).pin_memory()
        _pinned.clear()
        _pinned[shape] = buf
    arr = buf.numpy()
    for e, inp in enumerate(inputs):
        n = inp.shape[0]
        arr[e, :n, :] = inp
        if n < max_t:
            arr[e, n:, :] = 0.0
    return buf
