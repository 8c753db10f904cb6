"""CPU oracle for the encoder / CTC model (TEST INFRASTRUCTURE - never imported by speech_b200/).

Restates, with stock torch.nn modules on the CPU, what the reference computes:
  encode            speech/models/model.py:60-79  (Conv2d+ReLU stack -> (B,T',C*F') channel-major
                    flatten -> nn.GRU -> bidirectional halves summed)
  CTC forward/loss  speech/models/ctc_model.py:25-40 (encoder -> Linear -> warp-ctc on raw logits,
                    blank = last class, sum over the minibatch)
The CTC arithmetic itself is the un-vendored awni/warp-ctc (Makefile:4-7); here it is
log_softmax + torch.nn.functional.ctc_loss(reduction='sum'), which tests/test_oracle.py pins
against the independent float64 alpha/beta of oracle/ctc_ref.py.
This module is also the `cpu_baseline` / `--impl reference` arm of bench.py: the reference is pure
Python over torch CPU ops, so this is the same ATen/MKL work the reference would run on the host.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class RefEncoder(nn.Module):
    def __init__(self, input_dim, config):
        super().__init__()
        convs = []
        in_c = 1
        f = input_dim
        for out_c, h, w, s in config["encoder"]["conv"]:
            convs += [nn.Conv2d(in_c, out_c, (h, w), stride=(s, s)), nn.ReLU()]
            if config["dropout"] != 0:
                convs.append(nn.Dropout(config["dropout"]))
            in_c = out_c
            f = int(math.ceil((f - w + 1) / s))
        self.conv = nn.Sequential(*convs)
        r = config["encoder"]["rnn"]
        self.rnn = nn.GRU(in_c * f, r["dim"], r["layers"], batch_first=True,
                          dropout=config["dropout"], bidirectional=r["bidirectional"])
        self.dim = r["dim"]

    def time_out(self, n):
        for c in self.conv:
            if isinstance(c, nn.Conv2d):
                n = int(math.ceil((n - c.kernel_size[0] + 1) / c.stride[0]))
        return n

    def forward(self, x):
        y = self.conv(x.unsqueeze(1))                       # (B, C, T', F')
        b, c, t, f = y.shape
        y = y.transpose(1, 2).reshape(b, t, c * f)          # channel-major features
        y, _ = self.rnn(y)
        if self.rnn.bidirectional:
            y = y[..., :self.dim] + y[..., self.dim:]
        return y


class RefCTC(nn.Module):
    def __init__(self, input_dim, vocab, config):
        super().__init__()
        self.enc = RefEncoder(input_dim, config)
        self.fc = nn.Linear(self.enc.dim, vocab + 1)
        self.blank = vocab

    def logits(self, x):
        return self.fc(self.enc(x))

    def loss(self, x, labels_flat, label_lens):
        out = self.logits(x)                                 # (B, T', V+1)
        B, T = out.shape[:2]
        lp = F.log_softmax(out, 2).transpose(0, 1)
        return F.ctc_loss(lp, labels_flat.long(), torch.full((B,), T, dtype=torch.long),
                          label_lens.long(), blank=self.blank, reduction="sum")

    def load_from_dropin(self, sd):
        """Copy a speech_b200.models.CTC / reference CTC state_dict (conv.*, rnn.*, fc.fc.*)."""
        mine = {}
        for k, v in sd.items():
            if k.startswith("conv.") or k.startswith("rnn."):
                mine["enc." + k] = v
            elif k.startswith("fc.fc."):
                mine["fc." + k[len("fc.fc."):]] = v
        self.load_state_dict(mine)
