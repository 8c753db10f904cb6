"""CTC model - host-side mirror of speech/models/ctc_model.py:13-70 (reference)."""
import numpy as np
import torch

from . import model
from .ctc_decoder import decode, decode_batch
from ..functions import ctc


class CTC(model.Model):

    def __init__(self, freq_dim, output_dim, config):
        super().__init__(freq_dim, config)
        self.blank = output_dim                      # blank is the LAST class (ctc_model.py:18)
        self.fc = model.LinearND(self.encoder_dim, output_dim + 1)

    def _collated(self, batch):
        """reference-style (inputs, labels) pair, or a loader.StagedBatch prepared ahead of time"""
        from ..loader import StagedBatch
        if isinstance(batch, StagedBatch):
            return batch.tensors()
        return self.collate(*batch)

    def forward(self, batch):
        x, y, x_lens, y_lens = self._collated(batch)
        with self._grad_ctx():
            return self.forward_impl(x)

    def forward_impl(self, x, softmax=False):
        if self.is_cuda:
            x = x.cuda(non_blocking=True)
        x = self.encode_logits(x)
        if softmax:
            return torch.nn.functional.softmax(x, dim=2)
        return x

    def encode_logits(self, x):
        """encode + fc with the halves-sum and the projection fused into one contraction on the
        bf16 top-layer output (same math as model.py:75-77 followed by ctc_model.py:29)."""
        from .. import _lib, ops
        _lib.require_cuda(x, "forward_impl() input")
        if getattr(self, "parity_mode", False) and not torch.is_grad_enabled():
            # reference-precision forward (split-bf16 GEMMs, fp32 recurrence): a measuring stick
            # for the bf16 operand path, see ops.encode_logits_parity
            return ops.encode_logits_parity(x, self.conv, self.rnn, self.fc.fc)
        x = ops.conv_stack(x, self.conv, self.training)
        p = self.rnn.dropout if self.training else 0.0
        return ops.gru_stack_logits(x, self.rnn, self.fc.fc, dropout=p)

    def loss(self, batch):
        x, y, x_lens, y_lens = self._collated(batch)
        with self._grad_ctx():
            out = self.forward_impl(x)
            return self.ctc_loss(out, y, x_lens, y_lens)

    def ctc_loss(self, out, y, x_lens, y_lens):
        """warp-ctc call of the reference (ctc_model.py:38-39): raw logits, blank = last class."""
        return ctc.CTCLoss()(out, y, x_lens, y_lens)

    def collate(self, inputs, labels):
        return self.lattice_batch(inputs, labels)

    def infer(self, batch, beam_size=1):
        x, y, x_lens, y_lens = self._collated(batch)
        with torch.no_grad():
            probs = self.forward_impl(x, softmax=True)
        return decode_batch(probs, beam_size=beam_size, blank=self.blank)

    @staticmethod
    def max_decode(pred, blank):
        """Greedy collapse: merge repeats, then drop blanks (reference ctc_model.py:62-70)."""
        seq = []
        prev = None
        for p in pred:
            if p != blank and p != prev:
                seq.append(p)
            prev = p
        return seq
