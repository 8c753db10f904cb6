"""RNN-Transducer model - host-side mirror of speech/models/transducer_model.py:14-116.

Encoder and prediction network run on the sm_100a GRU kernels (ops.gru_stack), the loss on
sb_rnnt_fwd_bwd.  The joint (fc1 shared by both streams, transducer_model.py:71-73) and the
log-softmax are torch ops in round 1; fusing joint -> log-softmax -> lattice so that the
(B,T',U+1,H) intermediate is never materialised is the next kernel on this row.
"""
import torch
import torch.nn as nn

from . import model
from .. import ops
from ..functions import transducer as transducer_fn
from ..transducer_decoders import decode_static


class Transducer(model.Model):

    def __init__(self, freq_dim, vocab_size, config):
        super().__init__(freq_dim, config)
        dec = config["decoder"]
        rnn_dim = self.encoder_dim
        self.embedding = nn.Embedding(vocab_size, dec["embedding_dim"])
        self.dec_rnn = nn.GRU(input_size=dec["embedding_dim"], hidden_size=rnn_dim,
                              num_layers=dec["layers"], batch_first=True,
                              dropout=config["dropout"])
        self.blank = vocab_size                       # blank is the LAST class (:28)
        self.fc1 = model.LinearND(rnn_dim, rnn_dim)
        self.fc2 = model.LinearND(rnn_dim, vocab_size + 1)

    def forward(self, batch):
        x, y, x_lens, y_lens = self.collate(*batch)
        y_mat = self.label_collate(batch[1])
        with self._grad_ctx():
            return self.forward_impl(x, y_mat)

    def forward_impl(self, x, y):
        if self.is_cuda:
            x = x.cuda(non_blocking=True)
            y = y.cuda(non_blocking=True)
        return self.decode(self.encode(x), y)

    def loss(self, batch):
        x, y, x_lens, y_lens = self.collate(*batch)
        y_mat = self.label_collate(batch[1])
        with self._grad_ctx():
            out = self.forward_impl(x, y_mat)
            return transducer_fn.TransducerLoss()(out, y, x_lens, y_lens)

    def decode(self, x, y):
        """x (B, T', H) encoder states, y (B, U) labels -> (B, T', U+1, V+1) log-probs (:54-77)."""
        emb = self.embedding(y)
        start = torch.zeros((emb.shape[0], 1, emb.shape[2]), device=emb.device, dtype=emb.dtype)
        pred_in = torch.cat([start, emb], dim=1)           # zero vector stands for "no label yet"
        p = self.dec_rnn.dropout if self.training else 0.0
        pred = ops.gru_stack(pred_in, self.dec_rnn, dropout=p)
        joint = self.fc1(x).unsqueeze(2) + self.fc1(pred).unsqueeze(1)   # the SAME fc1 for both
        out = self.fc2(torch.relu(joint))
        return torch.log_softmax(out, dim=3)

    def collate(self, inputs, labels):
        return self.lattice_batch(inputs, labels)

    def infer(self, batch, beam_size=4):
        """Beam search on the TEACHER-FORCED lattice, as the reference does (:92-101), including
        its use of the un-subsampled input length as the time bound (clamped by slicing)."""
        with torch.no_grad():
            out = self(batch)
        preds = []
        for e, (i, l) in enumerate(zip(*batch)):
            T = min(i.shape[0], out.shape[1])
            U = len(l) + 1
            preds.append(decode_static(out[e, :T, :U, :], beam_size, blank=self.blank)[0])
        return preds

    def label_collate(self, labels):
        return torch.from_numpy(model.end_pad_labels(labels))
