"""Drop-in for `transducer.decoders` of awni/transducer (imported at
speech/models/transducer_model.py:10, used at :100 as `td.decode_static(lp, beam_size, blank)[0]`).

The beam search runs on the GPU (csrc/tdecode.cu, one CTA per utterance; hypotheses in a canonical
trie, float64 scores); `decode_static` keeps the reference's per-utterance call signature and
returns (labels, log_probability) so that `[0]` is the label list, `decode_static_batch` searches
a whole minibatch of lattices with one launch and one device->host copy (what
`Transducer.infer` uses).  The dependency is un-vendored (Makefile:10-12), so the algorithm is the
standard transducer beam search on a static lattice; its CPU restatement and how it is pinned are
in oracle/decode_static_ref.py (test infrastructure - not imported here).  There is no CPU path.
"""
import ctypes

import torch

from . import _lib


def decode_static_batch(lp, tlens, ulens, beam_size, blank):
    """lp (B, T, U1, V) float32 CUDA log-probabilities; tlens / ulens: frames and lattice rows
    (labels + 1) of every utterance.  Returns (list of label lists, list of log-probabilities)."""
    from . import ops
    _lib.require_cuda(lp, "lp")
    lib = _lib.load()
    lp = lp.detach().float().contiguous()
    B, T, U1, V = lp.shape
    dev = lp.device
    lens = torch.tensor([min(int(t), T) for t in tlens] + [min(int(u), U1) for u in ulens],
                        dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.sb_rnnt_decode_static_workspace_size(B, T, U1, int(beam_size),
                                                        ctypes.byref(nbytes)), "ws")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    labels = torch.empty(B, U1, dtype=torch.int32, device=dev)
    olens = torch.empty(B, dtype=torch.int32, device=dev)
    scores = torch.empty(B, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        sp = _lib.stream_ptr()
        ops._launch("rnnt_decode_static", 0.0,
                    lambda: lib.sb_rnnt_decode_static(lp.data_ptr(), lens[:B].data_ptr(),
                                                      lens[B:].data_ptr(), B, T, U1, V,
                                                      int(beam_size), int(blank),
                                                      labels.data_ptr(), olens.data_ptr(),
                                                      scores.data_ptr(), ws.data_ptr(),
                                                      nbytes.value, sp))
    lab = labels.cpu()
    n = olens.cpu().tolist()
    sc = scores.cpu().tolist()
    return [lab[b, :n[b]].tolist() for b in range(B)], sc


def decode_static(lp, beam_size, blank=0):
    """One utterance: lp (T, U, V) log-probabilities (numpy array or tensor; copied to the GPU if
    it is not there already) -> (labels, log_probability)."""
    lp = torch.as_tensor(lp)
    if not lp.is_cuda:
        lp = lp.cuda()
    T, U, V = lp.shape
    labels, scores = decode_static_batch(lp.unsqueeze(0), [T], [U], beam_size, blank)
    return labels[0], scores[0]
