"""Training objectives of the reference, device side (first slice).

Mirrors audio_sheet_retrieval/models/objectives.py:30-69: `get_contrastive_cos_loss(weight, gamma,
symmetric)` returns `loss(lv1, lv2)`.  Where the reference builds a Theano expression and leaves the
gradient to autodiff, the returned callable evaluates the loss on the GPU (`asr_contrastive_loss`) and
`loss.with_grads(lv1, lv2)` also returns d loss / d lv1 and d loss / d lv2.
Inputs: (n, 32) float32 CUDA tensors, row i of lv1 matching row i of lv2.
"""
import torch

from .. import _lib


def contrastive_cos_loss(lv1, lv2, weight, gamma, symmetric=False, want_grads=False):
    if not (lv1.is_cuda and lv2.is_cuda):
        raise _lib.AsrError("contrastive_cos_loss needs CUDA tensors (there is no CPU path)")
    lv1 = lv1.contiguous().float()
    lv2 = lv2.contiguous().float()
    if lv1.shape != lv2.shape or lv1.dim() != 2 or lv1.shape[1] != _lib.DIM:
        raise ValueError("expected two (n, %d) code matrices, got %s and %s" % (_lib.DIM, tuple(lv1.shape), tuple(lv2.shape)))
    n = lv1.shape[0]
    scratch = torch.empty(n, dtype=torch.float64, device=lv1.device)
    loss = torch.empty((), dtype=torch.float32, device=lv1.device)
    g1 = torch.empty_like(lv1) if want_grads else None
    g2 = torch.empty_like(lv2) if want_grads else None
    _lib.check(_lib.lib.asr_contrastive_loss(_lib.dptr(lv1), _lib.dptr(lv2), n, float(weight), float(gamma), int(bool(symmetric)),
                                             _lib.dptr(scratch), _lib.dptr(loss), _lib.dptr(g1), _lib.dptr(g2), _lib.stream_ptr()))
    return (loss, g1, g2) if want_grads else loss


def get_contrastive_cos_loss(weight, gamma, symmetric=False):
    """objectives.py:30 -- same factory signature as the reference."""

    def loss(lv1, lv2):
        return contrastive_cos_loss(lv1, lv2, weight, gamma, symmetric)

    def with_grads(lv1, lv2):
        return contrastive_cos_loss(lv1, lv2, weight, gamma, symmetric, want_grads=True)

    loss.with_grads = with_grads
    return loss
