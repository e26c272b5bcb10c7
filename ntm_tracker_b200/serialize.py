"""The data formats either side of the NTM path, on the device.

``tracker_inputs`` replaces the tf.concat / tf.tile / tf.reshape chain of
direct_offset_output.py:439-500 (training layout: delimiter row LAST in every frame) and of
test_tracker.py:392-404 (serve layout: delimiter row FIRST); ``gather_offsets`` replaces the
slice / reshape / tanh of direct_offset_output.py:581-593.  Both are single HBM-bound kernels
behind the C ABI (csrc/ntm_b200_io.cu).
"""
import ctypes as C

import torch

from . import _cabi


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def tracker_inputs(features, target, delimiter_first=False):
    """features [B, L, F, Cch] (conv4_3 vectors at the F sampled points of each of L frames),
    target [B, F] (first-frame ground-truth map) -> inputs [B, L*(F+1), Cch+2]."""
    if features.dim() != 4 or target.dim() != 2 or target.shape != (features.shape[0], features.shape[2]):
        raise ValueError("expected features [B,L,F,C] and target [B,F], got %s and %s"
                         % (tuple(features.shape), tuple(target.shape)))
    if not features.is_cuda:
        raise RuntimeError("ntm_tracker_b200 needs CUDA tensors; there is no CPU fallback")
    features = features.float().contiguous()
    target = target.to(features.device, torch.float32).contiguous()
    B, L, F, Cch = features.shape
    out = torch.empty(B, L * (F + 1), Cch + 2, dtype=torch.float32, device=features.device)
    _cabi.check(_cabi.load().ntm_b200_serialize_tracker_inputs(
        features.data_ptr(), target.data_ptr(), out.data_ptr(), B, L, F, Cch, int(bool(delimiter_first)),
        _stream(features.device)), "serialize_tracker_inputs")
    return out


def gather_offsets(output_logits, num_features):
    """output_logits [B, L*(F+1), O] -> tanh of the logits at each frame's delimiter step, first
    frame dropped: [B, L-1, O] (the (dy, dx) offset predictions)."""
    if output_logits.dim() != 3 or output_logits.shape[1] % (num_features + 1) != 0:
        raise ValueError("logits of shape %s do not hold whole frames of %d rows"
                         % (tuple(output_logits.shape), num_features + 1))
    if not output_logits.is_cuda:
        raise RuntimeError("ntm_tracker_b200 needs CUDA tensors; there is no CPU fallback")
    lg = output_logits.float().contiguous()
    B, T, O = lg.shape
    L = T // (num_features + 1)
    if L < 2:
        raise ValueError("need at least two frames")
    out = torch.empty(B, L - 1, O, dtype=torch.float32, device=lg.device)
    _cabi.check(_cabi.load().ntm_b200_gather_offsets(lg.data_ptr(), out.data_ptr(), B, L, num_features, O,
                                                     _stream(lg.device)), "gather_offsets")
    return out
