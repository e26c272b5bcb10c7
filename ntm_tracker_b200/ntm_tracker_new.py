"""LoopNTMTracker -- drop-in for the reference's batched sequence driver
(ntm_tracker_new.py:4-64): unrolls the NTM cell over the T frames-steps of B
independent sequences.  The reference does it with tf.while_loop + TensorArrays
(one graph dispatch per op per step); here the whole loop is ONE persistent CUDA
kernel call through the C ABI (ntm_b200_forward_seq).
"""
import numpy as np
import torch

from .ntm_cell import NTMCell, random_uniform_initializer


class LoopNTMTracker(object):
    def __init__(self, sequence_length, output_dim,
                 initializer=random_uniform_initializer(-.1, .1), **kwargs):
        self.cell = NTMCell(output_dim, **kwargs)
        self.initializer = initializer
        self.sequence_length = sequence_length
        self.final_state = None

    def __call__(self, inputs, state=None, scope=None):
        """inputs [B, T, D] batch-major -> (outputs [B,T,O], output_logits [B,T,O])
        (ntm_tracker_new.py:42-49).  T must equal ``sequence_length``.  Host
        inputs (NumPy / CPU tensors, ideally pinned) are copied to the device and
        the results are returned on the host, like a ``sess.run`` of the reference
        graph; CUDA inputs stay on the device.  The final state dict is kept in
        ``self.final_state`` (the reference's loop_vars M, w, read, controller_state)."""
        host_in = isinstance(inputs, np.ndarray) or not inputs.is_cuda
        as_numpy = isinstance(inputs, np.ndarray)
        x = self.cell._prepare_inputs(inputs, 3)
        B, T, D = x.shape
        if T != self.sequence_length:
            raise ValueError("inputs have %d steps but sequence_length is %d" % (T, self.sequence_length))
        if self.cell.input_dim is None:
            self.cell.build(D, self.initializer)
        state = state or self.cell.zero_state(B, self.initializer)
        logits, outputs, new_state, _ = self.cell._run(x, state, T)
        self.final_state = new_state
        if host_in:
            outputs, logits = outputs.cpu(), logits.cpu()
            if as_numpy:
                outputs, logits = outputs.numpy(), logits.numpy()
        return (outputs, logits)
