"""ResidentTracker -- the serve path's unit of work with the state kept on the device
(SURVEY.md s8f rank 3).

The reference's VOT tracker (test_tracker.py:104-299) crops a frame, runs VGG, builds 65 rows
``[delimiter row first | 64 feature rows]`` (``_preprocess_image``, :392-404) and then makes 65
``sess.run`` calls, one NTM cell step each, with the whole state (M, w, read, controller_state)
fed and fetched as NumPy every step (``_run_tracker``, :284-299); the last row's two outputs are the
(dy, dx) offsets.  Here a frame is ONE persistent-kernel launch of F+1 steps and the state never
leaves the GPU between frames.
"""
import torch

from .ntm_cell import NTMCell


class ResidentTracker(object):
    def __init__(self, cell: NTMCell, num_features=64, batch_size=1):
        self.cell = cell
        self.num_features = int(num_features)
        self.batch_size = int(batch_size)
        self.state = None
        self.frame_index = 0
        self._spare = [None, None]     # two sets of state buffers, written alternately (no allocation per frame)
        self._zero_target = None

    def reset(self, state=None):
        """Start a new sequence (test_tracker.py:146: ``states=[sess.run(zero_state)]``)."""
        self.state = state if state is not None else self.cell.zero_state(self.batch_size)
        self.frame_index = 0
        return self

    def track(self, features, target=None):
        """features [B, F, Cch] of one frame (CUDA); target [B, F] ground-truth map, given for the
        first frame only (zeros afterwards, test_tracker.py:398-399).  Returns the frame's offsets
        tanh(logit) of the last step, [B, O]."""
        if self.state is None:
            self.reset()
        if features.dim() != 3 or features.shape[0] != self.batch_size or features.shape[1] != self.num_features:
            raise ValueError("expected features [%d, %d, C], got %s"
                             % (self.batch_size, self.num_features, tuple(features.shape)))
        B, F, _ = features.shape
        if target is None:
            if self._zero_target is None or self._zero_target.device != features.device:
                self._zero_target = torch.zeros(B, F, device=features.device)
            target = self._zero_target
        slot = self.frame_index & 1
        if self._spare[slot] is None:
            self._spare[slot] = self.cell.state_placeholder(B)
        # feature-layout call: the library builds the [delimiter | 64 feature rows] steps itself (serve layout)
        logits, _, self.state = self.cell._run_features(features.unsqueeze(1), target, self.state, delimiter_first=True,
                                                        out_state=self._spare[slot])
        self.frame_index += 1
        return torch.tanh(logits[:, -1])
