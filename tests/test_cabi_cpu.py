"""CPU: the C-ABI library loads, exports every symbol include/ntm_b200.h declares,
and its host-side logic (shape validation, launch planning) behaves.  No compute
calls -- there is no GPU here."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as entry
from ntm_tracker_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _cabi.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ntm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(ntm_b200_[a-z_0-9]+)\s*\(", src))


def test_every_declared_symbol_is_exported_and_bound(lib):
    declared = header_symbols()
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_status_strings(lib):
    assert lib.ntm_b200_abi_version() == 2
    for code in range(10):
        assert len(lib.ntm_b200_status_string(code)) > 0
    assert b"no CPU fallback" in lib.ntm_b200_status_string(7)


def shape(**kw):
    d = dict(input_dim=514, output_dim=2, mem_size=128, mem_dim=512, shift_range=1,
             controller_hidden_size=200, controller_num_layers=1, write_head_size=1,
             read_head_size=4, write_first=0)
    d.update(kw)
    return _cabi.Shape(**d)


def query(lib, shp, B, T):
    plan = _cabi.Plan()
    st = lib.ntm_b200_query(C.byref(shp), B, T, C.byref(plan))
    return st, plan


def test_plan_tracker_config_uses_cta_pairs(lib):
    """N*M*4 = 256 KiB > 227 KiB: the memory must be split over a 2-CTA cluster."""
    st, plan = query(lib, shape(), 64, 32)
    assert st == 0
    assert plan.cluster_size == 2 and plan.rows_per_cta == 64
    assert plan.threads_per_cta == 512 and plan.ctas_per_sm == 1 and plan.teams == 1
    assert plan.sequences_resident == 64
    assert plan.smem_bytes_per_cta <= 232448
    st, plan = query(lib, shape(), 4096, 64)
    assert plan.sequences_resident == 74          # 148 SMs / 2
    assert plan.workspace_bytes >= 4096 * 64 * 800 * 4


def test_plan_large_memory_uses_8_cta_clusters(lib):
    st, plan = query(lib, shape(mem_size=1024, mem_dim=256), 512, 128)
    assert st == 0 and plan.cluster_size == 8 and plan.rows_per_cta == 128
    assert plan.threads_per_cta == 512 and plan.ctas_per_sm == 1 and plan.teams == 1
    assert plan.smem_bytes_per_cta <= 232448


def test_plan_copy_config_single_cta(lib):
    st, plan = query(lib, shape(input_dim=4, output_dim=4, mem_dim=20, controller_hidden_size=100,
                                read_head_size=1), 16, 20)
    assert st == 0 and plan.cluster_size == 1 and plan.sequences_resident == 16
    assert plan.debug_floats_per_sequence == 2 * 20 + 2 * 3 + 2 * 3 + 2 * 20 + 5 * 2 * 128


@pytest.mark.parametrize("kw,code", [
    (dict(mem_size=0), 1), (dict(controller_num_layers=17), 1), (dict(input_dim=0), 1),
    (dict(shift_range=5), 2), (dict(mem_size=2, shift_range=1), 2),
    (dict(read_head_size=5), 4), (dict(write_head_size=0), 4),
    (dict(mem_size=8192, mem_dim=1024), 5),
])
def test_bad_shapes_are_rejected(lib, kw, code):
    st, _ = query(lib, shape(**kw), 4, 4)
    assert st == code


def test_streaming_only_shapes_are_accepted(lib, monkeypatch):
    """A per-sequence state too large for the persistent kernel (N*M*4 beyond an 8-CTA cluster's shared memory) is
    still a valid shape when the streaming kernels cover it: the query succeeds, reports no resident geometry, sizes
    the workspace for the streaming mode, and the mode choice is "stream" at every batch size."""
    monkeypatch.delenv("NTM_B200_MODE", raising=False)
    shp = shape(mem_size=1024, mem_dim=512)
    st, plan = query(lib, shp, 4, 8)
    assert st == 0 and plan.cluster_size == 0 and plan.sequences_resident == 0 and plan.smem_bytes_per_cta == 0
    assert plan.workspace_bytes > 4 * 8 * 800 * 4 and plan.packed_bytes > 0
    m = C.c_int32(-1)
    for B in (1, 64, 4096):
        assert lib.ntm_b200_query_mode(C.byref(shp), B, C.byref(m)) == 0 and m.value == 1
    monkeypatch.setenv("NTM_B200_MODE", "resident")       # cannot be forced into a kernel that cannot hold it
    assert lib.ntm_b200_query_mode(C.byref(shp), 4, C.byref(m)) == 0 and m.value == 1


def test_null_pointers_and_no_device(lib):
    plan = _cabi.Plan()
    assert lib.ntm_b200_query(None, 1, 1, C.byref(plan)) == 3
    shp = shape()
    import torch
    if not torch.cuda.is_available():
        # compute entry points must fail loudly without a device -- never fall back
        w = _cabi.Weights()
        buf = (C.c_char * 16)()
        st = lib.ntm_b200_pack_weights(C.byref(shp), C.byref(w), buf, 1 << 40, None)
        assert st == 7


def test_python_boundary_mirrors_reference_errors():
    from ntm_tracker_b200 import LoopNTMTracker, NTMCell
    with pytest.raises(ValueError):
        NTMCell(2, mem_size=2, shift_range=2)           # ops.py:231 assert
    cell = NTMCell(4)                                     # the reference's defaults
    assert (cell.mem_size, cell.mem_dim, cell.controller_num_layers,
            cell.read_head_size, cell.write_head_size) == (128, 20, 10, 3, 3)
    names = cell.variable_shapes(6)
    assert names["ntm-tracker/ntm-cell/lstm-controller/cell_0/basic_lstm_cell/weights"] == (6 + 60 + 100, 400)
    assert names["ntm-tracker/ntm-cell/lstm-controller/cell_9/basic_lstm_cell/weights"] == (200, 400)
    assert names["ntm-tracker/ntm-cell/addressing/weights"] == (100, 6 * 20 + 6 * 3 + 6 * 3 + 2 * 20 * 3)
    trk = LoopNTMTracker(5, 2, mem_size=16, mem_dim=8, controller_num_layers=1)
    assert trk.sequence_length == 5 and trk.cell.output_dim == 2


def test_variable_names_match_reference_source(golden_dir):
    """The names the reference graph really creates were recorded by
    oracle/make_golden.py (the shim raises on any unknown name); the oracle's
    table and the package's table must be identical to each other."""
    from ntm_tracker_b200 import NTMCell
    from oracle import ntm_oracle as O
    s = O.NTMShape(output_dim=3, input_dim=5, mem_size=16, mem_dim=8, controller_hidden_size=12,
                   controller_num_layers=2, write_head_size=1, read_head_size=2)
    cell = NTMCell(3, mem_size=16, mem_dim=8, controller_hidden_size=12, controller_num_layers=2,
                   write_head_size=1, read_head_size=2)
    assert cell.variable_shapes(5) == O.param_shapes(s)


def test_workspace_covers_the_streaming_mode(lib):
    """ntm_b200_query sizes the workspace for whichever mode the call will pick: for a large batch it
    must hold the streaming mode's buffers (hoisted projection [B,T,4C], activation rows, K-slice slabs,
    head-parameter rows, operand tiles), and it grows monotonically with the batch."""
    prev = 0
    for B in (64, 512, 4096):
        st, plan = query(lib, shape(), B, 64)
        assert st == 0
        xw = B * 64 * 800 * 4                       # hoisted projection alone
        slabs = 5 * B * 800 * 4                     # K-slice partials of the controller GEMM
        tiles = ((B + 127) // 128) * 36 * 32768     # bf16 hi/lo operand records, K = 2248 -> 36 atoms
        assert plan.workspace_bytes >= xw + slabs + tiles
        assert plan.workspace_bytes > prev
        prev = plan.workspace_bytes


def test_copy_frames_argument_checks(lib):
    """ntm_b200_copy_frames_h2d: null pointers and bad step ranges are rejected before any CUDA call;
    with valid arguments a machine without an sm_100 device answers NO_DEVICE (no CPU path)."""
    buf = (C.c_float * 64)()
    p = C.cast(buf, C.c_void_p)
    assert lib.ntm_b200_copy_frames_h2d(None, p, 1, 4, 4, 0, 2, None) == 3
    assert lib.ntm_b200_copy_frames_h2d(p, p, 1, 4, 4, 2, 2, None) == 1      # empty range
    assert lib.ntm_b200_copy_frames_h2d(p, p, 1, 4, 4, 0, 5, None) == 1      # beyond T
    assert lib.ntm_b200_copy_frames_h2d(p, p, 0, 4, 4, 0, 2, None) == 1
    import torch
    if not torch.cuda.is_available():
        assert lib.ntm_b200_copy_frames_h2d(p, p, 1, 4, 4, 0, 2, None) == 7


def test_time_block_bounds_cover_every_step():
    """Host pipeline of LoopNTMTracker: the automatic blocks are a partition of [0, T) with a short
    first block (only its upload is exposed)."""
    from ntm_tracker_b200 import LoopNTMTracker
    for T in (1, 3, 4, 5, 16, 17, 32, 64, 100):
        b = LoopNTMTracker._auto_bounds(T)
        assert b[0][0] == 0 and b[-1][1] == T
        assert all(lo < hi for lo, hi in b) and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        assert b[0][1] - b[0][0] <= 4


def test_stream_profiling_hooks_answer_without_a_call(lib):
    """The streaming-mode measurement hooks are safe to call when no streaming call ran on this thread."""
    assert _cabi.last_stream_ms()["steps"] == 0
    assert _cabi.last_backward_ms()["steps"] == 0
    assert _cabi.stream_phase_ns()["ctas"] == 0


def test_execution_mode_choice(lib, monkeypatch):
    """The planner's choice between the persistent resident kernel and the streaming kernels (host arithmetic):
    small batches stay resident, batches of many waves stream, shapes the streaming kernels do not cover
    (mem_dim not a multiple of 4) stay resident whatever the batch, and the environment switch wins."""
    monkeypatch.delenv("NTM_B200_MODE", raising=False)
    monkeypatch.delenv("NTM_B200_STREAM_MIN_BATCH", raising=False)

    def mode(shp, B):
        m = C.c_int32(-1)
        assert lib.ntm_b200_query_mode(C.byref(shp), B, C.byref(m)) == 0
        return m.value

    assert mode(shape(), 64) == 0            # BASELINE config 2: one wave of 74 resident sequences
    assert mode(shape(), 74) == 0            # exactly one wave
    assert mode(shape(), 111) == 1           # a second wave: streaming wins from here on (measured, round 2)
    assert mode(shape(controller_num_layers=2), 222) == 0        # fallback streaming kernels: three waves as before
    assert mode(shape(controller_num_layers=2), 223) == 1
    assert mode(shape(), 256) == 1           # BASELINE config 5 (per GPU)
    assert mode(shape(), 4096) == 1          # BASELINE config 3
    assert mode(shape(mem_size=1024, mem_dim=256), 512) == 1     # BASELINE config 4
    assert mode(shape(mem_dim=510), 4096) == 0                   # not covered by the streaming kernels
    monkeypatch.setenv("NTM_B200_MODE", "resident")
    assert mode(shape(), 4096) == 0
    monkeypatch.setenv("NTM_B200_MODE", "stream")
    assert mode(shape(), 8) == 1
    monkeypatch.delenv("NTM_B200_MODE")
    monkeypatch.setenv("NTM_B200_STREAM_MIN_BATCH", "1000")
    assert mode(shape(), 512) == 0 and mode(shape(), 1001) == 1
    assert lib.ntm_b200_query_mode(C.byref(shape()), 0, C.byref(C.c_int32())) == 1


def test_feature_layout_entry_point_argument_checks(lib):
    """ntm_b200_forward_seq_features / ntm_b200_features_workspace_bytes: shape errors come back as codes (or -1),
    and without an sm_100 device the compute call answers NTM_B200_ERR_NO_DEVICE like every other entry point."""
    shp = _cabi.Shape(514, 2, 128, 512, 1, 200, 1, 1, 4, 0)
    base = _cabi.Plan()
    assert lib.ntm_b200_query(C.byref(shp), 8, 2 * 65, C.byref(base)) == 0
    need = lib.ntm_b200_features_workspace_bytes(C.byref(shp), 8, 2, 64)
    assert need >= base.workspace_bytes + 8 * 130 * 514 * 4         # room for the materialised rows as a fallback
    assert lib.ntm_b200_features_workspace_bytes(C.byref(shp), 0, 2, 64) == -1
    assert lib.ntm_b200_features_workspace_bytes(None, 8, 2, 64) == -1
    wts, st = _cabi.Weights(), _cabi.State()
    assert lib.ntm_b200_forward_seq_features(C.byref(shp), C.byref(wts), None, 8, 2, 64, None, None, 0, C.byref(st),
                                             C.byref(st), None, None, None, 0, None) == 3     # null pointers
