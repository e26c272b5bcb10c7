"""ctypes binding of include/ntm_b200.h (the C-ABI shared library).

The library is the product; this file is the thin host-side stub a maintainer
of the reference would add (INTEGRATION.md shows the same stub against the
reference's files).  There is no fallback: if libntm_b200.so is missing or a
call fails, an exception is raised.
"""
import ctypes as C
import os

MAX_LAYERS = 16
_HERE = os.path.dirname(os.path.abspath(__file__))
# NTM_B200_LIB (development only): another build of the same library, for A/B measurements in one process tree
LIB_PATH = os.environ.get("NTM_B200_LIB") or os.path.join(_HERE, "libntm_b200.so")

OK = 0
STATUS_VALUE_ERRORS = (1, 2, 4, 5)   # bad shape / bad shift / heads / too large -> ValueError


class Shape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "input_dim", "output_dim", "mem_size", "mem_dim", "shift_range",
        "controller_hidden_size", "controller_num_layers", "write_head_size",
        "read_head_size", "write_first")]


class Weights(C.Structure):
    _fields_ = [("lstm_w", C.c_void_p * MAX_LAYERS), ("lstm_b", C.c_void_p * MAX_LAYERS),
                ("addr_w", C.c_void_p), ("addr_b", C.c_void_p),
                ("out_w", C.c_void_p), ("out_b", C.c_void_p)]


class State(C.Structure):
    _fields_ = [("M", C.c_void_p), ("w", C.c_void_p), ("read", C.c_void_p),
                ("controller_state", C.c_void_p),
                ("stride_M", C.c_int64), ("stride_w", C.c_int64), ("stride_read", C.c_int64),
                ("stride_controller_state", C.c_int64)]


class History(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("M_prev", "w_prev", "params", "z", "c", "h", "read", "sim", "cn")]


class Grads(C.Structure):
    _fields_ = [("lstm_w", C.c_void_p * MAX_LAYERS), ("lstm_b", C.c_void_p * MAX_LAYERS),
                ("addr_w", C.c_void_p), ("addr_b", C.c_void_p), ("out_w", C.c_void_p), ("out_b", C.c_void_p),
                ("init_M", C.c_void_p), ("init_w", C.c_void_p), ("init_read", C.c_void_p)]


class Plan(C.Structure):
    _fields_ = [("cluster_size", C.c_int32), ("rows_per_cta", C.c_int32),
                ("sequences_resident", C.c_int32), ("threads_per_cta", C.c_int32),
                ("ctas_per_sm", C.c_int32), ("teams", C.c_int32),
                ("smem_bytes_per_cta", C.c_int64), ("workspace_bytes", C.c_int64),
                ("packed_bytes", C.c_int64), ("debug_floats_per_sequence", C.c_int64)]


# every symbol include/ntm_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "ntm_b200_abi_version": (C.c_int32, []),
    "ntm_b200_status_string": (C.c_char_p, [C.c_int32]),
    "ntm_b200_last_cuda_error": (C.c_char_p, []),
    "ntm_b200_query": (C.c_int32, [C.POINTER(Shape), C.c_int64, C.c_int64, C.POINTER(Plan)]),
    "ntm_b200_query_mode": (C.c_int32, [C.POINTER(Shape), C.c_int64, C.POINTER(C.c_int32)]),
    "ntm_b200_pack_weights": (C.c_int32, [C.POINTER(Shape), C.POINTER(Weights), C.c_void_p,
                                          C.c_int64, C.c_void_p]),
    "ntm_b200_forward_seq": (C.c_int32, [C.POINTER(Shape), C.POINTER(Weights), C.c_void_p,
                                         C.c_int64, C.c_int64, C.c_void_p, C.POINTER(State),
                                         C.POINTER(State), C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int64, C.c_void_p]),
    "ntm_b200_forward_seq_train": (C.c_int32, [C.POINTER(Shape), C.POINTER(Weights), C.c_void_p,
                                               C.c_int64, C.c_int64, C.c_void_p, C.POINTER(State),
                                               C.POINTER(State), C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.POINTER(History), C.c_void_p, C.c_int64, C.c_void_p]),
    "ntm_b200_forward_seq_continue": (C.c_int32, [C.POINTER(Shape), C.POINTER(Weights), C.c_void_p,
                                                  C.c_int64, C.c_int64, C.c_void_p, C.POINTER(State),
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ntm_b200_memory_backward_step": (C.c_int32, [C.POINTER(Shape), C.c_int64] + [C.c_void_p] * 9),
    "ntm_b200_serialize_tracker_inputs": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                                      C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ntm_b200_gather_offsets": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_void_p]),
    "ntm_b200_copy_frames_h2d": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                             C.c_int64, C.c_void_p]),
    "ntm_b200_lstm_backward_step": (C.c_int32, [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                                C.c_int64, C.c_void_p]),
    "ntm_b200_offset_loss": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int64, C.c_int64,
                                         C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ntm_b200_backward_workspace_bytes": (C.c_int64, [C.POINTER(Shape), C.c_int64, C.c_int64]),
    "ntm_b200_backward_seq": (C.c_int32, [C.POINTER(Shape), C.POINTER(Weights), C.c_void_p, C.c_int64, C.c_int64,
                                          C.c_void_p, C.POINTER(History), C.c_void_p, C.POINTER(State),
                                          C.POINTER(Grads), C.c_void_p, C.c_int64, C.c_void_p]),
    "ntm_b200_rmsprop_step": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                                          C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "ntm_b200_gemm_nt_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64, C.c_int64, C.c_int32]),
    "ntm_b200_gemm_nt": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "ntm_b200_step": (C.c_int32, [C.POINTER(Shape), C.POINTER(Weights), C.c_void_p, C.c_int64,
                                  C.c_void_p, C.POINTER(State), C.POINTER(State), C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ntm_b200_finish": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "ntm_b200_set_profiling": (C.c_int32, [C.c_int32]),
    "ntm_b200_zero_state": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ntm_b200_features_workspace_bytes": (C.c_int64, [C.POINTER(Shape), C.c_int64, C.c_int32, C.c_int32]),
    "ntm_b200_forward_seq_features": (C.c_int32, [C.POINTER(Shape), C.POINTER(Weights), C.c_void_p, C.c_int64, C.c_int32,
                                                  C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(State),
                                                  C.POINTER(State), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ntm_b200_last_kernel_ms": (C.c_int32, [C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ntm_b200_last_stream_ms": (C.c_int32, [C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "ntm_b200_stream_phase_ns": (C.c_int32, [C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "ntm_b200_last_backward_ms": (C.c_int32, [C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "ntm_b200_phase_cycles": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]),
    "ntm_b200_last_launch_info": (C.c_int32, [C.POINTER(C.c_int32)]),
    "ntm_b200_launch_count": (C.c_int64, []),
}

_lib = None


def load():
    """dlopen libntm_b200.so and bind every declared symbol (raises if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "ntm_tracker_b200: %s not found -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status == OK:
        return
    lib = load()
    msg = "%s: %s" % (what, lib.ntm_b200_status_string(status).decode())
    if status == 8:
        msg += " [%s]" % lib.ntm_b200_last_cuda_error().decode()
    if status in STATUS_VALUE_ERRORS:
        raise ValueError(msg)        # the reference raises ValueError on bad shapes
    raise RuntimeError(msg)


def last_launch_info():
    """{'tensor_path', 'sequences_resident', 'ctas', 'cluster_size', ...} of this thread's last launch."""
    buf = (C.c_int32 * 16)()
    check(load().ntm_b200_last_launch_info(buf), "last_launch_info")
    keys = ("tensor_path", "sequences_resident", "ctas", "cluster_size", "ks_ctrl", "kw_ctrl", "ks_heads",
            "kw_heads", "teams", "threads_per_cta", "ctas_per_sm", "smem_bytes_per_cta", "xproj_tensor_path", "streaming", "continued")
    return dict(zip(keys, list(buf)))


def last_stream_ms():
    """Streaming mode, profiling enabled: device ms of the last call summed over its timesteps
    ({'controller', 'head_params', 'memory', 'init', 'steps'}); steps == 0 if it was not a streaming call."""
    buf = (C.c_float * 4)()
    steps = C.c_int32(0)
    check(load().ntm_b200_last_stream_ms(buf, C.byref(steps)), "last_stream_ms")
    return {"controller": buf[0], "head_params": buf[1], "memory": buf[2], "init": buf[3], "steps": steps.value}


def last_backward_ms():
    """Training, profiling enabled: device ms of the last ntm_b200_backward_seq
    ({'memory_backward', 'loop_rest', 'weight_grads', 'total', 'steps'}); steps == 0 if nothing was recorded."""
    buf = (C.c_float * 4)()
    steps = C.c_int32(0)
    check(load().ntm_b200_last_backward_ms(buf, C.byref(steps)), "last_backward_ms")
    return {"memory_backward": buf[0], "loop_rest": buf[1], "weight_grads": buf[2], "total": buf[3], "steps": steps.value}


def stream_phase_ns():
    """Streaming mode, profiling enabled: mean ns per phase of the memory kernel's CTAs in the last launch."""
    buf = (C.c_double * 12)()
    n = C.c_int32(0)
    check(load().ntm_b200_stream_phase_ns(buf, C.byref(n)), "stream_phase_ns")
    keys = ("wait_params", "activations", "pass1", "addressing", "pass2", "store_drain", "finalize", "cta", "launch_span",
            "p1_wait_cycles", "p1_compute_cycles", "p1_sync_issue_cycles")
    d = dict(zip(keys, list(buf)))
    d["ctas"] = n.value
    return d
