"""ctypes binding of the C ABI declared in include/mcan_b200.h.

No torch types cross this boundary: only raw device pointers (ints), sizes and a
cudaStream_t.  There is no CPU fallback -- if the shared library cannot be loaded every
entry point raises.
"""
import ctypes
import threading
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmcan_b200.so")

MAX_SEG = 3

c_void_p = ctypes.c_void_p
c_int32 = ctypes.c_int32
c_int64 = ctypes.c_int64
c_uint32 = ctypes.c_uint32
c_float = ctypes.c_float


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("a", c_void_p * MAX_SEG),
        ("b", c_void_p * MAX_SEG),
        ("num_seg", c_int32),
        ("a_layout", c_int32),
        ("b_layout", c_int32),
        ("m", c_int64),
        ("n", c_int64),
        ("k", c_int64),
        ("lda", c_int64),
        ("ldb", c_int64),
        ("bias", c_void_p),
        ("relu", c_int32),
        ("dropout_p", c_float),
        ("dropout_seed", c_uint32),
        ("dropout_seed_dev", c_void_p),
        ("gate", c_void_p),
        ("ldg", c_int64),
        ("gate_scale", c_float),
        ("resid", c_void_p),
        ("ldr", c_int64),
        ("out_f32", c_void_p),
        ("ldo_f32", c_int64),
        ("out_bf16", c_void_p),
        ("out_bf16_lo", c_void_p),
        ("ldo_bf16", c_int64),
        ("colsum", c_void_p),
        ("accumulate", c_int32),
        ("split_k", c_int32),
        ("block_n", c_int32),
        ("cta_group", c_int32),
        ("stream", c_void_p),
    ]


MAX_GROUPS = 8


class GemmGroup(ctypes.Structure):
    _fields_ = [
        ("a", c_void_p),
        ("b", c_void_p),
        ("m", c_int64),
        ("n", c_int64),
        ("lda", c_int64),
        ("ldb", c_int64),
        ("out", c_void_p),
        ("ldo", c_int64),
    ]


class GemmGroupedArgs(ctypes.Structure):
    _fields_ = [
        ("g", GemmGroup * MAX_GROUPS),
        ("num_groups", c_int32),
        ("split_k", c_int32),
        ("accumulate", c_int32),
        ("out_bf16", c_int32),
        ("k", c_int64),
        ("stream", c_void_p),
    ]


class GemmLnArgs(ctypes.Structure):
    _fields_ = [
        ("a", c_void_p),
        ("b", c_void_p),
        ("m", c_int64),
        ("n", c_int64),
        ("k", c_int64),
        ("lda", c_int64),
        ("ldb", c_int64),
        ("bias", c_void_p),
        ("dropout_p", c_float),
        ("dropout_seed", c_uint32),
        ("dropout_seed_dev", c_void_p),
        ("resid", c_void_p),
        ("ldr", c_int64),
        ("ln_a2", c_void_p),
        ("ln_b2", c_void_p),
        ("eps", c_float),
        ("s_f32", c_void_p),
        ("y_f32", c_void_p),
        ("y_bf16", c_void_p),
        ("mean", c_void_p),
        ("sigma", c_void_p),
        ("stream", c_void_p),
    ]


class LstmArgs(ctypes.Structure):
    _fields_ = [
        ("xw", c_void_p),
        ("w_hh", c_void_p),
        ("b_hh", c_void_p),
        ("batch", c_int32),
        ("steps", c_int32),
        ("hidden", c_int32),
        ("hbuf", c_void_p),
        ("h_out", c_void_p),
        ("cbuf", c_void_p),
        ("gates", c_void_p),
        ("dout", c_void_p),
        ("da", c_void_p),
        ("barrier", c_void_p),
        ("stream", c_void_p),
    ]


class AttnArgs(ctypes.Structure):
    _fields_ = [
        ("q", c_void_p),
        ("k", c_void_p),
        ("v", c_void_p),
        ("ldq", c_int64),
        ("ldk", c_int64),
        ("ldv", c_int64),
        ("key_mask", c_void_p),
        ("out", c_void_p),
        ("ldo", c_int64),
        ("batch", c_int32),
        ("heads", c_int32),
        ("sq", c_int32),
        ("sk", c_int32),
        ("head_dim", c_int32),
        ("scale", c_float),
        ("dropout_p", c_float),
        ("dropout_seed", c_uint32),
        ("dropout_seed_dev", c_void_p),
        ("q_lo", c_void_p),
        ("k_lo", c_void_p),
        ("v_lo", c_void_p),
        ("out_lo", c_void_p),
        ("stream", c_void_p),
    ]


class AttnBwdArgs(ctypes.Structure):
    _fields_ = [
        ("fwd", AttnArgs),
        ("dout", c_void_p),
        ("lddo", c_int64),
        ("dq", c_void_p),
        ("dk", c_void_p),
        ("dv", c_void_p),
        ("lddq", c_int64),
        ("lddk", c_int64),
        ("lddv", c_int64),
        ("dbq", c_void_p),
        ("dbk", c_void_p),
        ("dbv", c_void_p),
    ]


# every symbol include/mcan_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mcan_version": (ctypes.c_int, []),
    "mcan_last_error": (ctypes.c_char_p, []),
    "mcan_num_sms": (ctypes.c_int, []),
    "mcan_set_sm_limit": (ctypes.c_int, [ctypes.c_int]),
    "mcan_set_gemm_schedule": (ctypes.c_int, [ctypes.c_int]),
    "mcan_set_pdl": (ctypes.c_int, [ctypes.c_int]),
    "mcan_set_attn_impl": (ctypes.c_int, [ctypes.c_int]),
    "mcan_gemm": (ctypes.c_int, [ctypes.POINTER(GemmArgs)]),
    "mcan_gemm_ln": (ctypes.c_int, [ctypes.POINTER(GemmLnArgs)]),
    "mcan_gemm_grouped": (ctypes.c_int, [ctypes.POINTER(GemmGroupedArgs)]),
    "mcan_attn_fwd": (ctypes.c_int, [ctypes.POINTER(AttnArgs)]),
    "mcan_attn_bwd": (ctypes.c_int, [ctypes.POINTER(AttnBwdArgs)]),
    "mcan_embed_gather": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32,
                                         c_void_p, c_void_p]),
    "mcan_embed_scatter_add": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                              c_void_p]),
    "mcan_lstm_fwd": (ctypes.c_int, [ctypes.POINTER(LstmArgs)]),
    "mcan_lstm_bwd": (ctypes.c_int, [ctypes.POINTER(LstmArgs)]),
    "mcan_layernorm_fwd": (ctypes.c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_float,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcan_layernorm_add_fwd": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_float,
                                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcan_rowmask_cast": (ctypes.c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "mcan_sigmoid_bce_fwd": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                            c_void_p]),
    "mcan_sigmoid_bce_bwd": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int64,
                                            c_void_p, c_void_p]),
    "mcan_layernorm_bwd": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                          c_int64, c_int64, c_void_p, c_void_p, c_float, c_uint32,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcan_attflat_pool_fwd": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                             c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                             c_void_p, c_void_p]),
    "mcan_attflat_pool_bwd": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_int32, c_int32, c_int32, c_int32, c_int32, c_float,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcan_cast_bf16": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "mcan_cast_multi": (ctypes.c_int, [c_void_p, c_int32, c_int64, c_void_p]),
    "mcan_gemm_plan": (ctypes.c_int, [c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "mcan_adamw_multi": (ctypes.c_int, [c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_float, c_float, c_float,
                                        c_float, c_int32, c_void_p]),
    "mcan_debug_hog": (ctypes.c_int, [c_int32, c_int64, c_int32, c_void_p]),
    "mcan_gate_bf16": (ctypes.c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int64, c_void_p]),
    "mcan_colsum_bf16": (ctypes.c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "mcan_colsum_f32": (ctypes.c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
}

_lib = None


class McanError(RuntimeError):
    pass


def load():
    """Loads libmcan_b200.so (once).  Raises McanError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McanError(
            "libmcan_b200.so not found at %s -- build it with `python mcan-vqa_b200/build.py` "
            "(there is no CPU or PyTorch fallback for the MCAN hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    if os.environ.get("MCAN_PDL", "1") == "0":      # debugging / A-B timing: plain stream-ordered launches
        lib.mcan_set_pdl(0)
    _lib = lib
    return lib


class StructPacker(object):
    """Fills a ctypes.Structure with ONE struct.pack_into call instead of one attribute store per field (~0.2 us each,
    ~30 fields for mcan_gemm_args: the marshalling was half of the Python cost of a launch on the eager route).
    The format string is derived from the Structure's own _fields_ (native alignment), so the two layouts cannot drift
    apart; tests/test_capi_cpu.py compares the packed bytes with a field-by-field fill.  Values are given in field order,
    arrays flattened, None pointers as 0.  The buffer is reused: the C ABI reads its arguments during the call only."""

    _CODES = {ctypes.c_void_p: "P", ctypes.c_int32: "i", ctypes.c_uint32: "I", ctypes.c_int64: "q", ctypes.c_float: "f",
              ctypes.c_uint8: "B"}

    def __init__(self, struct_type):
        import struct as _struct
        fmt = "@"
        self.count = 0
        for _, ctype in struct_type._fields_:
            n = 1
            if hasattr(ctype, "_length_"):
                n, ctype = ctype._length_, ctype._type_
            fmt += self._CODES[ctype] * n
            self.count += n
        fmt += "0P"                                    # trailing padding up to pointer alignment, as in C
        self.packer = _struct.Struct(fmt)
        if self.packer.size != ctypes.sizeof(struct_type):
            raise McanError("StructPacker: %s is %d bytes in ctypes, %d packed" %
                            (struct_type.__name__, ctypes.sizeof(struct_type), self.packer.size))
        self._tls = threading.local()
        self.struct_type = struct_type

    def pack(self, *values):
        """-> a ctypes pointer to the filled structure (valid until the next pack() on this thread)."""
        tls = self._tls
        buf = getattr(tls, "buf", None)
        if buf is None:
            buf = tls.buf = ctypes.create_string_buffer(self.packer.size)
            tls.ptr = ctypes.cast(buf, ctypes.POINTER(self.struct_type))
        self.packer.pack_into(buf, 0, *values)
        return tls.ptr


launch_count = 0   # successful C-ABI calls == kernels enqueued by this process


def check(rc, what, launch=True):
    """Raises McanError for a non-zero return code; counts the call as a kernel launch unless told otherwise
    (configuration / query entry points enqueue nothing)."""
    global launch_count
    if launch:
        launch_count += 1
    if rc != 0:
        msg = load().mcan_last_error()
        raise McanError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
