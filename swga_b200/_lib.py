"""ctypes binding of libswga_b200.so (include/swga_b200.h).  There is no CPU fallback: a missing
library or a missing CUDA device raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libswga_b200.so")

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_vp = ctypes.c_void_p
c_ctx = ctypes.c_void_p

# name -> (restype, argtypes).  Every symbol include/swga_b200.h declares is listed here; the CPU
# test tests/test_abi_cpu.py checks the two stay in sync.
SIGNATURES = {
    "swga_abi_version": (ctypes.c_int, []),
    "swga_last_error": (ctypes.c_char_p, []),
    "swga_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_ctx)]),
    "swga_ctx_destroy": (None, [c_ctx]),
    "swga_ctx_sm_count": (ctypes.c_int, [c_ctx]),
    "swga_ctx_set_count_impl": (ctypes.c_int, [c_ctx, ctypes.c_int]),
    "swga_malloc": (ctypes.c_int, [c_ctx, ctypes.POINTER(c_vp), ctypes.c_size_t]),
    "swga_free": (ctypes.c_int, [c_ctx, c_vp]),
    "swga_host_alloc": (ctypes.c_int, [c_ctx, ctypes.POINTER(c_vp), ctypes.c_size_t]),
    "swga_host_free": (ctypes.c_int, [c_ctx, c_vp]),
    "swga_memset": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_int, ctypes.c_size_t, c_vp]),
    "swga_h2d": (ctypes.c_int, [c_ctx, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "swga_d2h": (ctypes.c_int, [c_ctx, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "swga_stream_sync": (ctypes.c_int, [c_ctx, c_vp]),
    "swga_tables_words": (ctypes.c_uint64, [ctypes.c_int, ctypes.c_int]),
    "swga_table_offset": (ctypes.c_uint64, [ctypes.c_int, ctypes.c_int]),
    "swga_packed_words": (ctypes.c_uint64, [ctypes.c_uint64]),
    "swga_brk_words": (ctypes.c_uint64, [ctypes.c_uint64]),
    "swga_fasta_scan": (ctypes.c_int, [c_vp, ctypes.c_uint64, c_vp, c_vp, c_vp, c_u64p, c_u64p]),
    "swga_dsk_mode": (ctypes.c_int, [c_vp, ctypes.c_uint64]),
    "swga_pack": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_uint64, ctypes.c_int, c_vp, c_vp, c_vp]),
    "swga_mark_records": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_uint64, ctypes.c_uint64, c_vp, c_vp]),
    "swga_count_accumulate": (ctypes.c_int, [c_ctx, c_vp, c_vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                             c_vp, c_vp]),
    "swga_count_accumulate_ascii": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_uint64, c_vp, ctypes.c_uint64, ctypes.c_int,
                                                   ctypes.c_uint64, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, c_vp]),
    "swga_ctx_set_count_via_pack": (ctypes.c_int, [c_ctx, ctypes.c_int]),
    "swga_count_partition_stats": (ctypes.c_int, [c_ctx, c_u64p]),
    "swga_ctx_set_stage_timing": (ctypes.c_int, [c_ctx, ctypes.c_int]),
    "swga_count_stage_ms": (ctypes.c_int, [c_ctx, ctypes.POINTER(ctypes.c_float)]),
    "swga_count_finalize": (ctypes.c_int, [c_ctx, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "swga_fold_canonical": (ctypes.c_int, [c_ctx, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp]),
    "swga_merge_supported": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "swga_merge_finalize": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]),
    "swga_merge_finalize_multicast": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]),
    "swga_count_genome_device": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_uint64, c_vp, ctypes.c_uint64, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_int,
                                                c_vp, c_vp, c_vp]),
    "swga_count_genome_host": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_uint64, c_vp, ctypes.c_uint64, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, c_vp]),
    "swga_count_stream_host": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_uint64, c_vp, ctypes.c_uint64, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_uint64, c_vp]),
    "swga_ctx_compute_stream": (c_vp, [c_ctx]),
    "swga_count_fasta_stream": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               c_vp, c_u64p, c_u64p, ctypes.POINTER(ctypes.c_int)]),
    "swga_count_fasta_host": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             c_vp, c_u64p, c_u64p, ctypes.POINTER(ctypes.c_int)]),
    "swga_filter_primers": (ctypes.c_int, [c_ctx, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                           ctypes.c_int64, ctypes.c_int, ctypes.c_uint32, c_vp, ctypes.c_uint32,
                                           c_vp, c_vp]),
    "swga_build_edges": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_uint32, ctypes.c_uint32, c_vp, ctypes.c_uint32, ctypes.c_int,
                                        c_vp, ctypes.c_uint32, c_vp]),
    "swga_parse_set_lines": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, c_vp, c_vp,
                                            c_vp, c_u64p, c_u64p]),
    "swga_write_dsk_file": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, c_vp, ctypes.c_uint32, c_u64p]),
    "swga_index_offsets_words": (ctypes.c_uint64, [ctypes.c_int]),
    "swga_build_index": (ctypes.c_int, [c_ctx, c_vp, c_vp, ctypes.c_uint64, ctypes.c_int, c_vp, c_vp, c_u64p, c_vp]),
    "swga_locate_counts": (ctypes.c_int, [c_ctx, ctypes.c_int, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp]),
    "swga_locate_gather": (ctypes.c_int, [c_ctx, ctypes.c_int, c_vp, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp]),
    "swga_union_strands": (ctypes.c_int, [c_ctx, c_vp, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp]),
    "swga_tile_offsets": (ctypes.c_int, [c_ctx, c_vp, c_vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, c_vp, c_vp]),
    "swga_ctx_set_score_sort_cap": (ctypes.c_int, [c_ctx, ctypes.c_uint32]),
    "swga_ctx_set_score_dense_impl": (ctypes.c_int, [c_ctx, ctypes.c_int]),
    "swga_ctx_set_score_sort_impl": (ctypes.c_int, [c_ctx, ctypes.c_int]),
    "swga_score_tile_shift": (ctypes.c_uint32, []),
    "swga_score_max_hist_bins": (ctypes.c_uint32, []),
    "swga_score_sets": (ctypes.c_int, [c_ctx, c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint32, c_vp, ctypes.c_uint32,
                                       ctypes.c_uint32, c_vp, ctypes.c_uint32, ctypes.c_int,
                                       c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "swga_gini_unbounded": (ctypes.c_int, [c_ctx, c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint32, c_vp, ctypes.c_uint32,
                                           ctypes.c_uint64, c_f64p, c_vp]),
    "swga_synth_bases": (ctypes.c_int, [c_ctx, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                                        ctypes.c_uint32, ctypes.c_uint32, c_vp, c_vp]),
    "swga_synth_sets": (ctypes.c_int, [c_ctx, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, c_vp, c_vp]),
    "swga_launch_count": (ctypes.c_uint64, []),
    "swga_bench_red_global": (ctypes.c_int, [c_ctx, c_vp, ctypes.c_uint64, ctypes.c_uint64,
                                             ctypes.POINTER(ctypes.c_float)]),
    "swga_bench_atom_shared": (ctypes.c_int, [c_ctx, ctypes.c_uint32, ctypes.c_uint64,
                                              ctypes.POINTER(ctypes.c_float)]),
}

_lib = None


class SwgaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libswga_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """dlopen libswga_b200.so and type its symbols.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m swga_b200.build` (nvcc, sm_100a). "
                "swga_b200 has no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise SwgaError(rc, load().swga_last_error().decode("utf-8", "replace"))


def ptr(x):
    """void* of a numpy array / torch tensor / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return ctypes.c_void_p(x.data_ptr())
    if hasattr(x, "ctypes"):
        return ctypes.c_void_p(x.ctypes.data)
    raise TypeError(type(x))


class Context:
    """swga_ctx for one device.  Creating it without a usable B200 raises SwgaError."""

    def __init__(self, device=0):
        self.lib = load()
        self.device = device
        h = c_ctx()
        check(self.lib.swga_ctx_create(int(device), ctypes.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.swga_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self):
        return self.lib.swga_ctx_sm_count(self.h)


_contexts = {}


def context(device=None):
    """Process-wide context per device (device defaults to torch's current CUDA device)."""
    if device is None:
        import torch
        if not torch.cuda.is_available():
            raise SwgaError(-1, "no CUDA device: swga_b200 runs its hot path on a B200 only")
        device = torch.cuda.current_device()
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
