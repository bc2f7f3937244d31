"""ctypes binding of libpysdr_b200.so (include/pysdr_b200.h).  No pybind, no torch types in signatures.

The library is REQUIRED: there is no CPU fallback.  Import of this module succeeds without a GPU (so the
symbol-export test can run on CPU); any compute call without a CUDA device fails loudly in the CUDA runtime.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpysdr_b200.so")

c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_vp = ctypes.c_void_p
c_int = ctypes.c_int
c_dbl = ctypes.c_double


class PysdrError(RuntimeError):
    pass


PSD_RAW, PSD_FLIP = 1, 2          # include/pysdr_b200.h PYSDR_PSD_*


class BankConfig(ctypes.Structure):
    _fields_ = [("srate", c_dbl), ("up", ctypes.c_int32), ("down", ctypes.c_int32), ("in_chunk", c_i64),
                ("n_rx", ctypes.c_int32), ("filt_len", ctypes.c_int32), ("af_len", ctypes.c_int32),
                ("max_in", c_i64)]


# name -> (restype, argtypes); every symbol include/pysdr_b200.h declares
SIGNATURES = {
    "pysdr_last_error": (ctypes.c_char_p, []),
    "pysdr_version": (c_int, []),
    "pysdr_freq_to_phase_inc": (c_u64, [c_dbl, c_dbl]),
    "pysdr_phase_inc_to_freq": (c_dbl, [c_u64, c_dbl]),
    "pysdr_quad_mixer": (c_int, [c_vp, c_vp, c_i64, c_u64, c_u64, c_vp]),
    "pysdr_cs16_to_cf32": (c_int, [c_vp, c_vp, c_i64, c_dbl, c_vp]),
    "pysdr_mean_power": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "pysdr_fir_valid": (c_int, [c_vp, c_int, c_vp, c_int, c_i64, c_vp, c_vp]),
    "pysdr_bank_create": (c_int, [ctypes.POINTER(BankConfig), ctypes.POINTER(c_vp)]),
    "pysdr_bank_destroy": (c_int, [c_vp]),
    "pysdr_bank_reset": (c_int, [c_vp]),
    "pysdr_bank_set_stereo": (c_int, [c_vp, c_int, c_dbl]),
    "pysdr_bank_pll_reset": (c_int, [c_vp, c_int]),
    "pysdr_bank_pll_get": (c_int, [c_vp, c_int, c_vp, c_vp]),
    "pysdr_bank_set_lo": (c_int, [c_vp, c_int, c_u64]),
    "pysdr_bank_set_dec_taps": (c_int, [c_vp, c_int, c_vp, c_int]),
    "pysdr_bank_set_demod": (c_int, [c_vp, c_int, c_int, c_vp, c_int, c_int, c_u64]),
    "pysdr_bank_agc_reset": (c_int, [c_vp, c_int]),
    "pysdr_bank_agc_config": (c_int, [c_vp, c_int, c_dbl, c_dbl]),
    "pysdr_bank_agc_get": (c_int, [c_vp, c_int, ctypes.POINTER(c_dbl), c_vp]),
    "pysdr_bank_n_out": (c_i64, [c_vp, c_i64]),
    "pysdr_bank_position": (c_i64, [c_vp]),
    "pysdr_bank_n_blocks": (c_i64, [c_vp, c_i64]),
    "pysdr_bank_process": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(c_i64), c_vp]),
    "pysdr_bank_process_host": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp),
                                        ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_vp]),
    "pysdr_bank_host_chunk_ptr": (c_int, [c_vp, ctypes.POINTER(c_vp)]),
    "pysdr_bank_process_front": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, ctypes.POINTER(c_i64), c_vp]),
    "pysdr_bank_process_back": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "pysdr_bank_seek": (c_int, [c_vp, c_i64, c_vp]),
    "pysdr_bank_set_timing": (c_int, [c_vp, c_int]),
    "pysdr_bank_get_timing": (c_int, [c_vp, ctypes.POINTER(c_dbl), c_vp]),
    "pysdr_bank_state_size": (c_i64, [c_vp]),
    "pysdr_bank_get_state": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "pysdr_bank_set_state": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "pysdr_bank_k1_variant": (c_int, [c_vp]),
    "pysdr_bank_set_k1_mma": (c_int, [c_vp, c_int]),
    "pysdr_bank_k1_mma_available": (c_int, [c_vp]),
    "pysdr_bank_k1_last": (c_int, [c_vp]),
    "pysdr_k1chan_debug_plan": (c_i64, [c_int, c_int, c_int, c_int, c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_u64, c_i64, c_vp, c_vp, c_i64]),
    "pysdr_bank_force_generic": (c_int, [c_vp, c_int]),
    "pysdr_bank_set_k1_only": (c_int, [c_vp, c_int]),
    "pysdr_bank_set_real_input": (c_int, [c_vp, c_int]),
    "pysdr_fm_disc": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "pysdr_fir_spectrum": (c_int, [c_vp, c_int, c_vp, c_vp]),
    "pysdr_wfm_video_disc": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_vp, c_int, c_u64, c_u64, c_vp, c_vp]),
    "pysdr_bank_force_direct_fir": (c_int, [c_vp, c_int]),
    "pysdr_bank_adopt_c_memory": (c_int, [c_vp, c_vp, c_i64]),
    "pysdr_bank_set_k1_external": (c_int, [c_vp, c_int]),
    "pysdr_bank_c_memory": (c_int, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), ctypes.POINTER(ctypes.c_int32)]),
    "pysdr_bank_launch_count": (c_i64, [c_vp]),
    "pysdr_bank_agc_summary": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "pysdr_bank_agc_enter": (c_int, [c_vp, c_vp, c_int, c_vp]),
    "pysdr_bank_process_back_carry": (c_int, [c_vp, c_vp, c_int, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "pysdr_bank_force_unfused": (c_int, [c_vp, c_int]),
    "pysdr_xchg_bytes": (c_i64, [c_int, c_int]),
    "pysdr_bank_agc_summary_push": (c_int, [c_vp, c_i64, c_vp, c_int, c_int, c_u64, c_vp]),
    "pysdr_bank_process_back_xchg": (c_int, [c_vp, c_vp, c_int, c_int, c_u64, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "pysdr_bank_process_shard_xchg": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_int, c_int, c_u64, c_i64, c_vp, c_vp, c_i64,
                                              ctypes.POINTER(c_i64), c_vp]),
    "pysdr_find_peaks": (c_int, [c_vp, ctypes.c_int32, c_vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_vp]),
    "pysdr_bank_agc_trace": (c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(c_i64), c_vp]),
    "pysdr_lfilter_set_mode": (c_int, [c_int]),
    "pysdr_lfilter": (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_i64, c_int, c_i64, c_vp, c_vp]),
    "pysdr_abs_f32": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "pysdr_ratio_f32": (c_int, [c_vp, c_vp, c_vp, ctypes.c_float, c_i64, c_vp]),
    "pysdr_psd_create": (c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_vp, ctypes.POINTER(c_vp)]),
    "pysdr_psd_destroy": (c_int, [c_vp]),
    "pysdr_psd_lines": (c_int, [c_vp, c_vp, c_i64, c_int, ctypes.c_int32, c_int, c_vp, ctypes.POINTER(c_i64), c_vp]),
    "pysdr_psd_launch_count": (c_i64, [c_vp]),
    "pysdr_waterfall_rgba": (c_int, [c_vp, c_i64, c_vp, c_vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_float, c_vp, c_vp, c_vp]),
    "pysdr_fft_pos_to_freq": (c_int, [c_int, c_int]),
    "pysdr_czt_create": (c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "pysdr_czt_destroy": (c_int, [c_vp]),
    "pysdr_czt_lines": (c_int, [c_vp, c_vp, c_i64, c_int, ctypes.c_int32, c_int, c_vp, ctypes.POINTER(c_i64), c_vp]),
    "pysdr_czt_launch_count": (c_i64, [c_vp]),
    "pysdr_wola_channelize": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                      c_vp, ctypes.c_int32, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "pysdr_psd_configure": (c_int, [c_vp, ctypes.c_int32, c_vp, ctypes.c_int32]),
    "pysdr_waterfall_push": (c_int, [c_vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_vp, ctypes.c_int32,
                                     ctypes.c_int32, ctypes.c_float, c_vp, c_vp, c_vp, c_vp]),
}

_lib = None


def load():
    """Load the shared library (built by pysdr_b200._build / __graft_entry__.build()); raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PysdrError("libpysdr_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
                         "there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)                      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().pysdr_last_error()
        raise PysdrError("libpysdr_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
