"""ctypes binding of include/b200gnss.h.  Loads the in-tree libb200gnss.so; fails loudly if absent."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200gnss.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m gnss_sdr_b200.build` "
        "(there is no CPU fallback for the B200 hot path)")

lib = C.CDLL(LIB_PATH)

B200_MAX_TAPS = 8
ERRORS = {0: "OK", -1: "ERR_ARG", -2: "ERR_CUDA", -3: "ERR_NOMEM", -4: "ERR_STATE", -5: "ERR_RANGE", -6: "ERR_NODEV"}


class B200Error(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib.b200_last_error()
        super().__init__(f"{where}: {ERRORS.get(code, code)}: {msg.decode() if msg else ''}")


class TrkItem(C.Structure):
    _fields_ = [("channel", C.c_int32), ("n", C.c_int32), ("sample_index", C.c_uint64),
                ("rem_carrier_phase_rad", C.c_float), ("phase_step_rad", C.c_float), ("phase_rate_step_rad", C.c_float),
                ("rem_code_phase_chips", C.c_float), ("code_phase_step_chips", C.c_float),
                ("code_phase_rate_step_chips", C.c_float)]


TRK_ITEM_DTYPE = np.dtype([("channel", "<i4"), ("n", "<i4"), ("sample_index", "<u8"),
                           ("rem_carrier_phase_rad", "<f4"), ("phase_step_rad", "<f4"), ("phase_rate_step_rad", "<f4"),
                           ("rem_code_phase_chips", "<f4"), ("code_phase_step_chips", "<f4"),
                           ("code_phase_rate_step_chips", "<f4")])
assert TRK_ITEM_DTYPE.itemsize == 40 and C.sizeof(TrkItem) == 40

lib.b200_last_error.restype = C.c_char_p
_vp = C.c_void_p

_SIGS = {
    "b200_version": ([], C.c_int),
    "b200_device_count": ([C.POINTER(C.c_int)], C.c_int),
    "b200_engine_create": ([C.POINTER(_vp), C.c_int, _vp], C.c_int),
    "b200_engine_destroy": ([_vp], C.c_int),
    "b200_engine_sync": ([_vp], C.c_int),
    "b200_engine_timer_start": ([_vp], C.c_int),
    "b200_engine_timer_stop_ms": ([_vp, C.POINTER(C.c_float)], C.c_int),
    "b200_engine_launch_count": ([_vp, C.POINTER(C.c_uint64)], C.c_int),
    "b200_iq_create": ([_vp, C.c_int, C.c_uint64], C.c_int),
    "b200_iq_push": ([_vp, C.c_int, _vp, C.c_uint64, C.POINTER(C.c_uint64)], C.c_int),
    "b200_iq_attach_dev": ([_vp, C.c_int, _vp, C.c_uint64, C.c_uint64], C.c_int),
    "b200_trk_create": ([_vp, C.POINTER(_vp), C.c_int, C.c_int], C.c_int),
    "b200_trk_set_high_dynamics_resampler": ([_vp, C.c_int], C.c_int),
    "b200_trk_set_local_code_and_taps": ([_vp, C.c_int, _vp, _vp], C.c_int),
    "b200_trk_correlate": ([_vp, _vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _vp], C.c_int),
    "b200_trk_destroy": ([_vp], C.c_int),
    "b200_trk_channel_create": ([_vp, C.c_int, C.c_int, C.POINTER(C.c_int)], C.c_int),
    "b200_trk_channel_set_code": ([_vp, C.c_int, C.c_int, _vp, _vp, C.c_int], C.c_int),
    "b200_trk_batch": ([_vp, _vp, C.c_int, _vp, C.c_int], C.c_int),
    "b200_trk_batch_dev": ([_vp, _vp, C.c_int, _vp, C.c_int, C.c_int], C.c_int),
}
for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = _res


def exported_symbols():
    """Names declared in include/b200gnss.h that this binding knows about."""
    return list(_SIGS) + ["b200_last_error"]


def _chk(rc: int, where: str):
    if rc != 0:
        raise B200Error(rc, where)


def device_count() -> int:
    n = C.c_int(0)
    lib.b200_device_count(C.byref(n))
    return n.value


class Engine:
    """b200_engine handle."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = _vp()
        _chk(lib.b200_engine_create(C.byref(h), device, _vp(stream) if stream else None), "b200_engine_create")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            lib.b200_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _chk(lib.b200_engine_sync(self.h), "b200_engine_sync")

    def timer_start(self):
        _chk(lib.b200_engine_timer_start(self.h), "timer_start")

    def timer_stop_ms(self) -> float:
        ms = C.c_float(0)
        _chk(lib.b200_engine_timer_stop_ms(self.h, C.byref(ms)), "timer_stop")
        return ms.value

    def launch_count(self) -> int:
        n = C.c_uint64(0)
        _chk(lib.b200_engine_launch_count(self.h, C.byref(n)), "launch_count")
        return n.value

    # bands
    def iq_create(self, band: int, capacity: int):
        _chk(lib.b200_iq_create(self.h, band, capacity), "b200_iq_create")

    def iq_push(self, band: int, iq: np.ndarray) -> int:
        iq = np.ascontiguousarray(iq, np.complex64)
        first = C.c_uint64(0)
        _chk(lib.b200_iq_push(self.h, band, iq.ctypes.data, iq.size, C.byref(first)), "b200_iq_push")
        return first.value

    def iq_push_ptr(self, band: int, host_ptr: int, n: int) -> int:
        first = C.c_uint64(0)
        _chk(lib.b200_iq_push(self.h, band, host_ptr, n, C.byref(first)), "b200_iq_push")
        return first.value

    def iq_attach_dev(self, band: int, dev_ptr: int, n_samples: int, first_index: int = 0):
        _chk(lib.b200_iq_attach_dev(self.h, band, dev_ptr, n_samples, first_index), "b200_iq_attach_dev")

    # channels
    def channel_create(self, band: int, taps: int) -> int:
        cid = C.c_int(-1)
        _chk(lib.b200_trk_channel_create(self.h, band, taps, C.byref(cid)), "b200_trk_channel_create")
        return cid.value

    def channel_set_code(self, cid: int, code, shifts, high_dyn: bool = False):
        code = np.ascontiguousarray(code, np.float32)
        shifts = np.ascontiguousarray(shifts, np.float32)
        _chk(lib.b200_trk_channel_set_code(self.h, cid, code.size, code.ctypes.data, shifts.ctypes.data, int(high_dyn)),
             "b200_trk_channel_set_code")

    def trk_batch(self, items: np.ndarray, out_stride: int) -> np.ndarray:
        items = np.ascontiguousarray(items, TRK_ITEM_DTYPE)
        out = np.zeros((items.size, out_stride), np.complex64)
        _chk(lib.b200_trk_batch(self.h, items.ctypes.data, items.size, out.ctypes.data, out_stride), "b200_trk_batch")
        return out

    def trk_batch_dev(self, items_dev_ptr: int, n_items: int, out_dev_ptr: int, out_stride: int, slices: int = 1):
        _chk(lib.b200_trk_batch_dev(self.h, items_dev_ptr, n_items, out_dev_ptr, out_stride, slices), "b200_trk_batch_dev")


class Multicorrelator:
    """b200_trk handle; same call sequence as the reference's Cpu_Multicorrelator_Real_Codes."""

    def __init__(self, engine: Engine, max_signal_length_samples: int, n_correlators: int):
        h = _vp()
        _chk(lib.b200_trk_create(engine.h, C.byref(h), max_signal_length_samples, n_correlators), "b200_trk_create")
        self.h = h
        self.engine = engine
        self.taps = n_correlators

    def set_high_dynamics_resampler(self, flag: bool):
        _chk(lib.b200_trk_set_high_dynamics_resampler(self.h, int(flag)), "set_high_dynamics_resampler")

    def set_local_code_and_taps(self, code, shifts):
        code = np.ascontiguousarray(code, np.float32)
        shifts = np.ascontiguousarray(shifts, np.float32)
        assert shifts.size == self.taps
        _chk(lib.b200_trk_set_local_code_and_taps(self.h, code.size, code.ctypes.data, shifts.ctypes.data),
             "set_local_code_and_taps")

    def Carrier_wipeoff_multicorrelator_resampler(self, sig_in, rem_carrier_phase_in_rad, phase_step_rad,
                                                  phase_rate_step_rad, rem_code_phase_chips, code_phase_step_chips,
                                                  code_phase_rate_step_chips, signal_length_samples=None) -> np.ndarray:
        sig_in = np.ascontiguousarray(sig_in, np.complex64)
        n = sig_in.size if signal_length_samples is None else signal_length_samples
        out = np.zeros(self.taps, np.complex64)
        _chk(lib.b200_trk_correlate(self.h, sig_in.ctypes.data, rem_carrier_phase_in_rad, phase_step_rad,
                                    phase_rate_step_rad, rem_code_phase_chips, code_phase_step_chips,
                                    code_phase_rate_step_chips, n, out.ctypes.data), "b200_trk_correlate")
        return out

    def free(self):
        if self.h:
            lib.b200_trk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
