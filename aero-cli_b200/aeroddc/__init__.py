"""aeroddc - thin ctypes binding of the C ABI in include/aeroddc.h (libaeroddc.so).

This package is plumbing only: it loads the CUDA library built from aero-cli_b200/csrc and exposes
the bank the way the reference drives its `vfo` objects (/root/reference/publish/vfo.h:16-40,
publisher.cpp:118-148,159-219,301-305). There is no CPU implementation behind it: if the shared
library is missing or no CUDA device is present, construction raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libaeroddc.so")

CU8, CS16, CF32 = 0, 1, 2
MODE_EXACT, MODE_FAST, MODE_TENSOR = 0, 1, 2
_NP_DTYPE = {CU8: np.uint8, CS16: np.int16, CF32: np.float32}


class AeroDdcError(RuntimeError):
    pass


class SegmentPlan(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("warmup", "boundary_warmup", "segment_len", "n_segments", "parts", "part_len", "vfo_groups", "ctas")]


class VfoDesc(ctypes.Structure):
    _fields_ = [
        ("mixer_freq", ctypes.c_double),
        ("decim_count", ctypes.c_int),
        ("late_decimate", ctypes.c_int),
        ("filter_bw", ctypes.c_int),
        ("gain", ctypes.c_float),
        ("demod_usb", ctypes.c_int),
        ("compress_style", ctypes.c_int),
        ("scale_comp", ctypes.c_int),
        ("topic", ctypes.c_char * 64),
        ("parent", ctypes.c_int),
    ]


# every symbol include/aeroddc.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "aeroddc_bank_create", "aeroddc_bank_add_vfo", "aeroddc_bank_finalize", "aeroddc_bank_process",
    "aeroddc_bank_submit", "aeroddc_bank_wait", "aeroddc_bank_host_slot", "aeroddc_bank_submit_device",
    "aeroddc_bank_output", "aeroddc_bank_topic", "aeroddc_bank_stage_d", "aeroddc_bank_num_vfos",
    "aeroddc_bank_last_timing", "aeroddc_bank_last_main_ms", "aeroddc_bank_device_bytes",
    "aeroddc_bank_destroy", "aeroddc_last_error", "aeroddc_measure_fp32_peak", "aeroddc_abi_version",
    "aeroddc_design_lowpass", "aeroddc_design_hilbert", "aeroddc_design_rotation", "aeroddc_bank_stopwatch", "aeroddc_bank_set_mode",
    "aeroddc_fleet_create", "aeroddc_fleet_add_vfo", "aeroddc_fleet_set_mode", "aeroddc_fleet_finalize", "aeroddc_fleet_host_slot",
    "aeroddc_fleet_submit", "aeroddc_fleet_wait", "aeroddc_fleet_process", "aeroddc_fleet_output", "aeroddc_fleet_num_devices",
    "aeroddc_fleet_device_of", "aeroddc_fleet_destroy", "aeroddc_bank_set_dc_correction", "aeroddc_fleet_set_dc_correction",
    "aeroddc_bank_reset", "aeroddc_fleet_reset", "aeroddc_fleet_exchange", "aeroddc_dev_upload_async", "aeroddc_bank_submit_device_sliced",
    "aeroddc_dev_alloc", "aeroddc_dev_free", "aeroddc_dev_upload", "aeroddc_ipc_export", "aeroddc_ipc_import", "aeroddc_ipc_close", "aeroddc_enable_peer", "aeroddc_plan_segments",
    "aeroddc_plan_tensor_stretches",
]

_lib = None


def lib():
    """Load libaeroddc.so; raises AeroDdcError if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AeroDdcError("%s not found: build it with `make -C aero-cli_b200/csrc` "
                               "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp, ci, cz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        L.aeroddc_bank_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci]
        L.aeroddc_bank_add_vfo.argtypes = [vp, ctypes.POINTER(VfoDesc)]
        L.aeroddc_bank_finalize.argtypes = [vp]
        L.aeroddc_bank_process.argtypes = [vp, vp, cz]
        L.aeroddc_bank_submit.argtypes = [vp, vp, cz]
        L.aeroddc_bank_wait.argtypes = [vp]
        L.aeroddc_bank_reset.argtypes = [vp]
        L.aeroddc_bank_submit_device_sliced.argtypes = [vp, ctypes.POINTER(vp), ci, cz, cz, ctypes.POINTER(vp), ci]
        L.aeroddc_fleet_reset.argtypes = [vp]
        L.aeroddc_bank_host_slot.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(cz)]
        L.aeroddc_bank_submit_device.argtypes = [vp, vp, cz, vp]
        L.aeroddc_bank_output.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(cz), ctypes.POINTER(ctypes.c_uint32)]
        L.aeroddc_bank_topic.argtypes = [vp, ci]
        L.aeroddc_bank_topic.restype = ctypes.c_char_p
        L.aeroddc_bank_stage_d.argtypes = [vp, ci, vp, cz]
        L.aeroddc_bank_num_vfos.argtypes = [vp]
        L.aeroddc_bank_last_timing.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ci)]
        L.aeroddc_bank_last_main_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
        L.aeroddc_bank_device_bytes.argtypes = [vp, ctypes.POINTER(cz)]
        L.aeroddc_bank_stopwatch.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_float)]
        L.aeroddc_bank_set_mode.argtypes = [vp, ci]
        L.aeroddc_plan_tensor_stretches.argtypes = [ci, ci, ci, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ci]
        L.aeroddc_bank_set_dc_correction.argtypes = [vp, ci]
        L.aeroddc_fleet_set_dc_correction.argtypes = [vp, ci]
        L.aeroddc_dev_alloc.argtypes = [ci, cz, ctypes.POINTER(vp)]
        L.aeroddc_dev_free.argtypes = [ci, vp]
        L.aeroddc_dev_upload.argtypes = [ci, vp, vp, cz]
        L.aeroddc_dev_upload_async.argtypes = [ci, vp, vp, cz, vp]
        L.aeroddc_ipc_export.argtypes = [ci, vp, ctypes.c_char_p]
        L.aeroddc_ipc_import.argtypes = [ci, ctypes.c_char_p, ctypes.POINTER(vp)]
        L.aeroddc_ipc_close.argtypes = [ci, vp]
        L.aeroddc_enable_peer.argtypes = [ci, ci]
        L.aeroddc_plan_segments.argtypes = [ci, ci, ci, ci, ctypes.c_double, ci, ctypes.POINTER(SegmentPlan)]
        L.aeroddc_fleet_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ctypes.POINTER(ci), ci]
        L.aeroddc_fleet_add_vfo.argtypes = [vp, ctypes.POINTER(VfoDesc)]
        L.aeroddc_fleet_set_mode.argtypes = [vp, ci]
        L.aeroddc_fleet_finalize.argtypes = [vp]
        L.aeroddc_fleet_host_slot.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(cz)]
        L.aeroddc_fleet_submit.argtypes = [vp, vp, cz]
        L.aeroddc_fleet_wait.argtypes = [vp]
        L.aeroddc_fleet_process.argtypes = [vp, vp, cz]
        L.aeroddc_fleet_output.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(cz), ctypes.POINTER(ctypes.c_uint32)]
        L.aeroddc_fleet_num_devices.argtypes = [vp]
        L.aeroddc_fleet_exchange.argtypes = [vp]
        L.aeroddc_fleet_device_of.argtypes = [vp, ci]
        L.aeroddc_fleet_destroy.argtypes = [vp]
        L.aeroddc_fleet_destroy.restype = None
        L.aeroddc_bank_destroy.argtypes = [vp]
        L.aeroddc_bank_destroy.restype = None
        L.aeroddc_last_error.restype = ctypes.c_char_p
        L.aeroddc_measure_fp32_peak.argtypes = [ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        cd, fp = ctypes.c_double, ctypes.POINTER(ctypes.c_float)
        L.aeroddc_design_lowpass.argtypes = [cd, cd, cd, cd, vp, ci]
        L.aeroddc_design_hilbert.argtypes = [ci, ci, vp, ci]
        L.aeroddc_design_rotation.argtypes = [cd, cd, fp, fp]
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise AeroDdcError("aeroddc error %d: %s" % (rc, lib().aeroddc_last_error().decode()))
    return rc


def design_lowpass(gain, fs, cutoff, transition):
    """float32 taps of the Hamming low-pass the bank would upload (host code, no GPU needed)."""
    n = _check(lib().aeroddc_design_lowpass(gain, fs, cutoff, transition, None, 0))
    out = np.empty(n, np.float32)
    _check(lib().aeroddc_design_lowpass(gain, fs, cutoff, transition, out.ctypes.data, n))
    return out


def design_hilbert(length, fs_param):
    out = np.empty(length, np.float32)
    _check(lib().aeroddc_design_hilbert(length, fs_param, out.ctypes.data, length))
    return out


def design_rotation(fs, freq):
    c, s = ctypes.c_float(), ctypes.c_float()
    _check(lib().aeroddc_design_rotation(fs, freq, ctypes.byref(c), ctypes.byref(s)))
    return c.value, s.value


def dev_alloc(device, nbytes):
    p = ctypes.c_void_p()
    _check(lib().aeroddc_dev_alloc(device, nbytes, ctypes.byref(p)))
    return p.value


def dev_free(device, ptr):
    _check(lib().aeroddc_dev_free(device, ptr))


def dev_upload(device, ptr, array):
    a = np.ascontiguousarray(array)
    _check(lib().aeroddc_dev_upload(device, ptr, a.ctypes.data, a.nbytes))


def dev_upload_async(device, ptr, pinned_array, stream):
    """cudaMemcpyAsync H2D of a contiguous view of pinned host memory on the given cudaStream_t."""
    _check(lib().aeroddc_dev_upload_async(device, ptr, pinned_array.ctypes.data, pinned_array.nbytes, stream))


def ipc_export(device, ptr):
    """64-byte CUDA IPC handle of a block allocated with dev_alloc in this process."""
    buf = ctypes.create_string_buffer(64)
    _check(lib().aeroddc_ipc_export(device, ptr, buf))
    return buf.raw


def ipc_import(device, handle):
    """Device address, valid on `device` of THIS process, of a block another process exported."""
    p = ctypes.c_void_p()
    _check(lib().aeroddc_ipc_import(device, handle, ctypes.byref(p)))
    return p.value


def enable_peer(device, peer):
    _check(lib().aeroddc_enable_peer(device, peer))


def ipc_close(device, ptr):
    _check(lib().aeroddc_ipc_close(device, ptr))


def plan_segments(block_len, decim_count, n_vfos, n_sm=148, waves=1.0, parts=0):
    """The bank's segmentation of one VFO group (host arithmetic only)."""
    p = SegmentPlan()
    _check(lib().aeroddc_plan_segments(block_len, decim_count, n_vfos, n_sm, waves, parts, ctypes.byref(p)))
    return {n: getattr(p, n) for n, _ in SegmentPlan._fields_}


def plan_tensor_stretches(n_tiles, n_mid, n_sm=148):
    """Host-side plan of the tensor mode: (number of CTAs, [(cta, tile, first output, end output)])."""
    L = lib()
    P = _check(L.aeroddc_plan_tensor_stretches(n_tiles, n_mid, n_sm, None, None, 0))
    first = (ctypes.c_int * (P + 1))()
    cap = P + n_tiles + 1
    flat = (ctypes.c_int * (3 * cap))()
    n = _check(L.aeroddc_plan_tensor_stretches(n_tiles, n_mid, n_sm, first, flat, cap))
    out = []
    for c in range(P):
        for i in range(first[c], first[c + 1]):
            out.append((c, flat[3 * i], flat[3 * i + 1], flat[3 * i + 2]))
    assert len(out) == n
    return P, out


def measure_fp32_peak(device=0):
    """(TFLOP/s counting FMA as 2, SM clock in MHz) from the library's FFMA issue-rate probe."""
    t, c = ctypes.c_double(), ctypes.c_double()
    _check(lib().aeroddc_measure_fp32_peak(device, ctypes.byref(t), ctypes.byref(c)))
    return t.value, c.value


class Bank:
    """All VFOs of one raw-IQ stream. Mirrors create -> add_vfo* -> finalize -> process* of the ABI."""

    def __init__(self, sample_rate, block_len, in_format=CF32, device=0):
        self._L = lib()
        self._h = ctypes.c_void_p()
        self.sample_rate, self.block_len, self.in_format, self.device = sample_rate, block_len, in_format, device
        _check(self._L.aeroddc_bank_create(ctypes.byref(self._h), sample_rate, block_len, in_format, device))

    def add_vfo(self, mixer_freq, decim_count, late_decimate=0, filter_bw=0, gain=0.01, demod_usb=1,
                compress_style=1, scale_comp=1, topic="", parent=-1):
        d = VfoDesc(float(mixer_freq), int(decim_count), int(late_decimate), int(filter_bw), float(gain),
                    int(demod_usb), int(compress_style), int(scale_comp), topic.encode()[:63], int(parent))
        return _check(self._L.aeroddc_bank_add_vfo(self._h, ctypes.byref(d)))

    def set_mode(self, mode):
        _check(self._L.aeroddc_bank_set_mode(self._h, mode))

    def set_dc_correction(self, enable=True):
        _check(self._L.aeroddc_bank_set_dc_correction(self._h, 1 if enable else 0))

    def finalize(self):
        _check(self._L.aeroddc_bank_finalize(self._h))

    @property
    def num_vfos(self):
        return self._L.aeroddc_bank_num_vfos(self._h)

    def _ptr(self, block):
        if isinstance(block, np.ndarray):
            if block.dtype != _NP_DTYPE[self.in_format] or not block.flags.c_contiguous:
                raise AeroDdcError("block must be a C-contiguous %s array" % _NP_DTYPE[self.in_format].__name__)
            if block.size != 2 * self.block_len:
                raise AeroDdcError("block has %d values, expected %d" % (block.size, 2 * self.block_len))
            return block.ctypes.data
        return int(block)

    def process(self, block):
        _check(self._L.aeroddc_bank_process(self._h, self._ptr(block), self.block_len))

    def submit(self, block):
        _check(self._L.aeroddc_bank_submit(self._h, self._ptr(block), self.block_len))

    def submit_device(self, dev_ptr, ready_event=None):
        _check(self._L.aeroddc_bank_submit_device(self._h, int(dev_ptr), self.block_len, ready_event))

    def submit_device_sliced(self, slice_ptrs, slice_len, ready_events=()):
        """The block lies in len(slice_ptrs) device buffers of slice_len complex samples each (possibly on different GPUs)."""
        sl = (ctypes.c_void_p * len(slice_ptrs))(*[int(p) for p in slice_ptrs])
        ev = (ctypes.c_void_p * max(len(ready_events), 1))(*[int(e) if e else None for e in ready_events])
        _check(self._L.aeroddc_bank_submit_device_sliced(self._h, sl, len(slice_ptrs), int(slice_len), self.block_len, ev, len(ready_events)))

    def wait(self):
        _check(self._L.aeroddc_bank_wait(self._h))

    def reset(self):
        """Rewind to stream position 0 (fresh filter/oscillator state); nothing may be in flight."""
        _check(self._L.aeroddc_bank_reset(self._h))

    def host_slot(self, slot):
        """numpy view of pinned staging slot 0/1 (dtype of the input format, 2*block_len values)."""
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._L.aeroddc_bank_host_slot(self._h, slot, ctypes.byref(p), ctypes.byref(n)))
        dt = np.dtype(_NP_DTYPE[self.in_format])
        buf = (ctypes.c_ubyte * n.value).from_address(p.value)
        return np.frombuffer(buf, dtype=dt)

    def output(self, vfo):
        """(payload bytes, output rate) of the most recently completed block."""
        p, n, r = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_uint32()
        _check(self._L.aeroddc_bank_output(self._h, vfo, ctypes.byref(p), ctypes.byref(n), ctypes.byref(r)))
        return ctypes.string_at(p.value, n.value), r.value

    def topic(self, vfo):
        return self._L.aeroddc_bank_topic(self._h, vfo).decode()

    def stage_d(self, vfo, n_complex):
        out = np.empty(2 * n_complex, np.float32)
        n = _check(self._L.aeroddc_bank_stage_d(self._h, vfo, out.ctypes.data, n_complex))
        return out[: 2 * min(n, n_complex)]

    def last_timing(self):
        ms, n = ctypes.c_float(), ctypes.c_int()
        _check(self._L.aeroddc_bank_last_timing(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def last_main_ms(self):
        ms = ctypes.c_float()
        _check(self._L.aeroddc_bank_last_main_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def stopwatch_start(self, host_inputs=False):
        _check(self._L.aeroddc_bank_stopwatch(self._h, 1 if host_inputs else 0, None))

    def stopwatch_stop(self):
        ms = ctypes.c_float()
        _check(self._L.aeroddc_bank_stopwatch(self._h, 2, ctypes.byref(ms)))
        return ms.value

    def device_bytes(self):
        n = ctypes.c_size_t()
        _check(self._L.aeroddc_bank_device_bytes(self._h, ctypes.byref(n)))
        return n.value

    def close(self):
        if self._h:
            self._L.aeroddc_bank_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Fleet:
    """One bank per GPU of the node behind one handle (include/aeroddc.h, aeroddc_fleet_*): VFOs sharded over
    the devices, raw blocks uploaded once and broadcast with NCCL. Same call order as Bank."""

    def __init__(self, sample_rate, block_len, in_format=CF32, devices=(0,)):
        self._L = lib()
        self._h = ctypes.c_void_p()
        self.block_len, self.in_format = block_len, in_format
        devs = (ctypes.c_int * len(devices))(*devices)
        _check(self._L.aeroddc_fleet_create(ctypes.byref(self._h), sample_rate, block_len, in_format, devs, len(devices)))

    def add_vfo(self, mixer_freq, decim_count, late_decimate=0, filter_bw=0, gain=0.01, demod_usb=1,
                compress_style=1, scale_comp=1, topic="", parent=-1):
        d = VfoDesc(float(mixer_freq), int(decim_count), int(late_decimate), int(filter_bw), float(gain),
                    int(demod_usb), int(compress_style), int(scale_comp), topic.encode()[:63], int(parent))
        return _check(self._L.aeroddc_fleet_add_vfo(self._h, ctypes.byref(d)))

    def set_mode(self, mode):
        _check(self._L.aeroddc_fleet_set_mode(self._h, mode))

    def finalize(self):
        _check(self._L.aeroddc_fleet_finalize(self._h))

    def _ptr(self, block):
        if isinstance(block, np.ndarray):
            if block.dtype != _NP_DTYPE[self.in_format] or not block.flags.c_contiguous or block.size != 2 * self.block_len:
                raise AeroDdcError("block must be a C-contiguous %s array of %d values" % (_NP_DTYPE[self.in_format].__name__, 2 * self.block_len))
            return block.ctypes.data
        return int(block)

    def process(self, block):
        _check(self._L.aeroddc_fleet_process(self._h, self._ptr(block), self.block_len))

    def submit(self, block):
        _check(self._L.aeroddc_fleet_submit(self._h, self._ptr(block), self.block_len))

    def wait(self):
        _check(self._L.aeroddc_fleet_wait(self._h))

    def reset(self):
        _check(self._L.aeroddc_fleet_reset(self._h))

    def host_slot(self, slot):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._L.aeroddc_fleet_host_slot(self._h, slot, ctypes.byref(p), ctypes.byref(n)))
        return np.frombuffer((ctypes.c_ubyte * n.value).from_address(p.value), dtype=np.dtype(_NP_DTYPE[self.in_format]))

    def output(self, vfo):
        p, n, r = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_uint32()
        _check(self._L.aeroddc_fleet_output(self._h, vfo, ctypes.byref(p), ctypes.byref(n), ctypes.byref(r)))
        return ctypes.string_at(p.value, n.value), r.value

    @property
    def num_devices(self):
        return self._L.aeroddc_fleet_num_devices(self._h)

    def device_of(self, vfo):
        return self._L.aeroddc_fleet_device_of(self._h, vfo)

    @property
    def exchange(self):
        return {0: "single", 1: "peer", 2: "nccl"}[self._L.aeroddc_fleet_exchange(self._h)]

    def close(self):
        if self._h:
            self._L.aeroddc_fleet_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
