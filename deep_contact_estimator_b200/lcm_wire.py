"""Minimal LCM wire format for the output side of the pipeline.

`save2lcm` (/root/reference/src/inference_one_seq.py:91-133) writes three events per time step
to an `lcm.EventLog`.  liblcm is not installed in this image, so this module provides the few
pieces that function needs, written from the IDL (/root/reference/lcm_types/*.lcm):

  * encoders / decoders for `contact_t`, `leg_control_data_lcmt`, `microstrain_lcmt`: 8-byte big-endian type
    fingerprint, then the fields in declaration order, big-endian (the LCM marshalling rules);
  * `EventLog`: the LCM log-file container (sync word 0xEDA1DA01, event number, timestamp in
    microseconds, channel length, payload length, channel, payload — all big-endian).

When the real `lcm` package is importable, callers should prefer it; the byte streams are identical
(tests/test_host_cpu.py checks the encoders against golden bytes produced by the reference's
generated Python types).
"""
from __future__ import annotations

import struct
from typing import Sequence

_MASK = 0xFFFFFFFFFFFFFFFF


def _fingerprint(base_hash: int) -> bytes:
    # LCM: the packed fingerprint of a struct without nested types is rotl(base_hash, 1)
    h = base_hash & _MASK
    h = (((h << 1) & _MASK) + (h >> 63)) & _MASK
    return struct.pack(">Q", h)


# base hashes of the three IDL structs (lcm-gen output, lcm_types/python/*.py)
_FP_CONTACT = _fingerprint(0x12E312DF2F1D46F6)
_FP_LEG = _fingerprint(0xA7D2775A407DECA7)
_FP_IMU = _fingerprint(0x710A98F509C97D55)


def encode_contact(num_legs: int, timestamp: float, contact: Sequence[int]) -> bytes:
    """contact_t { int8_t num_legs; double timestamp; int8_t contact[num_legs]; }"""
    return _FP_CONTACT + struct.pack(">bd", num_legs, timestamp) + struct.pack(">%db" % num_legs, *[int(c) for c in contact[:num_legs]])


def decode_contact(data: bytes):
    if data[:8] != _FP_CONTACT:
        raise ValueError("not a contact_t")
    n, ts = struct.unpack(">bd", data[8:17])
    return n, ts, list(struct.unpack(">%db" % n, data[17:17 + n]))


def encode_leg_control_data(q, qd, p, v, tau_est) -> bytes:
    """leg_control_data_lcmt { float q[12]; float qd[12]; float p[12]; float v[12]; float tau_est[12]; }"""
    out = _FP_LEG
    for arr in (q, qd, p, v, tau_est):
        out += struct.pack(">12f", *[float(a) for a in arr[:12]])
    return out


def encode_microstrain(quat, rpy, omega, acc, good_packets: int = 0, bad_packets: int = 0) -> bytes:
    """microstrain_lcmt { float quat[4]; float rpy[3]; float omega[3]; float acc[3]; int64_t good_packets, bad_packets; }"""
    return (_FP_IMU + struct.pack(">4f", *[float(a) for a in quat[:4]]) + struct.pack(">3f", *[float(a) for a in rpy[:3]])
            + struct.pack(">3f", *[float(a) for a in omega[:3]]) + struct.pack(">3f", *[float(a) for a in acc[:3]])
            + struct.pack(">qq", int(good_packets), int(bad_packets)))


def decode_leg_control_data(data: bytes):
    """-> (q, qd, p, v, tau_est), five tuples of 12 floats"""
    if data[:8] != _FP_LEG or len(data) < 8 + 5 * 48:
        raise ValueError("not a leg_control_data_lcmt")
    return tuple(struct.unpack(">12f", data[8 + 48 * i: 8 + 48 * (i + 1)]) for i in range(5))


def decode_microstrain(data: bytes):
    """-> (quat[4], rpy[3], omega[3], acc[3], good_packets, bad_packets)"""
    if data[:8] != _FP_IMU or len(data) < 8 + 52 + 16:
        raise ValueError("not a microstrain_lcmt")
    quat = struct.unpack(">4f", data[8:24])
    rpy, omega, acc = (struct.unpack(">3f", data[24 + 12 * i: 36 + 12 * i]) for i in range(3))
    good, bad = struct.unpack(">qq", data[60:76])
    return quat, rpy, omega, acc, good, bad


class EventLog:
    """Write-only LCM event log (`lcm.EventLog(path, mode='w', overwrite=True)` + `write_event`)."""
    SYNC = 0xEDA1DA01

    def __init__(self, path: str, mode: str = "w", overwrite: bool = True):
        if mode != "w":
            raise ValueError("this shim only writes logs")
        self._f = open(path, "wb" if overwrite else "xb")
        self._n = 0

    def write_event(self, utime: int, channel: str, data: bytes):
        ch = channel.encode("utf-8")
        self._f.write(struct.pack(">IqqII", self.SYNC, self._n, int(utime), len(ch), len(data)))
        self._f.write(ch)
        self._f.write(data)
        self._n += 1

    def close(self):
        if self._f:
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_events(path: str):
    """Iterate `(event number, utime, channel, payload)` of a log (used by the tests)."""
    with open(path, "rb") as f:
        while True:
            head = f.read(28)
            if len(head) < 28:
                return
            sync, num, utime, clen, dlen = struct.unpack(">IqqII", head)
            if sync != EventLog.SYNC:
                raise ValueError("bad sync word")
            yield num, utime, f.read(clen).decode("utf-8"), f.read(dlen)
