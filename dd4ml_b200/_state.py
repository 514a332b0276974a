"""Host mirror of the device state block (``csrc/state.h``).

The layout is read from the loaded library (``dd4ml_num_fields`` / ``dd4ml_field``)
so the struct definition in C is the single source of truth.
"""
from __future__ import annotations

import ctypes


class StateLayout:
    def __init__(self, lib, prefix="dd4ml_"):
        nf = getattr(lib, prefix + "num_fields")
        ff = getattr(lib, prefix + "field")
        sd = getattr(lib, prefix + "state_doubles")
        nf.restype = ctypes.c_int
        sd.restype = ctypes.c_int
        ff.restype = ctypes.c_int
        ff.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                       ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        self.size = int(sd())
        self.fields: dict[str, tuple[int, int]] = {}
        for i in range(int(nf())):
            name, off, cnt = ctypes.c_char_p(), ctypes.c_int(), ctypes.c_int()
            if ff(i, ctypes.byref(name), ctypes.byref(off), ctypes.byref(cnt)) != 0:
                raise RuntimeError("dd4ml_b200: state field table query failed")
            self.fields[name.value.decode()] = (off.value, cnt.value)
        self.report_end = self.fields["report_end"][0]
        self.maxs = self.fields["cS"][1]
        self.maxm = self.fields["order"][1]

    def off(self, name: str) -> int:
        return self.fields[name][0]

    def slice(self, name: str) -> slice:
        o, c = self.fields[name]
        return slice(o, o + c)


class Report:
    """Attribute view over a host copy (list of python floats) of the state block or its report prefix."""

    def __init__(self, layout: StateLayout, values):
        self._l = layout
        self._v = values

    def __getattr__(self, name):
        try:
            o, c = self._l.fields[name]
        except KeyError:
            raise AttributeError(name)
        if c == 1:
            return self._v[o]
        return self._v[o:o + c]
