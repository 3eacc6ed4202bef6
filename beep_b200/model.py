"""BeepModelDesc on the Python side.

``include/beep_model_fields.def`` is the single source of truth for the flat
model description (the reference's ``Details`` + ``Data`` + ``Model`` after
construction, reference ``src/main.cc:125-151``).  This module parses that file
to build the ctypes mirror of ``struct BeepModelDesc`` and keeps the arrays
alive next to it.
"""
import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DEF = os.path.join(os.path.dirname(_HERE), "include", "beep_model_fields.def")


def _parse_fields():
    fields = []
    pat = re.compile(r"^\s*BEEP_(I32|F64|AI32|AF64)\((\w+)\)")
    with open(_DEF) as f:
        for line in f:
            mm = pat.match(line)
            if mm:
                fields.append((mm.group(1), mm.group(2)))
    return fields


FIELDS = _parse_fields()


def _ctypes_fields():
    out = []
    for kind, name in FIELDS:
        if kind == "I32":
            out.append((name, ctypes.c_int32))
        elif kind == "F64":
            out.append((name, ctypes.c_double))
        elif kind == "AI32":
            out.append((name, ctypes.POINTER(ctypes.c_int32)))
            out.append(("n_" + name, ctypes.c_int64))
        else:
            out.append((name, ctypes.POINTER(ctypes.c_double)))
            out.append(("n_" + name, ctypes.c_int64))
    return out


# arrays a description may lack (fixtures written before the feature existed): empty = the analysis does not use it
OPTIONAL = {"dcc_info", "dcc_area", "dcc_group", "dcc_frac", "table_type", "graph_step", "graph_type", "graph_table", "graph_factor",
            "graph_factor_spline", "mod_transmean", "mod_beta", "mod_efoi", "mod_spline_mult", "mod_spline_set", "graph_area_off", "graph_area", "graph_dp_off", "graph_dp", "graph_point_off", "graph_point_obs"}


class BeepModelDescStruct(ctypes.Structure):
    _fields_ = _ctypes_fields()


class ModelDesc:
    """Flat model description: a dict of numpy arrays plus the C struct view."""

    def __init__(self, arrays):
        self.arrays = {}
        self.c = BeepModelDescStruct()
        for kind, name in FIELDS:
            if name not in arrays:
                if name not in OPTIONAL:
                    raise KeyError(f"model description lacks field '{name}'")
                arrays = dict(arrays)
                arrays[name] = np.zeros(0)
            v = arrays[name]
            if kind == "I32":
                val = int(np.asarray(v).reshape(-1)[0])
                self.arrays[name] = val
                setattr(self.c, name, val)
            elif kind == "F64":
                val = float(np.asarray(v).reshape(-1)[0])
                self.arrays[name] = val
                setattr(self.c, name, val)
            else:
                dt = np.int32 if kind == "AI32" else np.float64
                a = np.ascontiguousarray(np.asarray(v, dtype=dt))
                self.arrays[name] = a
                ct = ctypes.c_int32 if kind == "AI32" else ctypes.c_double
                setattr(self.c, name, a.ctypes.data_as(ctypes.POINTER(ct)))
                setattr(self.c, "n_" + name, a.size)

    def __getattr__(self, name):
        arrays = self.__dict__.get("arrays", {})
        if name in arrays:
            return arrays[name]
        raise AttributeError(name)

    @property
    def ptr(self):
        return ctypes.byref(self.c)

    @classmethod
    def from_prefixed(cls, d, prefix="desc."):
        """Build from a dict whose model fields carry a prefix (BPK / npz fixtures)."""
        return cls({k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)})

    def to_prefixed(self, prefix="desc."):
        out = {}
        for kind, name in FIELDS:
            v = self.arrays[name]
            if kind == "I32":
                out[prefix + name] = np.array([v], dtype=np.int32)
            elif kind == "F64":
                out[prefix + name] = np.array([v], dtype=np.float64)
            else:
                out[prefix + name] = v
        return out

    # shapes of the reference-ordered arrays
    @property
    def traj_shape(self):
        return (self.T, self.A, self.J, self.D)

    @property
    def pop_shape(self):
        return (self.T, self.A, self.C, self.D)
