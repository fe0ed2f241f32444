"""TEST INFRASTRUCTURE ONLY -- ctypes front-end for the two CPU checkers.

  kind="port": oracle/libgoicp_oracle.so   (plain-C restatement, goicp_oracle.c)
  kind="ref" : oracle/_ref/libgoicp_ref.so (the reference's own sources compiled in place + ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libgoicp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgoicp_ref.so")


class Params(C.Structure):
    """config.txt keys (jly_main.cpp:231-270); identical layout to include/goicp_b200.h:goicp_params."""
    _fields_ = [("MSEThresh", C.c_float),
                ("rotMinX", C.c_float), ("rotMinY", C.c_float), ("rotMinZ", C.c_float), ("rotWidth", C.c_float),
                ("transMinX", C.c_float), ("transMinY", C.c_float), ("transMinZ", C.c_float), ("transWidth", C.c_float),
                ("trimFraction", C.c_float),
                ("regularization", C.c_float), ("regularizationNeighbors", C.c_float), ("regularizationFPFH", C.c_float),
                ("cfpfh", C.c_int), ("norm", C.c_int), ("ponderation", C.c_int),
                ("distTransSize", C.c_int),
                ("distTransExpandFactor", C.c_double)]


class Result(C.Structure):
    _fields_ = [("R", C.c_double * 9), ("t", C.c_double * 3), ("optError", C.c_float), ("optComp", C.c_int),
                ("counters", C.c_longlong * 8), ("seconds_dt", C.c_double), ("seconds_register", C.c_double)]


def shipped_config(**kw):
    """The shipped config.txt (config.txt:4-53)."""
    p = Params(0.01, -3.1416, -3.1416, -3.1416, 6.2832, -0.5, -0.5, -0.5, 1.0, 0.0,
               0.0005, 0.0, 0.0, 0, 2, 1, 20, 2.0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def upstream_config(**kw):
    """The upstream Go-ICP demo config (READMEGo-ICP.md:33-41): MSE 1e-3, DT 300^3, no fork terms."""
    p = Params(0.001, -3.1416, -3.1416, -3.1416, 6.2832, -0.5, -0.5, -0.5, 1.0, 0.0,
               0.0, 0.0, 0.0, 0, 2, 0, 300, 2.0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def build(kind):
    if kind == "port":
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    else:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def available(kind):
    return os.path.exists(PORT_SO if kind == "port" else REF_SO)


_libs = {}


def _lib(kind):
    if kind in _libs:
        return _libs[kind]
    path = PORT_SO if kind == "port" else REF_SO
    if not os.path.exists(path):
        if kind == "port" or os.path.isdir("/root/reference"):
            build(kind)
        else:
            raise FileNotFoundError(path)
    lib = C.CDLL(path)
    pre = "orc_" if kind == "port" else "ref_"
    f = lambda name: getattr(lib, pre + name)
    vp, fp, ip, dp = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)
    f("create").restype = vp
    f("create").argtypes = [fp, ip, fp, C.c_int, fp, ip, fp, C.c_int, C.POINTER(Params)]
    f("destroy").argtypes = [vp]
    f("set_nd").argtypes = [vp, C.c_int]
    f("build_dt").restype = C.c_double
    f("build_dt").argtypes = [vp]
    f("dt_info").argtypes = [vp, dp]
    f("dt_download").argtypes = [vp, fp, C.POINTER(C.c_short), ip, ip]
    f("dt_distance").argtypes = [vp, dp, C.c_int, fp, ip]
    f("initialize").argtypes = [vp]
    f("get_weights").argtypes = [vp, fp]
    f("get_maxrotdis").argtypes = [vp, fp]
    f("get_ssethresh").restype = C.c_float
    f("get_ssethresh").argtypes = [vp]
    f("get_inliernum").restype = C.c_int
    f("get_inliernum").argtypes = [vp]
    f("inner_bnb").restype = C.c_float
    f("inner_bnb").argtypes = [vp, fp, C.c_int, C.c_float, fp]
    f("icp").restype = C.c_float
    f("icp").argtypes = [vp, dp, dp, ip]
    f("register").argtypes = [vp, C.c_int, C.POINTER(Result), C.c_char_p, C.c_int]
    f("normalize").restype = C.c_double
    f("normalize").argtypes = [dp, C.c_int, dp]
    f("scale").argtypes = [dp, C.c_int, C.c_double]
    if kind == "port":
        lib.orc_eval_leaf.argtypes = [vp, fp, C.c_int, fp, C.c_int, fp, fp, ip, ip]
        lib.orc_dt_upload.argtypes = [vp, fp, ip]
        lib.orc_eval_inclusion.argtypes = [vp, fp, C.c_int, fp, C.c_int, C.POINTER(C.c_ubyte), fp]
        lib.orc_round6.argtypes = [dp, C.c_int, fp]
        lib.orc_rescale_translation.argtypes = [C.c_double, dp, dp, dp, dp, dp]
        lib.orc_apply_rigid.argtypes = [dp, C.c_int, dp, dp, dp]
        lib.orc_rmsd.restype = C.c_float
        lib.orc_rmsd.argtypes = [dp, dp, C.c_int]
    else:
        lib.ref_eval_inclusion.argtypes = [vp, fp, C.c_int, fp, C.c_int, fp, fp]
        lib.ref_read_mol2.restype = C.c_int
        lib.ref_read_mol2.argtypes = [C.c_char_p, dp, ip, C.c_int]
        lib.ref_write_xyz.restype = C.c_int
        lib.ref_write_xyz.argtypes = [C.c_char_p, dp, ip, C.c_int]
        lib.ref_rescale.restype = C.c_int
        lib.ref_rescale.argtypes = [C.c_char_p, C.c_double, dp, dp, dp, dp, C.c_double, C.c_double]
        lib.ref_rmsd.restype = C.c_float
        lib.ref_rmsd.argtypes = [C.c_char_p, C.c_char_p]
        lib.ref_apply_protein.restype = C.c_int
        lib.ref_apply_protein.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    _libs[kind] = (lib, f)
    return _libs[kind]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """One registration problem (model = target cloud, data = source cloud), reference semantics."""

    def __init__(self, kind, model_xyz, data_xyz, params, model_c=None, data_c=None, model_fpfh=None, data_fpfh=None):
        self.kind = kind
        self.lib, self.f = _lib(kind)
        self.m = _f32(model_xyz).reshape(-1, 3)
        self.d = _f32(data_xyz).reshape(-1, 3)
        self.Nm, self.NdAll = len(self.m), len(self.d)
        self.Nd = self.NdAll
        self.mc = None if model_c is None else np.ascontiguousarray(model_c, dtype=np.int32)
        self.dc = None if data_c is None else np.ascontiguousarray(data_c, dtype=np.int32)
        self.mf = None if model_fpfh is None else _f32(model_fpfh).reshape(-1, 41)
        self.df = None if data_fpfh is None else _f32(data_fpfh).reshape(-1, 41)
        self.params = params
        self.S = params.distTransSize
        self.h = self.f("create")(_ptr(self.m, C.c_float), _ptr(self.mc, C.c_int), _ptr(self.mf, C.c_float), self.Nm,
                                  _ptr(self.d, C.c_float), _ptr(self.dc, C.c_int), _ptr(self.df, C.c_float), self.NdAll,
                                  C.byref(params))

    def close(self):
        if self.h:
            self.f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_nd(self, nd):
        self.Nd = nd
        self.f("set_nd")(self.h, nd)

    def build_dt(self):
        return self.f("build_dt")(self.h)

    def dt_info(self):
        o = np.zeros(8)
        self.f("dt_info")(self.h, _ptr(o, C.c_double))
        return dict(xMin=o[0], xMax=o[1], yMin=o[2], yMax=o[3], zMin=o[4], zMax=o[5], scale=o[6], S=int(o[7]))

    def dt_download(self):
        S3 = self.S ** 3
        dist = np.zeros(S3, np.float32)
        off = np.zeros((S3, 3), np.int16)
        near = np.zeros((S3, 3), np.int32)
        cellc = np.zeros(S3, np.int32)
        self.f("dt_download")(self.h, _ptr(dist, C.c_float), _ptr(off, C.c_short), _ptr(near, C.c_int), _ptr(cellc, C.c_int))
        return dist, off, near, cellc

    def dt_upload(self, dist=None, nearest=None):
        assert self.kind == "port"
        d = None if dist is None else _f32(dist)
        n = None if nearest is None else np.ascontiguousarray(nearest, dtype=np.int32)
        self.lib.orc_dt_upload(self.h, _ptr(d, C.c_float), _ptr(n, C.c_int))

    def dt_distance(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(xyz), np.float32)
        cell = np.zeros((len(xyz), 3), np.int32)
        self.f("dt_distance")(self.h, _ptr(xyz, C.c_double), len(xyz), _ptr(out, C.c_float), _ptr(cell, C.c_int))
        return out, cell

    def initialize(self):
        self.f("initialize")(self.h)

    def weights(self):
        w = np.zeros(self.Nd, np.float32)
        self.f("get_weights")(self.h, _ptr(w, C.c_float))
        return w

    def maxrotdis(self):
        w = np.zeros((20, self.Nd), np.float32)
        self.f("get_maxrotdis")(self.h, _ptr(w, C.c_float))
        return w

    def ssethresh(self):
        return self.f("get_ssethresh")(self.h)

    def inliernum(self):
        return self.f("get_inliernum")(self.h)

    def inner_bnb(self, R, level, opt_error, want_node=True):
        Rp = None if R is None else _f32(R).reshape(9)
        tn = np.zeros(4, np.float32)
        e = self.f("inner_bnb")(self.h, _ptr(Rp, C.c_float), level, opt_error, _ptr(tn, C.c_float) if want_node else None)
        return float(np.float32(e)), tn

    def eval_leaf(self, R, level, tcubes):
        assert self.kind == "port"
        Rp = None if R is None else _f32(R).reshape(9)
        tc = _f32(tcubes).reshape(-1, 4)
        n = len(tc)
        ub, lb = np.zeros(n, np.float32), np.zeros(n, np.float32)
        inc, fp = np.zeros((n, 2), np.int32), np.zeros((n, 2), np.int32)
        self.lib.orc_eval_leaf(self.h, _ptr(Rp, C.c_float), level, _ptr(tc, C.c_float), n, _ptr(ub, C.c_float), _ptr(lb, C.c_float),
                               _ptr(inc, C.c_int), _ptr(fp, C.c_int))
        return ub, lb, inc, fp

    def eval_inclusion(self, R, level, tcubes):
        """trimmed-error inclusion per cube.  port: (mask[n,Nd] uint8, resid[n,Nd]); ref: (first-k values after the reference's own
        intro_select [n,inlierNum], resid[n,Nd])"""
        Rp = _f32(R).reshape(9)
        tc = _f32(tcubes).reshape(-1, 4)
        n = len(tc)
        resid = np.zeros((n, self.Nd), np.float32)
        if self.kind == "port":
            mask = np.zeros((n, self.Nd), np.uint8)
            self.lib.orc_eval_inclusion(self.h, _ptr(Rp, C.c_float), level, _ptr(tc, C.c_float), n, mask.ctypes.data_as(C.POINTER(C.c_ubyte)), _ptr(resid, C.c_float))
            return mask, resid
        firstk = np.zeros((n, self.inliernum()), np.float32)
        self.lib.ref_eval_inclusion(self.h, _ptr(Rp, C.c_float), level, _ptr(tc, C.c_float), n, _ptr(firstk, C.c_float), _ptr(resid, C.c_float))
        return firstk, resid

    def icp(self, R, t):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(9).copy()
        t = np.ascontiguousarray(t, dtype=np.float64).reshape(3).copy()
        corr = np.zeros(self.Nd, np.int32)
        e = self.f("icp")(self.h, _ptr(R, C.c_double), _ptr(t, C.c_double), _ptr(corr, C.c_int))
        return float(np.float32(e)), R.reshape(3, 3), t, corr

    def register(self, nd=0):
        r = Result()
        buf = C.create_string_buffer(1 << 20)
        self.f("register")(self.h, nd, C.byref(r), buf, len(buf))
        if nd > 0:
            self.Nd = nd
        return dict(R=np.array(r.R).reshape(3, 3), t=np.array(r.t), optError=float(np.float32(r.optError)), optComp=int(r.optComp),
                    counters=list(r.counters), seconds_dt=r.seconds_dt, seconds_register=r.seconds_register,
                    trace=buf.value.decode(errors="replace"))


def error_trace(trace):
    """The 'Error*:' improvement values of a stdout trace."""
    out = []
    for line in trace.splitlines():
        if line.startswith("Error*:"):
            out.append(line.split()[1].rstrip(","))
    return out


# ---- Transformation helpers -------------------------------------------------------------------------------
def normalize(kind, xyz):
    lib, f = _lib(kind)
    a = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3).copy()
    mean = np.zeros(3)
    s = f("normalize")(_ptr(a, C.c_double), len(a), _ptr(mean, C.c_double))
    return a, mean, s


def scale(kind, xyz, sc):
    lib, f = _lib(kind)
    a = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3).copy()
    f("scale")(_ptr(a, C.c_double), len(a), sc)
    return a


def read_mol2_ref(path):
    lib, _ = _lib("ref")
    cap = 20000
    xyz = np.zeros((cap, 3))
    c = np.zeros(cap, np.int32)
    n = lib.ref_read_mol2(path.encode(), _ptr(xyz, C.c_double), _ptr(c, C.c_int), cap)
    if n < 0:
        raise FileNotFoundError(path)
    return xyz[:n].copy(), c[:n].copy()
