"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")


class OracleConfig(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("mode", C.c_int32), ("use_antialiasing", C.c_int32),
                ("n_tex", C.c_int32), ("tex_w", C.c_int32), ("tex_h", C.c_int32), ("n_fields", C.c_int32),
                ("n_threads", C.c_int32), ("faithful_copies", C.c_int32), ("tex_sizes", C.c_void_p), ("tex_offsets", C.c_void_p)]


class OracleDebug(C.Structure):
    _fields_ = [("id0", C.c_void_p), ("id1", C.c_void_p), ("masks", C.c_void_p), ("max_objs", C.c_int32),
                ("frames8", C.c_void_p), ("flow_bw", C.c_void_p), ("occlusion", C.c_void_p)]


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("oracle.cpp", "warpfields.cpp", "oracle.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.run(["make", "-C", HERE, "-B", "CXX=g++"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_last_error.restype = C.c_char_p
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def render(task_struct, textures, W=512, H=384, mode=1, use_aa=True, fields=None, n_threads=8, faithful=False,
           debug=False, max_objs=24):
    """task_struct: ofdg_b200.TaskBatchStruct (same C layout). textures: n x 3 x th x tw uint8, or a list of
    3 x h x w uint8 arrays of different sizes."""
    sizes = offsets = None
    if isinstance(textures, (list, tuple)):
        parts = [np.ascontiguousarray(t, np.uint8) for t in textures]
        sizes = np.array([[t.shape[2], t.shape[1]] for t in parts], np.int32)
        offsets = np.cumsum([0] + [t.size for t in parts[:-1]]).astype(np.uint64)
        textures = np.concatenate([t.ravel() for t in parts])
        n_tex, th, tw = len(parts), 0, 0
    else:
        textures = np.ascontiguousarray(textures, np.uint8)
        n_tex, _, th, tw = textures.shape
    cfg = OracleConfig(W, H, mode, int(use_aa), n_tex, tw, th, 0 if fields is None else fields.shape[0], n_threads, int(faithful),
                       None if sizes is None else sizes.ctypes.data, None if offsets is None else offsets.ctypes.data)
    n = task_struct.n_tasks
    out = {"img0": np.empty((n, 3, H, W), np.float32), "img1": np.empty((n, 3, H, W), np.float32),
           "flow": np.empty((n, 2, H, W), np.float32)}
    dbg = None
    if debug:
        out["id0"] = np.empty((n, H, W), np.uint32)
        out["id1"] = np.empty((n, H, W), np.uint32)
        out["masks"] = np.zeros((n, max_objs, 4, H, W), np.uint8)
        out["frames8"] = np.empty((n, 2, 3, H, W), np.uint8)
        out["flow_bw"] = np.empty((n, 2, H, W), np.float32)
        out["occlusion"] = np.empty((n, 1, H, W), np.float32)
        dbg = OracleDebug(out["id0"].ctypes.data, out["id1"].ctypes.data, out["masks"].ctypes.data, max_objs, out["frames8"].ctypes.data,
                          out["flow_bw"].ctypes.data, out["occlusion"].ctypes.data)
    if fields is not None:
        fields = np.ascontiguousarray(fields, np.float32)
    rc = lib().oracle_render(C.byref(cfg), C.byref(task_struct), _p(textures), _p(fields), _p(out["img0"]), _p(out["img1"]),
                             _p(out["flow"]), C.byref(dbg) if dbg is not None else None)
    if rc:
        raise RuntimeError("oracle: " + lib().oracle_last_error().decode())
    return out


def raster_polygon(xy, W, H, aa=True):
    xy = np.ascontiguousarray(xy, np.float64)
    mask = np.empty((H, W), np.uint8)
    rc = lib().oracle_raster_polygon(_p(xy), C.c_int32(len(xy)), C.c_int32(W), C.c_int32(H), C.c_int32(int(aa)), _p(mask))
    assert rc == 0, lib().oracle_last_error()
    return mask


def raster_fixed(xy, W, H, aa=True):
    xy = np.ascontiguousarray(xy, np.int32)
    mask = np.empty((H, W), np.uint8)
    rc = lib().oracle_raster_fixed(_p(xy), C.c_int32(len(xy)), C.c_int32(W), C.c_int32(H), C.c_int32(int(aa)), _p(mask))
    assert rc == 0, lib().oracle_last_error()
    return mask


def transform_texture(img, m):
    img = np.ascontiguousarray(img, np.uint8)
    _, h, w = img.shape
    m = np.ascontiguousarray(m, np.float64)
    out = np.empty_like(img)
    rc = lib().oracle_transform_texture(_p(img), C.c_int32(w), C.c_int32(h), _p(m), _p(out))
    assert rc == 0, lib().oracle_last_error()
    return out


def randomized_crop(tex, out_w, out_h, angle=0.0, zoom=1.0, shift_x=0, shift_y=0):
    tex = np.ascontiguousarray(tex, np.uint8)
    _, th, tw = tex.shape
    out = np.empty((3, out_h, out_w), np.uint8)
    rc = lib().oracle_randomized_crop(_p(tex), C.c_int32(tw), C.c_int32(th), C.c_int32(out_w), C.c_int32(out_h),
                                      C.c_float(angle), C.c_float(zoom), C.c_int32(shift_x), C.c_int32(shift_y), _p(out))
    assert rc == 0, lib().oracle_last_error()
    return out


def composite_luts():
    a = np.empty((256, 256), np.uint8)
    s = np.empty((256, 256), np.uint8)
    lib().oracle_composite_luts(_p(a), _p(s))
    return a, s


def generate_fields(W=512, H=384, seed=1, n_fields=8):
    """(n, 2, 2, H+1, W+1) float32 pool of (flow, iflow) crops, WarpFields::CropGenerator restated."""
    out = np.empty((n_fields, 2, 2, H + 1, W + 1), np.float32)
    rc = lib().oracle_generate_fields(C.c_int32(W), C.c_int32(H), C.c_uint32(seed), C.c_int32(n_fields), _p(out))
    assert rc == 0
    return out
