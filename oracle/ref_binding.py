"""ctypes binding of oracle/_ref/libofdg_ref.so: the REFERENCE'S OWN sources (compiled untouched from /root/reference by
oracle/ref_build.sh against the stand-in AGG / CImg / Caffe headers of oracle/shim) behind a small C interface
(oracle/ref_api.cpp). TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package.

/root/reference exists only in the build container; the built library travels to the GPU box with the snapshot, so
available() is True there as long as it was built here first (__graft_entry__.build() does that)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libofdg_ref.so")
REFERENCE = os.environ.get("OFDG_REFERENCE", "/root/reference")
W, H = 512, 384  # DGEN_WIDTH x DGEN_HEIGHT: compile-time constants of the reference (DataGenerator.h:55-56)

BLUEPRINT_DTYPE = np.dtype([
    ("obj_id", "<i4"), ("obj_type", "<i4"), ("init_rot", "<f4"), ("init_scale", "<f4"),
    ("init_trans_x", "<f4"), ("init_trans_y", "<f4"), ("rot", "<f4"), ("scale", "<f4"),
    ("trans_x", "<f4"), ("trans_y", "<f4"), ("tex_id", "<i4"), ("tex_rot", "<f4"), ("tex_scale", "<f4"),
    ("tex_shift_x", "<i4"), ("tex_shift_y", "<i4"), ("ellipse_scale_x", "<f4"), ("ellipse_scale_y", "<f4"),
    ("seg_begin", "<i4"), ("seg_count", "<i4"), ("comp_begin", "<i4"), ("comp_count", "<i4"),
    ("parent", "<i4"), ("is_additive_component", "<i4"), ("do_warpfield_deformation", "<i4"),
    ("field_id", "<i4"),
])


class TaskBatchStruct(C.Structure):
    _fields_ = [("n_tasks", C.c_int32), ("n_blueprints", C.c_int32), ("n_segments", C.c_int32),
                ("task_begin", C.c_void_p), ("blueprints", C.c_void_p), ("seg_type", C.c_void_p),
                ("seg_x", C.c_void_p), ("seg_y", C.c_void_p), ("augment", C.c_void_p)]


class DebugStruct(C.Structure):  # oracle_debug (oracle/oracle.h)
    _fields_ = [("id0", C.c_void_p), ("id1", C.c_void_p), ("masks", C.c_void_p), ("max_objs", C.c_int32),
                ("frames8", C.c_void_p), ("flow_bw", C.c_void_p), ("occlusion", C.c_void_p)]


def sources_present():
    return os.path.exists(os.path.join(REFERENCE, "src", "caffe", "DataGenerator.cpp"))


def build(force=False):
    """Runs oracle/ref_build.sh when the reference sources are present (the build container). Returns the library path or None."""
    if not sources_present():
        return LIB_PATH if os.path.exists(LIB_PATH) else None
    deps = [os.path.join(HERE, f) for f in ("ref_api.cpp", "ref_build.sh", "oracle.h", "shim/agg/agg_shim.h", "shim/thirdparty/CImg/CImg.h",
                                            "shim/caffe/ofdg_caffe_shim.hpp", "shim/caffe/proto/caffe.pb.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps):
        r = subprocess.run(["sh", os.path.join(HERE, "ref_build.sh")], capture_output=True, text=True, env=dict(os.environ, REF=REFERENCE))
        if r.returncode:
            raise RuntimeError("oracle/ref_build.sh failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def available():
    return os.path.exists(LIB_PATH) or (sources_present() and build() is not None)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            if build() is None:
                raise RuntimeError("oracle/_ref/libofdg_ref.so is missing and /root/reference is not here to build it from")
        L = C.CDLL(LIB_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_describe.restype = C.c_char_p
        L.ref_params_create.restype = C.c_void_p
        L.ref_params_create.argtypes = [C.c_int, C.c_int]
        L.ref_params_destroy.argtypes = [C.c_void_p]
        L.ref_params_clear.argtypes = [C.c_void_p]
        L.ref_params_generate.argtypes = [C.c_void_p, C.c_int]
        L.ref_params_view.argtypes = [C.c_void_p, C.POINTER(TaskBatchStruct)]
        L.ref_generator_create.restype = C.c_void_p
        L.ref_generator_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ref_generator_destroy.argtypes = [C.c_void_p]
        L.ref_generator_add_texture.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_generator_load_list.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_generator_texture_count.argtypes = [C.c_void_p]
        L.ref_generator_texture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_generator_set_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_randomized_crop.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.ref_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_generate_fields.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        L.ref_layer_create.restype = C.c_void_p
        L.ref_layer_create.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_layer_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_layer_top_shape.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_layer_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def describe():
    return lib().ref_describe().decode()


def _err():
    return RuntimeError("reference: " + lib().ref_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _copy(ptr, n, dt):
    if n == 0:
        return np.zeros(0, dtype=dt)
    buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt).copy()


class ParamStream:
    """The reference's ObjectParametersGenerator driven by the commission loop of its layer
    (data_generation_layer.cpp:197-214); seed_offset re-seeds engine k with seed_offset + k."""

    def __init__(self, mode, seed_offset=0):
        self._h = lib().ref_params_create(mode, seed_offset)
        if not self._h:
            raise _err()

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_params_destroy(self._h)
            self._h = None

    def generate(self, n):
        """Draws n more tasks and returns ALL tasks drawn since the last clear() as flat arrays (ofdg_task_batch layout)."""
        if lib().ref_params_generate(self._h, n):
            raise _err()
        s = TaskBatchStruct()
        lib().ref_params_view(self._h, C.byref(s))
        return {"task_begin": _copy(s.task_begin, s.n_tasks + 1, "<i4"), "blueprints": _copy(s.blueprints, s.n_blueprints, BLUEPRINT_DTYPE),
                "seg_type": _copy(s.seg_type, s.n_segments, "<i4"), "seg_x": _copy(s.seg_x, s.n_segments, "<f4"),
                "seg_y": _copy(s.seg_y, s.n_segments, "<f4"), "augment": None}

    def clear(self):
        lib().ref_params_clear(self._h)


class Generator:
    """The reference's DataGenerator, driven task by task like one of its worker threads."""

    def __init__(self, mode, use_aa=True, second_level_threads=1, textures=None, fields=None):
        self._h = lib().ref_generator_create(mode, int(use_aa), second_level_threads)
        if not self._h:
            raise _err()
        self.mode = mode
        if textures is not None:
            self.add_textures(textures)
        if fields is not None:
            self.set_fields(fields)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_generator_destroy(self._h)
            self._h = None

    def add_textures(self, textures):
        """n x 3 x h x w uint8 (or a list of 3 x h x w arrays), planes in the order the reference holds after its R<->B swap."""
        for t in textures:
            t = np.ascontiguousarray(t, np.uint8)
            lib().ref_generator_add_texture(self._h, _p(t), t.shape[2], t.shape[1])

    def load_list(self, path):
        if lib().ref_generator_load_list(self._h, path.encode()):
            raise _err()

    def texture_count(self):
        return lib().ref_generator_texture_count(self._h)

    def texture(self, i):
        w, h = C.c_int(), C.c_int()
        lib().ref_generator_texture(self._h, i, None, C.byref(w), C.byref(h))
        out = np.empty((3, h.value, w.value), np.uint8)
        lib().ref_generator_texture(self._h, i, _p(out), C.byref(w), C.byref(h))
        return out

    def set_fields(self, fields):
        fields = np.ascontiguousarray(fields, np.float32)
        assert fields.shape[1:] == (2, 2, H + 1, W + 1)
        lib().ref_generator_set_fields(self._h, _p(fields), fields.shape[0])

    def randomized_crop(self, tex_id, out_w, out_h, angle=0.0, zoom=1.0, shift_x=0, shift_y=0):
        out = np.empty((3, out_h, out_w), np.uint8)
        if lib().ref_randomized_crop(self._h, tex_id, out_w, out_h, angle, zoom, shift_x, shift_y, _p(out)):
            raise _err()
        return out

    def render(self, task_struct, debug=False, max_objs=24, field_policy=0):
        """task_struct: any ctypes struct with ofdg_task_batch's layout. Returns the same dictionary as oracle.binding.render
        (without 'occlusion', which is not a reference output)."""
        n = task_struct.n_tasks
        out = {"img0": np.empty((n, 3, H, W), np.float32), "img1": np.empty((n, 3, H, W), np.float32), "flow": np.empty((n, 2, H, W), np.float32)}
        dbg = None
        if debug:
            out["id0"] = np.empty((n, H, W), np.uint32)
            out["id1"] = np.empty((n, H, W), np.uint32)
            out["masks"] = np.zeros((n, max_objs, 4, H, W), np.uint8)
            out["frames8"] = np.empty((n, 2, 3, H, W), np.uint8)
            out["flow_bw"] = np.empty((n, 2, H, W), np.float32)
            dbg = DebugStruct(out["id0"].ctypes.data, out["id1"].ctypes.data, out["masks"].ctypes.data, max_objs, out["frames8"].ctypes.data,
                              out["flow_bw"].ctypes.data, None)
        rc = lib().ref_render(self._h, C.addressof(task_struct), _p(out["img0"]), _p(out["img1"]), _p(out["flow"]),
                              C.addressof(dbg) if dbg is not None else None, field_policy)
        if rc:
            raise _err()
        return out


class Layer:
    """The reference's DataGenerationLayer<float>, created through its own REGISTER_LAYER_CLASS registration."""

    def __init__(self, mode, texture_list, batch, prefetch=4, first_level_threads=16, second_level_threads=1, use_aa=True):
        self._h = lib().ref_layer_create(mode, texture_list.encode(), batch, prefetch, first_level_threads, second_level_threads, int(use_aa))
        if not self._h:
            raise _err()
        self.batch = batch

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_layer_destroy(self._h)
            self._h = None

    __del__ = close

    def top_shape(self, i):
        s = (C.c_int * 4)()
        lib().ref_layer_top_shape(self._h, i, s)
        return tuple(s)

    def forward(self, copy=True):
        if not copy:
            if lib().ref_layer_forward(self._h, None, None, None):
                raise _err()
            return None
        out = [np.empty((self.batch, 3, H, W), np.float32), np.empty((self.batch, 3, H, W), np.float32), np.empty((self.batch, 2, H, W), np.float32)]
        if lib().ref_layer_forward(self._h, _p(out[0]), _p(out[1]), _p(out[2])):
            raise _err()
        return out


def generate_fields(seed=1, n_fields=8):
    """(n, 2, 2, H+1, W+1) float32 pool of (flow, iflow) crops from the reference's own DisplacementComposer / FlowField, with
    the displacer scene drawn as CropGenerator::worker_thread_loop draws it but from std::mt19937(seed)."""
    out = np.empty((n_fields, 2, 2, H + 1, W + 1), np.float32)
    if lib().ref_generate_fields(seed, n_fields, _p(out)):
        raise _err()
    return out


def write_ppm_pool(directory, textures_bgr):
    """Writes textures (n x 3 x h x w, planes B,G,R as the generators hold them) as binary PPM files (R,G,B on disk, what the
    reference's loader un-swaps) plus the list file the reference's TextureCollection reads. Returns the list path."""
    os.makedirs(directory, exist_ok=True)
    names = []
    for i, t in enumerate(textures_bgr):
        t = np.ascontiguousarray(t, np.uint8)
        rgb = np.stack([t[2], t[1], t[0]], axis=-1)
        p = os.path.join(directory, "tex%05d.ppm" % i)
        with open(p, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (t.shape[2], t.shape[1]))
            f.write(rgb.tobytes())
        names.append(p)
    lst = os.path.join(directory, "database.txt")
    with open(lst, "w") as f:
        f.write("".join(n + "\n" for n in names))
    return lst
