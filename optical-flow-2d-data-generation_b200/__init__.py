"""ofdg_b200 -- Python face of the B200-native optical-flow data generator.

Thin ctypes binding over the C ABI in include/ofdg/ofdg.h (the native library is
csrc/libofdg.so, built in-tree by build.py). PyTorch appears only as the owner of device
memory and streams. There is no CPU fallback: rendering without the native library or
without a B200 raises.

The directory name is not a valid Python identifier; import it through the repo-root shim:
    import ofdg_b200
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libofdg.so")

OBJ_DUMMY, OBJ_ELLIPSE, OBJ_POLYGON, OBJ_COMPOSITE = 0, 1, 2, 3
SEG_DUMMY, SEG_LINE, SEG_CURVE3 = 0, 1, 3
NUM_SLOTS = 45

# numpy view of ofdg_blueprint (include/ofdg/scene.h)
BLUEPRINT_DTYPE = np.dtype([
    ("obj_id", "<i4"), ("obj_type", "<i4"), ("init_rot", "<f4"), ("init_scale", "<f4"),
    ("init_trans_x", "<f4"), ("init_trans_y", "<f4"), ("rot", "<f4"), ("scale", "<f4"),
    ("trans_x", "<f4"), ("trans_y", "<f4"), ("tex_id", "<i4"), ("tex_rot", "<f4"), ("tex_scale", "<f4"),
    ("tex_shift_x", "<i4"), ("tex_shift_y", "<i4"), ("ellipse_scale_x", "<f4"), ("ellipse_scale_y", "<f4"),
    ("seg_begin", "<i4"), ("seg_count", "<i4"), ("comp_begin", "<i4"), ("comp_count", "<i4"),
    ("parent", "<i4"), ("is_additive_component", "<i4"), ("do_warpfield_deformation", "<i4"),
    ("field_id", "<i4"),
])
assert BLUEPRINT_DTYPE.itemsize == 100


class TaskBatchStruct(C.Structure):
    _fields_ = [("n_tasks", C.c_int32), ("n_blueprints", C.c_int32), ("n_segments", C.c_int32),
                ("task_begin", C.c_void_p), ("blueprints", C.c_void_p), ("seg_type", C.c_void_p),
                ("seg_x", C.c_void_p), ("seg_y", C.c_void_p), ("augment", C.c_void_p)]


# numpy view of ofdg_augment (include/ofdg/scene.h)
AUGMENT_DTYPE = np.dtype([("enabled", "<i4"), ("gain", "<f4", (3,)), ("brightness", "<f4"), ("contrast", "<f4"),
                          ("noise_sigma", "<f4"), ("noise_seed", "<u4", (2,))])
assert AUGMENT_DTYPE.itemsize == 36


class ExtraTopsStruct(C.Structure):
    _fields_ = [("flow_bw", C.c_void_p), ("id0", C.c_void_p), ("id1", C.c_void_p), ("occlusion", C.c_void_p)]


class ConfigStruct(C.Structure):
    _fields_ = [("device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("mode", C.c_int32),
                ("use_antialiasing", C.c_int32), ("max_batch", C.c_int32), ("reserved", C.c_int32 * 8)]


class OfdgError(RuntimeError):
    pass


_lib = None

# every symbol include/ofdg/ofdg.h declares
EXPORTS = [
    "ofdg_last_error", "ofdg_version", "ofdg_params_create", "ofdg_params_destroy", "ofdg_params_generate",
    "ofdg_params_skip", "ofdg_params_set_threads", "ofdg_params_enable_augmentation", "ofdg_params_tasks_generated", "ofdg_params_draws", "ofdg_params_field_draws", "ofdg_refresh_fields", "ofdg_reserve_fields", "ofdg_params_slot_name",
    "ofdg_tasks_create", "ofdg_tasks_destroy", "ofdg_tasks_clear", "ofdg_tasks_view", "ofdg_tasks_assign",
    "ofdg_flatten_ellipse", "ofdg_flatten_polygon", "ofdg_debug_raster_host", "ofdg_debug_expand_host", "ofdg_create", "ofdg_destroy", "ofdg_upload_textures", "ofdg_add_textures", "ofdg_clear_textures", "ofdg_texture_size", "ofdg_download_foreground_view",
    "ofdg_synth_textures", "ofdg_download_texture", "ofdg_set_fields", "ofdg_generate_fields", "ofdg_render", "ofdg_render_host",
    "ofdg_render_debug", "ofdg_debug_background", "ofdg_debug_composite_luts", "ofdg_prepare", "ofdg_prepared_destroy",
    "ofdg_render_prepared", "ofdg_render_prepared_host", "ofdg_generate", "ofdg_generate_host", "ofdg_generate_philox", "ofdg_philox_tasks", "ofdg_launch_count", "ofdg_kernel_times", "ofdg_last_shade_ms", "ofdg_last_bin_ms", "ofdg_last_render_stats", "ofdg_last_raster_ms", "ofdg_last_upload_bytes", "ofdg_last_download_bytes", "ofdg_set_extra_tops",
]
# include/ofdg/layer.h
LAYER_EXPORTS = [
    "ofdg_layer_last_error", "ofdg_layer_parse_prototxt", "ofdg_layer_registered_types", "ofdg_layer_producer_stats", "ofdg_layer_create", "ofdg_layer_destroy", "ofdg_layer_setup",
    "ofdg_layer_top_shape", "ofdg_layer_forward", "ofdg_layer_top_data", "ofdg_layer_type", "ofdg_decode_texture_file", "ofdg_read_texture_list",
]


def lib():
    """Loads csrc/libofdg.so; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OfdgError(f"{LIB_PATH} is missing: run `python {os.path.join(_HERE, 'build.py')}` "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.ofdg_last_error.restype = C.c_char_p
        L.ofdg_params_slot_name.restype = C.c_char_p
        L.ofdg_params_slot_name.argtypes = [C.c_int32]
        L.ofdg_params_tasks_generated.restype = C.c_uint64
        L.ofdg_params_tasks_generated.argtypes = [C.c_void_p]
        L.ofdg_params_draws.restype = C.c_uint64
        L.ofdg_params_draws.argtypes = [C.c_void_p, C.c_int32]
        L.ofdg_launch_count.restype = C.c_uint64
        L.ofdg_launch_count.argtypes = [C.c_void_p]
        L.ofdg_last_upload_bytes.restype = C.c_uint64
        L.ofdg_last_upload_bytes.argtypes = [C.c_void_p]
        L.ofdg_last_download_bytes.restype = C.c_uint64
        L.ofdg_last_download_bytes.argtypes = [C.c_void_p]
        L.ofdg_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.ofdg_last_shade_ms.argtypes = [C.c_void_p]
        L.ofdg_last_shade_ms.restype = C.c_double
        L.ofdg_params_set_threads.argtypes = [C.c_void_p, C.c_int32]
        L.ofdg_params_field_draws.argtypes = [C.c_void_p]
        L.ofdg_params_field_draws.restype = C.c_uint64
        L.ofdg_refresh_fields.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32]
        L.ofdg_reserve_fields.argtypes = [C.c_void_p, C.c_int32]
        L.ofdg_last_render_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        for f in (L.ofdg_last_bin_ms, L.ofdg_last_raster_ms):
            f.argtypes = [C.c_void_p]
            f.restype = C.c_double
        L.ofdg_params_create.argtypes = [C.c_int32] * 6 + [C.POINTER(C.c_void_p)]
        L.ofdg_params_destroy.argtypes = [C.c_void_p]
        L.ofdg_params_generate.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ofdg_params_skip.argtypes = [C.c_void_p, C.c_uint64]
        L.ofdg_params_enable_augmentation.argtypes = [C.c_void_p, C.c_int32]
        L.ofdg_tasks_create.argtypes = [C.POINTER(C.c_void_p)]
        L.ofdg_tasks_destroy.argtypes = [C.c_void_p]
        L.ofdg_tasks_clear.argtypes = [C.c_void_p]
        L.ofdg_tasks_view.argtypes = [C.c_void_p, C.POINTER(TaskBatchStruct)]
        L.ofdg_tasks_assign.argtypes = [C.c_void_p, C.POINTER(TaskBatchStruct)]
        L.ofdg_flatten_ellipse.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int32]
        L.ofdg_flatten_polygon.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.ofdg_debug_raster_host.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.ofdg_debug_expand_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32]
        L.ofdg_create.argtypes = [C.POINTER(ConfigStruct), C.POINTER(C.c_void_p)]
        L.ofdg_destroy.argtypes = [C.c_void_p]
        L.ofdg_upload_textures.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.ofdg_add_textures.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.ofdg_clear_textures.argtypes = [C.c_void_p]
        L.ofdg_texture_size.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.ofdg_download_foreground_view.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ofdg_synth_textures.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint64]
        L.ofdg_download_texture.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ofdg_set_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.ofdg_generate_fields.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p]
        L.ofdg_set_extra_tops.argtypes = [C.c_void_p, C.c_void_p]
        L.ofdg_render.argtypes = [C.c_void_p, C.POINTER(TaskBatchStruct)] + [C.c_void_p] * 4
        L.ofdg_render_host.argtypes = [C.c_void_p, C.POINTER(TaskBatchStruct)] + [C.c_void_p] * 3
        L.ofdg_render_debug.argtypes = [C.c_void_p, C.POINTER(TaskBatchStruct)] + [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 3
        L.ofdg_debug_background.argtypes = [C.c_void_p, C.POINTER(TaskBatchStruct), C.c_void_p, C.c_void_p]
        L.ofdg_debug_composite_luts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ofdg_prepare.argtypes = [C.c_void_p, C.POINTER(TaskBatchStruct), C.POINTER(C.c_void_p)]
        L.ofdg_prepared_destroy.argtypes = [C.c_void_p]
        L.ofdg_render_prepared.argtypes = [C.c_void_p] * 6
        L.ofdg_render_prepared_host.argtypes = [C.c_void_p] * 5
        L.ofdg_generate_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 3
        L.ofdg_generate_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32] + [C.c_void_p] * 4
        L.ofdg_philox_tasks.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p]
        L.ofdg_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 4
        L.ofdg_layer_last_error.restype = C.c_char_p
        L.ofdg_layer_type.restype = C.c_char_p
        L.ofdg_layer_type.argtypes = [C.c_void_p]
        L.ofdg_layer_parse_prototxt.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]
        L.ofdg_layer_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.ofdg_layer_destroy.argtypes = [C.c_void_p]
        L.ofdg_layer_setup.argtypes = [C.c_void_p]
        L.ofdg_layer_top_shape.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ofdg_layer_forward.argtypes = [C.c_void_p, C.c_int32]
        L.ofdg_layer_top_data.restype = C.c_void_p
        L.ofdg_layer_top_data.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise OfdgError(f"ofdg error {rc}: {lib().ofdg_last_error().decode()}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Tasks:
    """An owned task batch (the reference's queue of TaskBucket*, DataGenerator.h:423-437)."""

    def __init__(self):
        self._h = C.c_void_p()
        _check(lib().ofdg_tasks_create(C.byref(self._h)))
        self._keep = None

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ofdg_tasks_destroy(self._h)
            self._h = None

    def clear(self):
        lib().ofdg_tasks_clear(self._h)

    def struct(self):
        s = TaskBatchStruct()
        _check(lib().ofdg_tasks_view(self._h, C.byref(s)))
        s._owner = self  # the view borrows this batch's memory: keep it alive as long as the view
        return s

    def __len__(self):
        return self.struct().n_tasks

    def arrays(self):
        """Copies of the five flat arrays as numpy."""
        s = self.struct()

        def cp(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dtype=dt)
            buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dt).copy()
        return {
            "task_begin": cp(s.task_begin, s.n_tasks + 1, "<i4"),
            "blueprints": cp(s.blueprints, s.n_blueprints, BLUEPRINT_DTYPE),
            "seg_type": cp(s.seg_type, s.n_segments, "<i4"),
            "seg_x": cp(s.seg_x, s.n_segments, "<f4"),
            "seg_y": cp(s.seg_y, s.n_segments, "<f4"),
            "augment": cp(s.augment, s.n_tasks, AUGMENT_DTYPE) if s.augment else None,
        }

    @staticmethod
    def from_arrays(arrs):
        t = Tasks()
        s, keep = struct_from_arrays(arrs)
        _check(lib().ofdg_tasks_assign(t._h, C.byref(s)))
        return t

    def select(self, idx):
        """New batch holding only the tasks in idx (indices rebased)."""
        return Tasks.from_arrays(select_tasks(self.arrays(), idx))


def struct_from_arrays(arrs):
    keep = {k: np.ascontiguousarray(v) for k, v in arrs.items() if v is not None}
    s = TaskBatchStruct()
    s.n_tasks = len(keep["task_begin"]) - 1
    s.n_blueprints = len(keep["blueprints"])
    s.n_segments = len(keep["seg_type"])
    s.task_begin = keep["task_begin"].ctypes.data
    s.blueprints = keep["blueprints"].ctypes.data
    s.seg_type = keep["seg_type"].ctypes.data
    s.seg_x = keep["seg_x"].ctypes.data
    s.seg_y = keep["seg_y"].ctypes.data
    s.augment = keep["augment"].ctypes.data if "augment" in keep else None
    s._owner = keep
    return s, keep


def select_tasks(arrs, idx):
    tb = arrs["task_begin"]
    bps, st, sx, sy, begin = [], [], [], [], [0]
    nb = 0
    ns = 0
    for t in idx:
        b = arrs["blueprints"][tb[t]:tb[t + 1]].copy()
        for r in b:
            if r["seg_count"] > 0:
                s0, n = int(r["seg_begin"]), int(r["seg_count"])
                st.append(arrs["seg_type"][s0:s0 + n]); sx.append(arrs["seg_x"][s0:s0 + n]); sy.append(arrs["seg_y"][s0:s0 + n])
                r["seg_begin"] = ns
                ns += n
            if r["comp_count"] > 0:
                r["comp_begin"] = r["comp_begin"] - tb[t] + nb
            if r["parent"] >= 0:
                r["parent"] = r["parent"] - tb[t] + nb
        bps.append(b)
        nb += len(b)
        begin.append(nb)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return {"task_begin": np.asarray(begin, "<i4"), "blueprints": cat(bps, BLUEPRINT_DTYPE),
            "seg_type": cat(st, "<i4"), "seg_x": cat(sx, "<f4"), "seg_y": cat(sy, "<f4"),
            "augment": None if arrs.get("augment") is None else arrs["augment"][list(idx)].copy()}


class ParamStream:
    """Host RNG parameter stream (ObjectParametersGenerator, DataGenerator.cpp:1353-2835)."""

    def __init__(self, mode, width=512, height=384, seed_offset=0, n_fields=0, fg_override=0):
        self._h = C.c_void_p()
        _check(lib().ofdg_params_create(mode, width, height, seed_offset, n_fields, fg_override, C.byref(self._h)))
        self.mode = mode

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ofdg_params_destroy(self._h)
            self._h = None

    def generate(self, n, out=None):
        out = Tasks() if out is None else out
        _check(lib().ofdg_params_generate(self._h, n, out._h))
        return out

    def skip(self, n):
        _check(lib().ofdg_params_skip(self._h, n))

    def set_threads(self, n):
        """Look-ahead helper threads of generate() (ofdg_params_set_threads); the stream itself does not change."""
        _check(lib().ofdg_params_set_threads(self._h, n))

    def enable_augmentation(self, on=True):
        _check(lib().ofdg_params_enable_augmentation(self._h, int(bool(on))))

    def tasks_generated(self):
        return int(lib().ofdg_params_tasks_generated(self._h))

    def draws(self):
        return [int(lib().ofdg_params_draws(self._h, i)) for i in range(NUM_SLOTS)]

    def field_draws(self):
        """Mode 9: warp-field picks so far; pick k reads pool slot (k // 3) % n_fields."""
        return int(lib().ofdg_params_field_draws(self._h))


def slot_names():
    return [lib().ofdg_params_slot_name(i).decode() for i in range(NUM_SLOTS)]


def flatten_ellipse(rx, ry, m):
    m = np.ascontiguousarray(m, dtype=np.float64)
    xy = np.zeros((4096, 2), dtype=np.int32)
    n = lib().ofdg_flatten_ellipse(rx, ry, _ptr(m), _ptr(xy), 4096)
    if n < 0:
        raise OfdgError(lib().ofdg_last_error().decode())
    return xy[:n].copy()


def flatten_polygon(seg_type, seg_x, seg_y, m):
    st = np.ascontiguousarray(seg_type, dtype=np.int32)
    sx = np.ascontiguousarray(seg_x, dtype=np.float32)
    sy = np.ascontiguousarray(seg_y, dtype=np.float32)
    m = np.ascontiguousarray(m, dtype=np.float64)
    xy = np.zeros((65536, 2), dtype=np.int32)
    n = lib().ofdg_flatten_polygon(_ptr(st), _ptr(sx), _ptr(sy), len(st), _ptr(m), _ptr(xy), 65536)
    if n < 0:
        raise OfdgError(lib().ofdg_last_error().decode())
    return xy[:n].copy()


def raster_host(xy, W, H, aa=True):
    """The render kernel's tile rasteriser run on the host (csrc/raster_tile.h)."""
    xy = np.ascontiguousarray(xy, dtype=np.int32)
    mask = np.empty((H, W), np.uint8)
    _check(lib().ofdg_debug_raster_host(_ptr(xy), len(xy), W, H, int(aa), _ptr(mask)))
    return mask


def decode_texture_file(path):
    """The layer's texture-file decoder (binary PPM, uncompressed BMP, 8-bit PNG): (3, h, w) uint8, planes B,G,R."""
    L = lib()
    L.ofdg_decode_texture_file.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_uint64]
    w, h = C.c_int32(), C.c_int32()
    if L.ofdg_decode_texture_file(str(path).encode(), C.byref(w), C.byref(h), None, 0):
        raise OfdgError(L.ofdg_layer_last_error().decode())
    out = np.empty((3, h.value, w.value), np.uint8)
    if L.ofdg_decode_texture_file(str(path).encode(), C.byref(w), C.byref(h), _ptr(out), out.size):
        raise OfdgError(L.ofdg_layer_last_error().decode())
    return out


def read_texture_list(path):
    """Paths of a texture list file as the layer reads it (reference semantics: an unterminated last line is dropped)."""
    buf = C.create_string_buffer(1 << 20)
    n = C.c_int32()
    L = lib()
    L.ofdg_read_texture_list.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(C.c_int32)]
    if L.ofdg_read_texture_list(str(path).encode(), buf, 1 << 20, C.byref(n)):
        raise OfdgError(L.ofdg_layer_last_error().decode())
    out = [p for p in buf.value.decode().split("\n") if p]
    assert len(out) == n.value
    return out


def expand_host(src, dst, streaming=True):
    """The host routine of the uint8 transport (csrc/host/expand.cpp): dst[i] = float(src[i]), in place into `dst`."""
    assert src.dtype == np.uint8 and dst.dtype == np.float32 and src.size == dst.size
    assert src.flags["C_CONTIGUOUS"] and dst.flags["C_CONTIGUOUS"]
    _check(lib().ofdg_debug_expand_host(_ptr(src), _ptr(dst), src.size, int(streaming)))
    return dst


class Generator:
    """Device-side generator (DataGenerator::DataGenerator, DataGenerator.h:449-500): owns the HBM texture pool
    and renders task batches into caller-provided device blobs."""

    def __init__(self, device=0, width=512, height=384, mode=1, use_antialiasing=True, max_batch=64):
        cfg = ConfigStruct()
        cfg.device, cfg.width, cfg.height, cfg.mode = device, width, height, mode
        cfg.use_antialiasing, cfg.max_batch = int(bool(use_antialiasing)), max_batch
        self._h = C.c_void_p()
        _check(lib().ofdg_create(C.byref(cfg), C.byref(self._h)))
        self.W, self.H, self.mode, self.device, self.max_batch = width, height, mode, device, max_batch
        self.tex = None

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ofdg_destroy(self._h)
            self._h = None

    __del__ = close

    # -- texture pool
    def upload_textures(self, planar):
        """Replace the pool. `planar`: (n, 3, h, w) uint8, or a list of (3, h, w) arrays of different sizes."""
        if isinstance(planar, (list, tuple)):
            _check(lib().ofdg_clear_textures(self._h))
            for t in planar:
                self.add_textures(np.asarray(t)[None])
            return
        planar = np.ascontiguousarray(planar, dtype=np.uint8)
        n, c, h, w = planar.shape
        assert c == 3
        _check(lib().ofdg_upload_textures(self._h, _ptr(planar), n, w, h))

    def add_textures(self, planar):
        """Append (n, 3, h, w) uint8 textures of one size to the pool (ofdg_add_textures)."""
        planar = np.ascontiguousarray(planar, dtype=np.uint8)
        n, c, h, w = planar.shape
        assert c == 3
        _check(lib().ofdg_add_textures(self._h, _ptr(planar), n, w, h))

    def synth_textures(self, n, w=1024, h=768, seed=0):
        _check(lib().ofdg_synth_textures(self._h, n, w, h, seed))

    def texture_size(self, i):
        w, h = C.c_int32(), C.c_int32()
        _check(lib().ofdg_texture_size(self._h, i, C.byref(w), C.byref(h)))
        return w.value, h.value

    def download_texture(self, i):
        w, h = self.texture_size(i)
        out = np.empty((3, h, w), dtype=np.uint8)
        _check(lib().ofdg_download_texture(self._h, i, _ptr(out)))
        return out

    def download_foreground_view(self, i):
        """The (3, H, W) view of pool texture i that foreground objects are textured from."""
        out = np.empty((3, self.H, self.W), dtype=np.uint8)
        _check(lib().ofdg_download_foreground_view(self._h, i, _ptr(out)))
        return out

    def set_fields(self, fields):
        fields = np.ascontiguousarray(fields, dtype=np.float32)
        assert fields.shape[1:] == (2, 2, self.H + 1, self.W + 1), fields.shape
        _check(lib().ofdg_set_fields(self._h, _ptr(fields), fields.shape[0]))

    def generate_fields(self, seed, n):
        """Mode-9 field pool produced on the GPU and installed; returns a host copy (n, 2, 2, H+1, W+1)."""
        out = np.empty((n, 2, 2, self.H + 1, self.W + 1), np.float32)
        _check(lib().ofdg_generate_fields(self._h, seed, n, _ptr(out)))
        return out

    def reserve_fields(self, total):
        _check(lib().ofdg_reserve_fields(self._h, total))

    def refresh_fields(self, seed, first_slot, n):
        """Regenerates pool slots [first_slot, first_slot + n) in place on the GPU (ofdg_refresh_fields)."""
        _check(lib().ofdg_refresh_fields(self._h, seed, first_slot, n))

    # -- rendering
    def render(self, tasks, img0, img1, flow, stream=None):
        """img0/img1/flow: CUDA float32 tensors (N,3,H,W), (N,3,H,W), (N,2,H,W) -- the top blobs."""
        s = tasks.struct()
        _check(lib().ofdg_render(self._h, C.byref(s), img0.data_ptr(), img1.data_ptr(), flow.data_ptr(), stream))

    def render_host(self, tasks, img0=None, img1=None, flow=None):
        s = tasks.struct()
        n = s.n_tasks
        img0 = np.empty((n, 3, self.H, self.W), np.float32) if img0 is None else img0
        img1 = np.empty((n, 3, self.H, self.W), np.float32) if img1 is None else img1
        flow = np.empty((n, 2, self.H, self.W), np.float32) if flow is None else flow
        p = lambda a: a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        _check(lib().ofdg_render_host(self._h, C.byref(s), p(img0), p(img1), p(flow)))
        return img0, img1, flow

    def render_debug(self, tasks, max_objs=24, want_masks=True):
        s = tasks.struct()
        n, H, W = s.n_tasks, self.H, self.W
        out = {
            "img0": np.empty((n, 3, H, W), np.float32), "img1": np.empty((n, 3, H, W), np.float32),
            "flow": np.empty((n, 2, H, W), np.float32),
            "masks": np.empty((n, max_objs, 4, H, W), np.uint8) if want_masks else None,
            "id0": np.empty((n, H, W), np.uint32), "id1": np.empty((n, H, W), np.uint32),
            "frames8": np.empty((n, 2, 3, H, W), np.uint8),
        }
        _check(lib().ofdg_render_debug(self._h, C.byref(s), _ptr(out["img0"]), _ptr(out["img1"]), _ptr(out["flow"]),
                                       _ptr(out["masks"]), max_objs if want_masks else 0, _ptr(out["id0"]), _ptr(out["id1"]),
                                       _ptr(out["frames8"])))
        return out

    def debug_background(self, tasks):
        s = tasks.struct()
        out = np.empty((s.n_tasks, 3, 2 * self.H, 2 * self.W), np.uint8)
        need = np.empty((s.n_tasks, 4), np.int32)
        _check(lib().ofdg_debug_background(self._h, C.byref(s), _ptr(out), _ptr(need)))
        return out, need

    def debug_composite_luts(self):
        a = np.empty((256, 256), np.uint8)
        s = np.empty((256, 256), np.uint8)
        _check(lib().ofdg_debug_composite_luts(self._h, _ptr(a), _ptr(s)))
        return a, s

    def prepare(self, tasks):
        s = tasks.struct()
        h = C.c_void_p()
        _check(lib().ofdg_prepare(self._h, C.byref(s), C.byref(h)))
        return Prepared(h, s.n_tasks)

    def render_prepared(self, prepared, img0, img1, flow, stream=None):
        _check(lib().ofdg_render_prepared(self._h, prepared._h, img0.data_ptr(), img1.data_ptr(), flow.data_ptr(), stream))

    def render_prepared_host(self, prepared, img0=None, img1=None, flow=None):
        """A prepared batch into HOST blobs (numpy arrays or pinned torch tensors) through the pipelined host-blob path."""
        n = prepared.n
        img0 = np.empty((n, 3, self.H, self.W), np.float32) if img0 is None else img0
        img1 = np.empty((n, 3, self.H, self.W), np.float32) if img1 is None else img1
        flow = np.empty((n, 2, self.H, self.W), np.float32) if flow is None else flow
        p = lambda a: a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        _check(lib().ofdg_render_prepared_host(self._h, prepared._h, p(img0), p(img1), p(flow)))
        return img0, img1, flow

    def generate(self, params, batch, img0, img1, flow, stream=None):
        _check(lib().ofdg_generate(self._h, params._h, batch, img0.data_ptr(), img1.data_ptr(), flow.data_ptr(), stream))

    def generate_host(self, params, batch, img0, img1, flow):
        """Draw + render `batch` samples into HOST blobs (numpy arrays or pinned torch tensors)."""
        p = lambda a: a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        _check(lib().ofdg_generate_host(self._h, params._h, batch, p(img0), p(img1), p(flow)))

    def generate_philox(self, seed, first_sample, batch, img0, img1, flow, augment=False, stream=None):
        """Production mode: parameters drawn and flattened on the device (Philox4x32), then rendered."""
        _check(lib().ofdg_generate_philox(self._h, seed, first_sample, batch, int(bool(augment)), img0.data_ptr(), img1.data_ptr(),
                                          flow.data_ptr(), stream))

    def philox_tasks(self, seed, first_sample, batch, augment=False):
        """The blueprints the device stream draws for these samples, as an ordinary task batch."""
        t = Tasks()
        _check(lib().ofdg_philox_tasks(self._h, seed, first_sample, batch, int(bool(augment)), t._h))
        return t

    def launch_count(self):
        return int(lib().ofdg_launch_count(self._h))

    def kernel_times(self):
        """(prep_ms, render_ms, calls) accumulated since the last call; CUDA events on the launching stream."""
        p, r, n = C.c_double(), C.c_double(), C.c_int32()
        _check(lib().ofdg_kernel_times(self._h, C.byref(p), C.byref(r), C.byref(n)))
        return p.value, r.value, n.value

    def last_render_stats(self):
        """(pairs, prepared_px, source_px) of the batch rendered last (ofdg_last_render_stats)."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().ofdg_last_render_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def last_bin_raster_ms(self):
        """Pair binning / mask rasterisation parts of the last kernel_times() call (in-line runs only, else 0)."""
        return float(lib().ofdg_last_bin_ms(self._h)), float(lib().ofdg_last_raster_ms(self._h))

    def last_shade_ms(self):
        """Part of the render time of the last kernel_times() call spent in the shade kernel (0 with OFDG_RENDER=fused)."""
        return float(lib().ofdg_last_shade_ms(self._h))

    def set_extra_tops(self, flow_bw=None, id0=None, id1=None, occlusion=None):
        """Register device float tensors that every later device-blob call also fills (ofdg_set_extra_tops):
        backward flow (N,2,H,W), index images (N,1,H,W) and the occlusion mask (N,1,H,W). No arguments: off."""
        t = ExtraTopsStruct()
        keep = []
        for name, v in (("flow_bw", flow_bw), ("id0", id0), ("id1", id1), ("occlusion", occlusion)):
            if v is not None:
                assert v.is_cuda and v.dtype.is_floating_point and v.element_size() == 4 and v.is_contiguous()
                setattr(t, name, v.data_ptr())
                keep.append(v)
        self._extra_keep = keep
        _check(lib().ofdg_set_extra_tops(self._h, C.byref(t) if keep else None))

    def last_upload_bytes(self):
        return int(lib().ofdg_last_upload_bytes(self._h))

    def last_download_bytes(self):
        """Bytes the last render_host / generate_host call copied device-to-host (uint8 frames + float flow by default)."""
        return int(lib().ofdg_last_download_bytes(self._h))


class Prepared:
    def __init__(self, h, n):
        self._h, self.n = h, n

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ofdg_prepared_destroy(self._h)
            self._h = None


# ---------------------------------------------------------------------------------------------------
# Procedural texture pool in numpy; bit-identical to synth_textures_kernel (csrc/render.cu).
# ---------------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _hash8(key, gx, gy, o):
    with np.errstate(over="ignore"):
        z = (np.uint64(key) + gx.astype(np.uint64) * np.uint64(0xBF58476D1CE4E5B9)
             + gy.astype(np.uint64) * np.uint64(0x94D049BB133111EB) + np.uint64(o) * np.uint64(0xD6E8FEB86659FD93))
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z & np.uint64(255)).astype(np.uint32)


def synth_textures(n, w=1024, h=768, seed=0, first_index=0):
    """n x 3 x h x w uint8 procedural textures (value noise + stripes + checker, integer only)."""
    out = np.empty((n, 3, h, w), np.uint8)
    x = np.arange(w, dtype=np.uint32)[None, :].repeat(h, 0)
    y = np.arange(h, dtype=np.uint32)[:, None].repeat(w, 1)
    one = np.ones(1, np.uint32)
    for ti in range(n):
        t = ti + first_index
        with np.errstate(over="ignore"):
            tkey = np.uint64(seed) ^ (np.uint64(t + 1) * np.uint64(0x9E3779B97F4A7C15))
        sax = int(_hash8(tkey, one * 1, one * 2, 77)[0]) % 17
        say = int(_hash8(tkey, one * 3, one * 4, 77)[0]) % 17
        phase = (x * np.uint32(sax) + y * np.uint32(say)) & np.uint32(255)
        tri = np.where(phase < 128, phase, np.uint32(255) - phase) * np.uint32(2)
        checker = ((x >> np.uint32(5)) ^ (y >> np.uint32(5))) & np.uint32(1)
        for c in range(3):
            with np.errstate(over="ignore"):
                key = tkey + np.uint64(c + 1) * np.uint64(0xA24BAED4963EE407)
            acc = np.zeros((h, w), np.uint32)
            for o in range(4):
                cell = np.uint32(64 >> o)
                gx, gy = x // cell, y // cell
                fx, fy = (x % cell) * np.uint32(256) // cell, (y % cell) * np.uint32(256) // cell
                v00, v10 = _hash8(key, gx, gy, o), _hash8(key, gx + 1, gy, o)
                v01, v11 = _hash8(key, gx, gy + 1, o), _hash8(key, gx + 1, gy + 1, o)
                top = v00 * (np.uint32(256) - fx) + v10 * fx
                bot = v01 * (np.uint32(256) - fx) + v11 * fx
                val = (top * (np.uint32(256) - fy) + bot * fy) >> np.uint32(16)
                acc += val << np.uint32(3 - o)
            v = (acc // np.uint32(15)) * np.uint32(3) // np.uint32(4) + tri // np.uint32(4) + checker * np.uint32(24) + np.uint32(c * 5)
            out[ti, c] = np.minimum(v, 255).astype(np.uint8)
    return out


# ---------------------------------------------------------------------------------------------------
# Caffe-style layer surface (csrc/host/layer.hpp): prototxt in, three top blobs out.
# ---------------------------------------------------------------------------------------------------
def gather_blobs(blobs, dst=0, group=None, out=None):
    """The optional epilogue of sharded generation (BASELINE north_star): every rank's finished blobs
    (img0, img1, flow -- device tensors under NCCL, host tensors under gloo) are gathered to the training rank `dst`,
    which receives them stacked along the batch axis in rank order: [(world * N, 3, H, W), (world * N, 3, H, W),
    (world * N, 2, H, W)]; the other ranks get None. No collective sits on the generation path itself. `out`, on the
    destination rank, is an optional list of preallocated stacked tensors to receive into."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    res = []
    for i, b in enumerate(blobs):
        if rank == dst:
            stacked = out[i] if out is not None else torch.empty((world * b.shape[0],) + tuple(b.shape[1:]), dtype=b.dtype, device=b.device)
            parts = list(stacked.split(b.shape[0], dim=0))  # views: the gather lands in place
            dist.gather(b, parts, dst=dst, group=group)
            res.append(stacked)
        else:
            dist.gather(b, None, dst=dst, group=group)
    return res if rank == dst else None


def parse_prototxt(text):
    """Fields of a prototxt `layer { ... }` block as the layer sees them."""
    ints = (C.c_int32 * 7)()
    db = C.create_string_buffer(4096)
    ty = C.create_string_buffer(256)
    if lib().ofdg_layer_parse_prototxt(text.encode(), ints, db, 4096, ty, 256):
        raise OfdgError(lib().ofdg_layer_last_error().decode())
    keys = ["batch_size", "prefetch", "mode", "first_level_threads", "second_level_threads", "use_antialiasing", "top_size"]
    d = dict(zip(keys, [int(v) for v in ints]))
    d["use_antialiasing"] = bool(d["use_antialiasing"])
    d["texture_dbases"] = db.value.decode()
    d["type"] = ty.value.decode()
    return d


class DataGenerationLayer:
    """caffe::DataGenerationLayer<float> (type "DataGeneration", 0 bottoms, 3 tops) driven from Python.

    layer = DataGenerationLayer(prototxt_text, texture_db="synthetic:64"); layer.LayerSetUp()
    layer.Forward_gpu(); img0, img1, flow = layer.top_tensors()   # torch views of the device blobs
    """

    def __init__(self, prototxt, texture_db=None, solver_rank=0):
        self._h = C.c_void_p()
        if lib().ofdg_layer_create(prototxt.encode(), (texture_db or "").encode(), solver_rank, C.byref(self._h)):
            raise OfdgError(lib().ofdg_layer_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ofdg_layer_destroy(self._h)
            self._h = None

    __del__ = close

    def type(self):
        return lib().ofdg_layer_type(self._h).decode()

    def producer_stats(self):
        """(ms drawing per batch, ms in ofdg_prepare per batch, batches) of the prefetch threads since the last call."""
        a = (C.c_double * 3)()
        lib().ofdg_layer_producer_stats.argtypes = [C.c_void_p, C.c_void_p]
        if lib().ofdg_layer_producer_stats(self._h, a):
            raise OfdgError(lib().ofdg_layer_last_error().decode())
        n = max(a[2], 1.0)
        return a[0] / n, a[1] / n, int(a[2])

    @staticmethod
    def registered_types():
        """LayerRegistry<float>::LayerTypeList() of the shim (REGISTER_LAYER_CLASS(DataGeneration))."""
        buf = C.create_string_buffer(1024)
        lib().ofdg_layer_registered_types(buf, 1024)
        return [t for t in buf.value.decode().split(",") if t]

    def LayerSetUp(self):
        if lib().ofdg_layer_setup(self._h):
            raise OfdgError(lib().ofdg_layer_last_error().decode())

    def top_shape(self, i):
        s = (C.c_int32 * 4)()
        if lib().ofdg_layer_top_shape(self._h, i, s):
            raise OfdgError(lib().ofdg_layer_last_error().decode())
        return tuple(int(v) for v in s)

    def Forward_gpu(self):
        if lib().ofdg_layer_forward(self._h, 1):
            raise OfdgError(lib().ofdg_layer_last_error().decode())

    def Forward_cpu(self):
        if lib().ofdg_layer_forward(self._h, 0):
            raise OfdgError(lib().ofdg_layer_last_error().decode())

    def top_cpu(self, i):
        shape = self.top_shape(i)
        ptr = lib().ofdg_layer_top_data(self._h, i, 0)
        if not ptr:
            raise OfdgError(lib().ofdg_layer_last_error().decode())
        n = int(np.prod(shape))
        return np.frombuffer((C.c_float * n).from_address(ptr), dtype=np.float32).reshape(shape).copy()
