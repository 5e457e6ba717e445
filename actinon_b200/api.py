"""ctypes binding of libactinon_b200.so (see include/actinon_b200.h).

Host mirror of the reference interface for the hot path:

* ``Scene``            scene_s + the scene-edit API (objects.c:1463-1716, compound.c:380-455,
                       container.c:376-421, scene.c:238-331) — pure host code
* ``Tracer``           the device copy of one flattened scene
* ``lum_machine_run``  ``lum_machine_s_run(scene, lum_arr)`` (scene.c:1017-1028)
* ``Image``            lum_image_s + the pass controller (scene.c:760-885,1032-1165)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libactinon_b200.so"
_lib = None

SEED_POSITION_HASH = 0
SEED_INDEX_KEYED = 1
PRECISION_F32 = 0
PRECISION_F64 = 1
CSG_AUTO, CSG_INTERVALS, CSG_MARCH = 0, 1, 2
SPECIALIZE_AUTO, SPECIALIZE_ON, SPECIALIZE_OFF = 0, 1, 2

(KIND_COMPOUND, KIND_PLANE, KIND_SPHERE, KIND_SQUAROID, KIND_DIST_SPHERE, KIND_DIST_TORUS,
 KIND_PAIR_INSIDE, KIND_PAIR_OUTSIDE, KIND_NEG, KIND_SCALE) = range(10)


class AcnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"actinon_b200 error {code}: {msg}")
        self.code = code


# ----------------------------------------------------------------------------------------------
# C structs (include/actinon_b200.h)
# ----------------------------------------------------------------------------------------------
class FlatNode(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("child0", C.c_int32), ("child1", C.c_int32), ("material", C.c_int32),
        ("has_envelope", C.c_int32), ("reserved", C.c_int32),
        ("env_pos", C.c_double * 3), ("env_radius", C.c_double),
        ("pos", C.c_double * 3), ("rax", C.c_double * 9),
        ("surface_roughness", C.c_double), ("tail", C.c_double * 4),
    ]


class FlatMaterial(C.Structure):
    _fields_ = [
        ("color", C.c_double * 3), ("radiance", C.c_double), ("refractive_index", C.c_double),
        ("fresnel_reflectivity", C.c_double), ("chromatic_reflectivity", C.c_double),
        ("diffuse_reflectivity", C.c_double), ("sigma", C.c_double), ("transparency", C.c_double * 3),
        ("texture_kind", C.c_int32), ("reserved", C.c_int32),
        ("tex_color1", C.c_double * 3), ("tex_color2", C.c_double * 3), ("tex_scale", C.c_double),
    ]


class FlatParams(C.Structure):
    _fields_ = [
        ("image_width", C.c_int32), ("image_height", C.c_int32), ("gamma", C.c_double),
        ("background_color", C.c_double * 3), ("camera_position", C.c_double * 3),
        ("camera_view_direction", C.c_double * 3), ("camera_top_direction", C.c_double * 3),
        ("camera_focal_length", C.c_double),
        ("trace_depth", C.c_int32), ("direct_samples", C.c_int32), ("path_samples", C.c_int32),
        ("gradient_samples", C.c_int32), ("gradient_cycles", C.c_int32), ("threads", C.c_int32),
        ("trace_min_intensity", C.c_double), ("max_path_length", C.c_double), ("gradient_threshold", C.c_double),
    ]


class FlatSceneStruct(C.Structure):
    _fields_ = [
        ("params", FlatParams),
        ("n_nodes", C.c_int32), ("n_children", C.c_int32), ("n_materials", C.c_int32),
        ("light_root", C.c_int32), ("matter_root", C.c_int32), ("reserved", C.c_int32),
        ("nodes", C.POINTER(FlatNode)), ("children", C.POINTER(C.c_int32)), ("materials", C.POINTER(FlatMaterial)),
    ]


class Options(C.Structure):
    _fields_ = [
        ("seed_mode", C.c_int32), ("precision", C.c_int32), ("eps", C.c_double),
        ("wave_budget", C.c_int64), ("device", C.c_int32), ("csg_mode", C.c_int32),
        ("specialize", C.c_int32), ("reserved", C.c_int32),
    ]

    def __init__(self, seed_mode=SEED_POSITION_HASH, precision=PRECISION_F32, eps=0.0, wave_budget=0, device=-1, csg_mode=0,
                 specialize=0):
        super().__init__()
        self.seed_mode, self.precision, self.eps, self.wave_budget, self.device = seed_mode, precision, eps, wave_budget, device
        self.csg_mode = csg_mode
        self.specialize = specialize       # SPECIALIZE_AUTO / _ON / _OFF: kernels compiled for this scene's structure (NVRTC)


class Stats(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64), ("rays_primary", C.c_uint64), ("rays_reflection", C.c_uint64),
        ("rays_chromatic", C.c_uint64), ("rays_refraction", C.c_uint64), ("rays_path", C.c_uint64),
        ("rays_shadow", C.c_uint64), ("rays_light", C.c_uint64), ("diffuse_hits", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("waves", C.c_uint64), ("device_ms", C.c_double),
        ("kernel_ms", C.c_double * 8), ("kernel_launches_by_class", C.c_uint64 * 8),
    ]

    @property
    def rays(self) -> int:
        return (self.rays_primary + self.rays_reflection + self.rays_chromatic + self.rays_refraction
                + self.rays_path + self.rays_shadow)

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("kernel_")}
        d["kernel_ms"] = list(self.kernel_ms)
        d["kernel_launches_by_class"] = list(self.kernel_launches_by_class)
        d["rays"] = self.rays
        return d


# every symbol include/actinon_b200.h declares
EXPORTED_SYMBOLS = [
    "acn_options_default", "acn_tracer_create", "acn_tracer_destroy", "acn_render_samples",
    "acn_render_samples_device", "acn_accumulate_device", "acn_last_error", "acn_version", "acn_device_count",
    "acn_measure_fp32_peak_tflops", "acn_tracer_stream", "acn_spec_probe",
    "acn_scene_create", "acn_scene_destroy", "acn_scene_params",
    "acn_create_plane", "acn_create_sphere", "acn_create_squaroid", "acn_create_ellipsoid", "acn_create_cylinder",
    "acn_create_cone", "acn_create_hyperboloid1", "acn_create_hyperboloid2", "acn_create_torus", "acn_create_distance_sphere",
    "acn_clone", "acn_pair_inside", "acn_pair_outside", "acn_neg", "acn_scale_object",
    "acn_list_create", "acn_list_push", "acn_list_inside_composite", "acn_list_outside_composite", "acn_list_create_compound",
    "acn_move", "acn_rotate", "acn_scale", "acn_set_color", "acn_set_transparency", "acn_set_refractive_index",
    "acn_set_radiance", "acn_set_fresnel_reflectivity", "acn_set_chromatic_reflectivity", "acn_set_diffuse_reflectivity",
    "acn_set_sigma", "acn_set_surface_roughness", "acn_set_material", "acn_set_envelope", "acn_set_auto_envelope", "acn_set_bounding_envelope",
    "acn_set_texture_plain", "acn_set_texture_chess", "acn_scene_clear", "acn_scene_push",
    "acn_scene_load_acn", "acn_scene_image_name", "acn_scene_select_image", "acn_scene_flatten",
    "acn_image_create", "acn_image_destroy", "acn_image_size", "acn_image_cycle", "acn_image_rval", "acn_image_next_pass",
    "acn_image_push", "acn_image_average", "acn_image_sums", "acn_image_add_sums", "acn_image_write_pnm",
    "acn_image_save", "acn_image_load", "acn_image_set_state",
    "acn_pixel_owner", "acn_dimage_create", "acn_dimage_destroy", "acn_dimage_reset", "acn_dimage_set_shard", "acn_dimage_cycle", "acn_dimage_rval", "acn_dimage_stream",
    "acn_dimage_begin_pass", "acn_dimage_accumulate", "acn_dimage_delta", "acn_dimage_end_pass", "acn_dimage_render_pass",
    "acn_dimage_download", "acn_dimage_upload", "acn_dimage_copy_delta", "acn_dimage_set_delta", "acn_dimage_read_pass_xy",
    "acn_group_create", "acn_group_destroy", "acn_group_size", "acn_group_uses_peer_access", "acn_group_render_pass",
    "acn_group_download", "acn_group_upload", "acn_group_image",
]


def library_path() -> str:
    # ACN_B200_LIBRARY: development hook for A/B-ing kernel build variants (tools/exp_build.sh); always a CUDA build
    return os.environ.get("ACN_B200_LIBRARY") or os.path.join(_HERE, _LIB_NAME)


def load_library():
    """Loads the in-tree CUDA extension.  Fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise AcnError(-3, f"{path} is missing: run ./build.sh (or __graft_entry__.build()); there is no fallback path")
    lib = C.CDLL(path)
    D, I, V, P = C.c_double, C.c_int, C.c_void_p, C.POINTER
    pd = P(C.c_double)
    sig = {
        "acn_last_error": (C.c_char_p, []), "acn_version": (C.c_char_p, []), "acn_device_count": (I, []),
        "acn_options_default": (None, [P(Options)]),
        "acn_tracer_create": (I, [P(FlatSceneStruct), P(Options), P(V)]),
        "acn_tracer_destroy": (None, [V]),
        "acn_render_samples": (I, [V, V, C.c_uint64, C.c_uint64, V, V, P(Stats)]),
        "acn_render_samples_device": (I, [V, V, C.c_uint64, C.c_uint64, V, V, V, P(Stats)]),
        "acn_accumulate_device": (I, [V, V, V, C.c_uint64, V, V]),
        "acn_measure_fp32_peak_tflops": (D, [I]),
        "acn_tracer_stream": (V, [V]),
        "acn_spec_probe": (I, [P(FlatSceneStruct), P(Options), I, C.c_char_p, C.c_uint64, P(C.c_uint64), P(D)]),
        "acn_scene_create": (I, [P(V)]), "acn_scene_destroy": (None, [V]), "acn_scene_params": (P(FlatParams), [V]),
        "acn_create_plane": (I, [V]), "acn_create_sphere": (I, [V, D]), "acn_create_squaroid": (I, [V, D, D, D, D]),
        "acn_create_ellipsoid": (I, [V, D, D, D]), "acn_create_cylinder": (I, [V, D, D]), "acn_create_cone": (I, [V, D, D, D]),
        "acn_create_hyperboloid1": (I, [V, D, D, D]), "acn_create_hyperboloid2": (I, [V, D, D, D]),
        "acn_create_torus": (I, [V, D, D]), "acn_create_distance_sphere": (I, [V]),
        "acn_clone": (I, [V, I]), "acn_pair_inside": (I, [V, I, I]), "acn_pair_outside": (I, [V, I, I]), "acn_neg": (I, [V, I]),
        "acn_scale_object": (I, [V, I, pd]),
        "acn_list_create": (I, [V]), "acn_list_push": (I, [V, I, I]), "acn_list_inside_composite": (I, [V, I]),
        "acn_list_outside_composite": (I, [V, I]), "acn_list_create_compound": (I, [V, I]),
        "acn_move": (I, [V, I, pd]), "acn_rotate": (I, [V, I, pd]), "acn_scale": (I, [V, I, D]),
        "acn_set_color": (I, [V, I, pd]), "acn_set_transparency": (I, [V, I, pd]),
        "acn_set_refractive_index": (I, [V, I, D]), "acn_set_radiance": (I, [V, I, D]),
        "acn_set_fresnel_reflectivity": (I, [V, I, D]), "acn_set_chromatic_reflectivity": (I, [V, I, D]),
        "acn_set_diffuse_reflectivity": (I, [V, I, D]), "acn_set_sigma": (I, [V, I, D]), "acn_set_surface_roughness": (I, [V, I, D]),
        "acn_set_material": (I, [V, I, C.c_char_p]), "acn_set_envelope": (I, [V, I, pd, D]), "acn_set_auto_envelope": (I, [V, I]), "acn_set_bounding_envelope": (I, [V, I]),
        "acn_set_texture_plain": (I, [V, I, pd]), "acn_set_texture_chess": (I, [V, I, pd, pd, D]),
        "acn_scene_clear": (I, [V]), "acn_scene_push": (I, [V, I]),
        "acn_scene_load_acn": (I, [V, C.c_char_p, I, P(C.c_char_p), P(I)]),
        "acn_scene_image_name": (C.c_char_p, [V, I]), "acn_scene_select_image": (I, [V, I]),
        "acn_scene_flatten": (I, [V, P(P(FlatSceneStruct))]),
        "acn_image_create": (I, [C.c_int32, C.c_int32, P(V)]), "acn_image_destroy": (None, [V]),
        "acn_image_size": (I, [V, P(C.c_int32), P(C.c_int32)]),
        "acn_image_cycle": (C.c_int32, [V]), "acn_image_rval": (C.c_uint64, [V]),
        "acn_image_next_pass": (I, [V, P(FlatParams), P(pd), P(C.c_uint64)]),
        "acn_image_push": (I, [V, V, V, C.c_uint64]), "acn_image_average": (I, [V, V]),
        "acn_image_sums": (I, [V, V]), "acn_image_add_sums": (I, [V, V]),
        "acn_image_write_pnm": (I, [V, C.c_char_p, P(C.c_uint64)]),
        "acn_image_save": (I, [V, C.c_char_p]), "acn_image_load": (I, [C.c_char_p, P(V)]),
        "acn_image_set_state": (I, [V, C.c_int32, C.c_uint64, V]),
        "acn_dimage_create": (I, [I, C.c_int32, C.c_int32, P(V)]), "acn_dimage_destroy": (None, [V]),
        "acn_dimage_set_shard": (I, [V, C.c_int32, C.c_int32, C.c_int32]), "acn_dimage_reset": (I, [V]),
        "acn_pixel_owner": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
        "acn_dimage_cycle": (C.c_int32, [V]), "acn_dimage_rval": (C.c_uint64, [V]), "acn_dimage_stream": (V, [V]),
        "acn_dimage_begin_pass": (I, [V, P(FlatParams), P(V), P(C.c_uint64), P(C.c_uint64)]),
        "acn_dimage_accumulate": (I, [V, V, V, C.c_uint64, V]),
        "acn_dimage_delta": (V, [V, P(C.c_uint64)]), "acn_dimage_end_pass": (I, [V, V]),
        "acn_dimage_render_pass": (I, [V, V, P(FlatParams), C.c_uint64, P(C.c_uint64), P(C.c_uint64), V, P(Stats)]),
        "acn_dimage_download": (I, [V, V]), "acn_dimage_upload": (I, [V, V]),
        "acn_dimage_copy_delta": (I, [V, V, V]), "acn_dimage_set_delta": (I, [V, V, V]),
        "acn_dimage_read_pass_xy": (I, [V, V]),
        "acn_group_create": (I, [P(FlatSceneStruct), P(Options), P(C.c_int32), C.c_int32, P(V)]), "acn_group_destroy": (None, [V]),
        "acn_group_size": (I, [V]), "acn_group_uses_peer_access": (I, [V]),
        "acn_group_render_pass": (I, [V, C.c_uint64, P(C.c_uint64), V, P(Stats)]),
        "acn_group_download": (I, [V, V]), "acn_group_upload": (I, [V, V]), "acn_group_image": (V, [V, C.c_int32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int):
    if rc < 0:
        raise AcnError(rc, load_library().acn_last_error().decode("utf-8", "replace"))
    return rc


def spec_probe(flat: "FlatScene", options: Optional["Options"] = None, compile: bool = False):
    """The source the run-time specialiser writes for a scene's structure ('' if the scene does not qualify) and, with
    compile=True, the NVRTC compile time in seconds.  Needs no GPU (NVRTC cross-compiles for sm_100a)."""
    lib = load_library()
    options = options or Options()
    cap = 1 << 22
    buf = C.create_string_buffer(cap)
    n = C.c_uint64(0)
    sec = C.c_double(0)
    _check(lib.acn_spec_probe(flat.ptr, C.byref(options), 1 if compile else 0, buf, cap, C.byref(n), C.byref(sec)))
    return buf.raw[: n.value].decode(), sec.value


def device_count() -> int:
    """Number of CUDA devices; raises AcnError(-3) when there is none (no CPU fallback)."""
    return _check(load_library().acn_device_count())


def measure_fp32_peak_tflops(device: int = -1) -> float:
    v = load_library().acn_measure_fp32_peak_tflops(device)
    if v < 0:
        raise AcnError(int(v), load_library().acn_last_error().decode())
    return v


def _vec3(v) -> C.Array:
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


def rotx(deg: float):
    """closures.c:115-139: rotations take degrees; rows as in vectors.h:289-307"""
    a = np.deg2rad(deg); s, c = np.sin(a), np.cos(a)
    return [1, 0, 0, 0, c, -s, 0, s, c]


def roty(deg: float):
    a = np.deg2rad(deg); s, c = np.sin(a), np.cos(a)
    return [c, 0, s, 0, 1, 0, -s, 0, c]


def rotz(deg: float):
    a = np.deg2rad(deg); s, c = np.sin(a), np.cos(a)
    return [c, -s, 0, s, c, 0, 0, 0, 1]


# ----------------------------------------------------------------------------------------------
# Scene: host-side scene description (value semantics like the .acn language)
# ----------------------------------------------------------------------------------------------
class Obj:
    """Handle of a host-side value (shape, compound or list) owned by a Scene."""

    def __init__(self, scene: "Scene", handle: int):
        self.scene, self.h = scene, _check(handle)

    # --- combinators: clone their operands (objects.c:1011-1018,1161-1176,1315-1321,1388-1407)
    def __and__(self, other: "Obj") -> "Obj":
        return Obj(self.scene, self.scene._l.acn_pair_inside(self.scene._p, self.h, other.h))

    def __or__(self, other: "Obj") -> "Obj":
        return Obj(self.scene, self.scene._l.acn_pair_outside(self.scene._p, self.h, other.h))

    def __invert__(self) -> "Obj":
        return Obj(self.scene, self.scene._l.acn_neg(self.scene._p, self.h))

    def clone(self) -> "Obj":
        return Obj(self.scene, self.scene._l.acn_clone(self.scene._p, self.h))

    def scaled_by_vec(self, v) -> "Obj":
        return Obj(self.scene, self.scene._l.acn_scale_object(self.scene._p, self.h, _vec3(v)))

    def __add__(self, v) -> "Obj":          # obj + vec: moved clone (interpreter.c:818-924)
        o = self.clone(); o.move(v); return o

    def __sub__(self, v) -> "Obj":
        o = self.clone(); o.move([-v[0], -v[1], -v[2]]); return o

    def __mul__(self, f) -> "Obj":          # obj * num: scaled clone; obj * rot: rotated clone (interpreter.c:651-785)
        o = self.clone()
        if np.isscalar(f):
            o.scale(float(f))
        else:
            o.rotate(f)
        return o

    # --- in-place edits
    def move(self, v):
        _check(self.scene._l.acn_move(self.scene._p, self.h, _vec3(v))); return self

    def rotate(self, m):
        _check(self.scene._l.acn_rotate(self.scene._p, self.h, (C.c_double * 9)(*[float(x) for x in m]))); return self

    def scale(self, f: float):
        _check(self.scene._l.acn_scale(self.scene._p, self.h, float(f))); return self

    def set_color(self, rgb):
        _check(self.scene._l.acn_set_color(self.scene._p, self.h, _vec3(rgb))); return self

    def set_transparency(self, rgb):
        _check(self.scene._l.acn_set_transparency(self.scene._p, self.h, _vec3(rgb))); return self

    def set_material(self, name: str):
        _check(self.scene._l.acn_set_material(self.scene._p, self.h, name.encode())); return self

    def set_envelope(self, pos, radius: float):
        _check(self.scene._l.acn_set_envelope(self.scene._p, self.h, _vec3(pos), float(radius))); return self

    def set_bounding_envelope(self):
        """Analytic conservative bounding sphere (raises AcnError -5 for unbounded shapes)."""
        _check(self.scene._l.acn_set_bounding_envelope(self.scene._p, self.h)); return self

    def set_auto_envelope(self):
        _check(self.scene._l.acn_set_auto_envelope(self.scene._p, self.h)); return self

    def set_texture_plain(self, rgb):
        _check(self.scene._l.acn_set_texture_plain(self.scene._p, self.h, _vec3(rgb))); return self

    def set_texture_chess(self, rgb1, rgb2, scale: float):
        _check(self.scene._l.acn_set_texture_chess(self.scene._p, self.h, _vec3(rgb1), _vec3(rgb2), float(scale))); return self

    def push(self, item: "Obj"):
        _check(self.scene._l.acn_list_push(self.scene._p, self.h, item.h)); return self

    def create_inside_composite(self) -> "Obj":
        return Obj(self.scene, self.scene._l.acn_list_inside_composite(self.scene._p, self.h))

    def create_outside_composite(self) -> "Obj":
        return Obj(self.scene, self.scene._l.acn_list_outside_composite(self.scene._p, self.h))

    def create_compound(self) -> "Obj":
        return Obj(self.scene, self.scene._l.acn_list_create_compound(self.scene._p, self.h))


def _scalar_setter(cname):
    def f(self, v: float):
        _check(getattr(self.scene._l, cname)(self.scene._p, self.h, float(v))); return self
    return f


for _n in ("refractive_index", "radiance", "fresnel_reflectivity", "chromatic_reflectivity", "diffuse_reflectivity",
           "sigma", "surface_roughness"):
    setattr(Obj, "set_" + _n, _scalar_setter("acn_set_" + _n))


class FlatScene:
    """Borrowed pointer to a flattened scene (owned by its Scene until the next flatten())."""

    def __init__(self, ptr, owner):
        self.ptr, self._owner = ptr, owner

    @property
    def struct(self) -> FlatSceneStruct:
        return self.ptr.contents

    @property
    def params(self) -> FlatParams:
        return self.ptr.contents.params


class Scene:
    """scene_s (scene.c:153-213): render parameters + light and matter compounds."""

    def __init__(self):
        self._l = load_library()
        p = C.c_void_p()
        _check(self._l.acn_scene_create(C.byref(p)))
        self._p = p

    def __del__(self):
        try:
            if getattr(self, "_p", None):
                self._l.acn_scene_destroy(self._p); self._p = None
        except Exception:
            pass

    @property
    def params(self) -> FlatParams:
        return self._l.acn_scene_params(self._p).contents

    def set(self, **kw):
        prm = self.params
        for k, v in kw.items():
            cur = getattr(prm, k)
            if hasattr(cur, "__len__"):
                for i in range(len(cur)):
                    cur[i] = float(v[i])
            else:
                setattr(prm, k, v)
        return self

    # constructors (closures.c:417-591)
    def create_plane(self): return Obj(self, self._l.acn_create_plane(self._p))
    def create_sphere(self, r): return Obj(self, self._l.acn_create_sphere(self._p, float(r)))
    def create_squaroid(self, a, b, c, r): return Obj(self, self._l.acn_create_squaroid(self._p, float(a), float(b), float(c), float(r)))
    def create_ellipsoid(self, rx, ry, rz): return Obj(self, self._l.acn_create_ellipsoid(self._p, float(rx), float(ry), float(rz)))
    def create_cylinder(self, rx, ry): return Obj(self, self._l.acn_create_cylinder(self._p, float(rx), float(ry)))
    def create_cone(self, rx, ry, rz): return Obj(self, self._l.acn_create_cone(self._p, float(rx), float(ry), float(rz)))
    def create_hyperboloid1(self, rx, ry, rz): return Obj(self, self._l.acn_create_hyperboloid1(self._p, float(rx), float(ry), float(rz)))
    def create_hyperboloid2(self, rx, ry, rz): return Obj(self, self._l.acn_create_hyperboloid2(self._p, float(rx), float(ry), float(rz)))
    def create_torus(self, r1, r2): return Obj(self, self._l.acn_create_torus(self._p, float(r1), float(r2)))
    def create_distance_sphere(self): return Obj(self, self._l.acn_create_distance_sphere(self._p))
    def create_list(self, items: Sequence[Obj] = ()):
        l = Obj(self, self._l.acn_list_create(self._p))
        for it in items:
            l.push(it)
        return l

    def clear(self):
        _check(self._l.acn_scene_clear(self._p)); return self

    def push(self, o: Obj):
        _check(self._l.acn_scene_push(self._p, o.h)); return self

    def load_acn(self, path: str, args: Sequence[str] = ()) -> int:
        """Evaluates an .acn script; returns the number of recorded create_image calls."""
        n = C.c_int(0)
        arr = (C.c_char_p * max(1, len(args)))(*[a.encode() for a in args])
        _check(self._l.acn_scene_load_acn(self._p, path.encode(), len(args), arr, C.byref(n)))
        return n.value

    def image_name(self, i: int) -> str:
        s = self._l.acn_scene_image_name(self._p, i)
        return s.decode() if s else ""

    def select_image(self, i: int):
        _check(self._l.acn_scene_select_image(self._p, i)); return self

    def flatten(self) -> FlatScene:
        out = C.POINTER(FlatSceneStruct)()
        _check(self._l.acn_scene_flatten(self._p, C.byref(out)))
        return FlatScene(out, self)


# ----------------------------------------------------------------------------------------------
# Tracer
# ----------------------------------------------------------------------------------------------
class Tracer:
    """Device copy of one flattened scene + wavefront queues (acn_tracer)."""

    def __init__(self, flat: FlatScene, options: Optional[Options] = None):
        self._l = load_library()
        self._flat = flat
        self.options = options or Options()
        p = C.c_void_p()
        _check(self._l.acn_tracer_create(flat.ptr, C.byref(self.options), C.byref(p)))
        self._p = p
        self.width, self.height = flat.params.image_width, flat.params.image_height
        self.last_stats = Stats()

    def close(self):
        if getattr(self, "_p", None):
            self._l.acn_tracer_destroy(self._p); self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render_samples(self, xy: np.ndarray, index_base: int = 0) -> np.ndarray:
        """lum_machine_s_run with host buffers: xy float64 [n,2] -> rgb float32 [n,3] (gamma + clamp applied)."""
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        n = xy.shape[0]
        rgb = np.empty((n, 3), dtype=np.float32)
        st = Stats()
        _check(self._l.acn_render_samples(self._p, xy.ctypes.data, n, index_base, rgb.ctypes.data, None, C.byref(st)))
        self.last_stats = st
        return rgb

    def render_samples_device(self, d_xy, d_rgb=None, index_base: int = 0, stream=None):
        """Same with torch CUDA tensors: d_xy float64 [n,2] -> float32 [n,3]."""
        import torch
        assert d_xy.is_cuda and d_xy.dtype == torch.float64 and d_xy.is_contiguous()
        n = d_xy.shape[0]
        if d_rgb is None:
            d_rgb = torch.empty((n, 3), dtype=torch.float32, device=d_xy.device)
        st = Stats()
        s = stream.cuda_stream if stream is not None else torch.cuda.current_stream(d_xy.device).cuda_stream
        _check(self._l.acn_render_samples_device(self._p, d_xy.data_ptr(), n, index_base, d_rgb.data_ptr(), s, None, C.byref(st)))
        self.last_stats = st
        return d_rgb

    def accumulate_device(self, d_xy, d_rgb, d_accum, stream=None):
        """lum_image_s_push_arr on the device: d_accum float32 [h,w,4] += samples."""
        import torch
        s = stream.cuda_stream if stream is not None else torch.cuda.current_stream(d_xy.device).cuda_stream
        _check(self._l.acn_accumulate_device(self._p, d_xy.data_ptr(), d_rgb.data_ptr(), d_xy.shape[0], d_accum.data_ptr(), s))
        return d_accum


def lum_machine_run(tracer: Tracer, xy: np.ndarray, index_base: int = 0) -> np.ndarray:
    """``lum_machine_s_run(scene, lum_arr)`` (scene.c:1017-1028): colours of the given sample positions."""
    return tracer.render_samples(xy, index_base)


# ----------------------------------------------------------------------------------------------
# Image: lum_image_s + pass controller
# ----------------------------------------------------------------------------------------------
class Image:
    def __init__(self, width: int, height: int, _ptr=None):
        self._l = load_library()
        if _ptr is None:
            p = C.c_void_p()
            _check(self._l.acn_image_create(width, height, C.byref(p)))
            _ptr = p
        self._p = _ptr
        self.width, self.height = width, height

    def __del__(self):
        try:
            if getattr(self, "_p", None):
                self._l.acn_image_destroy(self._p); self._p = None
        except Exception:
            pass

    @property
    def cycle(self) -> int:
        return self._l.acn_image_cycle(self._p)

    @property
    def rval(self) -> int:
        return self._l.acn_image_rval(self._p)

    def next_pass(self, params: FlatParams) -> np.ndarray:
        """Sample positions of the next pass (scene.c:1108-1139); empty when all passes are done."""
        xy = C.POINTER(C.c_double)()
        n = C.c_uint64(0)
        _check(self._l.acn_image_next_pass(self._p, C.byref(params), C.byref(xy), C.byref(n)))
        if n.value == 0:
            return np.empty((0, 2), dtype=np.float64)
        return np.ctypeslib.as_array(xy, shape=(n.value, 2)).copy()

    def push(self, xy: np.ndarray, rgb: np.ndarray):
        xy = np.ascontiguousarray(xy, dtype=np.float64); rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        _check(self._l.acn_image_push(self._p, xy.ctypes.data, rgb.ctypes.data, xy.shape[0]))

    def average(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 3), dtype=np.float32)
        _check(self._l.acn_image_average(self._p, out.ctypes.data))
        return out

    def sums(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 6), dtype=np.float64)
        _check(self._l.acn_image_sums(self._p, out.ctypes.data))
        return out

    def add_sums(self, sums: np.ndarray):
        sums = np.ascontiguousarray(sums, dtype=np.float64)
        _check(self._l.acn_image_add_sums(self._p, sums.ctypes.data))

    def write_pnm(self, path: Optional[str]) -> int:
        h = C.c_uint64(0)
        _check(self._l.acn_image_write_pnm(self._p, path.encode() if path else None, C.byref(h)))
        return h.value

    def save(self, path: str):
        _check(self._l.acn_image_save(self._p, path.encode()))

    @staticmethod
    def load(path: str) -> "Image":
        lib = load_library()
        p = C.c_void_p()
        _check(lib.acn_image_load(path.encode(), C.byref(p)))
        im = Image.__new__(Image)
        im._l, im._p = lib, p
        w, h = C.c_int32(0), C.c_int32(0)
        _check(lib.acn_image_size(p, C.byref(w), C.byref(h)))
        im.width, im.height = w.value, h.value
        return im


def pixel_owner(x: int, y: int, tile: int, n_ranks: int) -> int:
    """Rank that owns pixel (x, y): tile x tile squares dealt along a Morton curve (acn_dimage_set_shard)."""
    return load_library().acn_pixel_owner(x, y, tile, n_ranks)


PIX_SCALE = float(1 << 44)      # fixed-point scale of the device image's sums (Q20.44)


class DeviceImage:
    """lum_image_s + the pass controller resident on the device (acn_dimage): gradient selection, sample list and
    accumulation run as kernels; sums are 64-bit fixed point (bit-identical images for any number of GPUs)."""

    def __init__(self, width: int, height: int, device: int = -1):
        self._l = load_library()
        p = C.c_void_p()
        _check(self._l.acn_dimage_create(device, width, height, C.byref(p)))
        self._p = p
        self.width, self.height = width, height
        self.n_ranks = 1

    def close(self):
        if getattr(self, "_p", None):
            self._l.acn_dimage_destroy(self._p); self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _check(self._l.acn_dimage_reset(self._p))

    def accumulate(self, d_xy, d_rgb, stream=None):
        """lum_image_s_push_arr of device-resident samples into the pass delta (torch CUDA tensors)."""
        s = stream.cuda_stream if stream is not None else self._l.acn_dimage_stream(self._p)
        _check(self._l.acn_dimage_accumulate(self._p, d_xy.data_ptr(), d_rgb.data_ptr(), d_xy.shape[0], s))

    def set_shard(self, n_ranks: int, rank: int, tile: int = 4):
        _check(self._l.acn_dimage_set_shard(self._p, n_ranks, rank, tile))
        self.n_ranks = n_ranks

    @property
    def cycle(self) -> int:
        return self._l.acn_dimage_cycle(self._p)

    @property
    def rval(self) -> int:
        return self._l.acn_dimage_rval(self._p)

    @property
    def words(self) -> int:
        return self.width * self.height * 6

    def render_pass(self, tracer: "Tracer", params: FlatParams, index_base: int = 0):
        """One pass (scene.c:1108-1159): builds this rank's sample list on the device, traces it, accumulates.  Returns
        (n_local, n_total); (0, 0) when all passes are done.  With several ranks the caller then sums the deltas over the
        ranks (copy_delta / all_reduce / set_delta) and calls end_pass()."""
        nl, nt = C.c_uint64(0), C.c_uint64(0)
        st = Stats()
        _check(self._l.acn_dimage_render_pass(self._p, tracer._p, C.byref(params), index_base, C.byref(nl), C.byref(nt), None, C.byref(st)))
        tracer.last_stats = st
        return nl.value, nt.value

    def pass_xy(self, n_local: int) -> np.ndarray:
        xy = np.empty((n_local, 2), dtype=np.float64)
        _check(self._l.acn_dimage_read_pass_xy(self._p, xy.ctypes.data))
        return xy

    def copy_delta(self, d_tensor, stream=None):
        """The pass delta -> a caller-owned int64 CUDA tensor (stream: a torch stream; default the image's own stream)."""
        s = stream.cuda_stream if stream is not None else self._l.acn_dimage_stream(self._p)
        _check(self._l.acn_dimage_copy_delta(self._p, d_tensor.data_ptr(), s))

    def set_delta(self, d_tensor, stream=None):
        s = stream.cuda_stream if stream is not None else self._l.acn_dimage_stream(self._p)
        _check(self._l.acn_dimage_set_delta(self._p, d_tensor.data_ptr(), s))

    def end_pass(self):
        _check(self._l.acn_dimage_end_pass(self._p, self._l.acn_dimage_stream(self._p)))

    def download(self, image: Optional["Image"] = None) -> "Image":
        image = image or Image(self.width, self.height)
        _check(self._l.acn_dimage_download(self._p, image._p))
        return image

    def upload(self, image: "Image"):
        _check(self._l.acn_dimage_upload(self._p, image._p))


class Group:
    """One image on several GPUs of one box inside one process (acn_group): a thread, a tracer and a device image per GPU,
    pixel tiles dealt to the GPUs, per-pass sums exchanged through peer memory."""

    def __init__(self, flat: "FlatScene", devices: Sequence[int], options: Optional[Options] = None):
        self._l = load_library()
        self._flat = flat
        self.options = options or Options()
        dv = (C.c_int32 * len(devices))(*devices)
        p = C.c_void_p()
        _check(self._l.acn_group_create(flat.ptr, C.byref(self.options), dv, len(devices), C.byref(p)))
        self._p = p
        self.width, self.height = flat.params.image_width, flat.params.image_height
        self.last_stats = Stats()

    def close(self):
        if getattr(self, "_p", None):
            self._l.acn_group_destroy(self._p); self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def uses_peer_access(self) -> bool:
        return bool(self._l.acn_group_uses_peer_access(self._p))

    def render_pass(self, index_base: int = 0) -> int:
        n = C.c_uint64(0)
        st = Stats()
        _check(self._l.acn_group_render_pass(self._p, index_base, C.byref(n), None, C.byref(st)))
        self.last_stats = st
        return n.value

    def render(self, passes: Optional[int] = None):
        """All passes (or the first `passes`); returns (samples, passes, rays)."""
        n_samples = n_pass = rays = 0
        while passes is None or n_pass < passes:
            n = self.render_pass(n_samples)
            if n == 0:
                break
            n_samples += n; n_pass += 1; rays += self.last_stats.rays
        return n_samples, n_pass, rays

    def download(self, image: Optional["Image"] = None) -> "Image":
        image = image or Image(self.width, self.height)
        _check(self._l.acn_group_download(self._p, image._p))
        return image


def render_image_device(flat: "FlatScene", tracer: "Tracer", passes: Optional[int] = None, dist=None, log=None):
    """scene_s_create_image_file with the image on the device; `dist` = torch.distributed (initialised, NCCL) shards the pixel
    tiles over the ranks and all-reduces the per-pass deltas.  Returns (Image, samples, passes, rays of this rank).
    log: optional callable(pass index, n_total, n_local, Stats, wall seconds of the pass)."""
    import time
    prm = flat.params
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    dimg = DeviceImage(prm.image_width, prm.image_height, tracer.options.device)
    buf = None
    if world > 1:
        import torch
        dimg.set_shard(world, rank, 4)
        dev = torch.device("cuda", tracer.options.device if tracer.options.device >= 0 else torch.cuda.current_device())
        buf = torch.zeros(dimg.words, dtype=torch.int64, device=dev)
    n_samples = n_pass = rays = 0
    while passes is None or n_pass < passes:
        t0 = time.perf_counter()
        nl, nt = dimg.render_pass(tracer, prm, n_samples)
        if nt == 0:
            break
        rays += tracer.last_stats.rays if nl else 0
        if world > 1:
            import torch
            dimg.copy_delta(buf)
            torch.cuda.synchronize()            # the copy ran on the image's stream, the all-reduce runs on torch's
            dist.all_reduce(buf)                # disjoint pixel tiles, integer sums: exact, order-independent
            torch.cuda.synchronize()
            dimg.set_delta(buf)
            dimg.end_pass()
        if log is not None:
            log(n_pass, nt, nl, tracer.last_stats, time.perf_counter() - t0)
        n_samples += nt
        n_pass += 1
    img = dimg.download()
    dimg.close()
    return img, n_samples, n_pass, rays


def render_image(scene: Scene, tracer: Optional[Tracer] = None, passes: Optional[int] = None,
                 pnm_path: Optional[str] = None, options: Optional[Options] = None, verbose: bool = False):
    """scene_s_create_image_file (scene.c:1032-1165): all passes, accumulation, optional .pnm after every pass."""
    flat = scene.flatten()
    own = tracer is None
    if tracer is None:
        tracer = Tracer(flat, options)
    prm = flat.params
    img = Image(prm.image_width, prm.image_height)
    total = Stats()
    index_base = 0
    n_pass = 0
    while True:
        if passes is not None and n_pass >= passes:
            break
        xy = img.next_pass(prm)
        if xy.shape[0] == 0:
            break
        rgb = tracer.render_samples(xy, index_base)
        index_base += xy.shape[0]
        img.push(xy, rgb)
        st = tracer.last_stats
        for k, _ in Stats._fields_:
            if not k.startswith("kernel_"):
                setattr(total, k, getattr(total, k) + getattr(st, k))
        if pnm_path:
            h = img.write_pnm(pnm_path)
            if verbose:
                print(f"pass {n_pass}: {xy.shape[0]} samples, {st.device_ms:.1f} ms, hash {h:016x}")
        n_pass += 1
    if own:
        tracer.close()
    return img, total


# ----------------------------------------------------------------------------------------------
# flat-scene files: the flattened form of a scripted scene as a compressed .npz (own format)
# ----------------------------------------------------------------------------------------------
class _OwnedFlat:
    """Keeps the ctypes buffers of a FlatScene loaded from disk alive."""


def save_flat(flat: FlatScene, path: str, name: str = ""):
    fs = flat.struct
    nodes = np.frombuffer(C.string_at(fs.nodes, C.sizeof(FlatNode) * fs.n_nodes), dtype=np.uint8)
    mats = np.frombuffer(C.string_at(fs.materials, C.sizeof(FlatMaterial) * fs.n_materials), dtype=np.uint8)
    children = np.ctypeslib.as_array(fs.children, shape=(max(fs.n_children, 1),))[: fs.n_children].copy()
    params = np.frombuffer(C.string_at(C.byref(fs.params), C.sizeof(FlatParams)), dtype=np.uint8)
    np.savez_compressed(path, nodes=nodes, materials=mats, children=children.astype(np.int32), params=params,
                        roots=np.array([fs.light_root, fs.matter_root], dtype=np.int32),
                        sizes=np.array([C.sizeof(FlatNode), C.sizeof(FlatMaterial), C.sizeof(FlatParams)], dtype=np.int32),
                        name=np.array(name))


def load_flat(path: str) -> FlatScene:
    z = np.load(path)
    assert list(z["sizes"]) == [C.sizeof(FlatNode), C.sizeof(FlatMaterial), C.sizeof(FlatParams)], "flat-scene ABI changed"
    own = _OwnedFlat()
    n_nodes = z["nodes"].size // C.sizeof(FlatNode)
    n_mats = z["materials"].size // C.sizeof(FlatMaterial)
    own.nodes = (FlatNode * n_nodes).from_buffer_copy(z["nodes"].tobytes())
    own.mats = (FlatMaterial * max(n_mats, 1)).from_buffer_copy(z["materials"].tobytes().ljust(C.sizeof(FlatMaterial), b"\0"))
    ch = z["children"].astype(np.int32)
    own.children = (C.c_int32 * max(len(ch), 1))(*ch.tolist())
    own.fs = FlatSceneStruct()
    C.memmove(C.byref(own.fs.params), z["params"].tobytes(), C.sizeof(FlatParams))
    own.fs.n_nodes, own.fs.n_children, own.fs.n_materials = n_nodes, len(ch), n_mats
    own.fs.light_root, own.fs.matter_root = int(z["roots"][0]), int(z["roots"][1])
    own.fs.nodes = C.cast(own.nodes, C.POINTER(FlatNode))
    own.fs.children = C.cast(own.children, C.POINTER(C.c_int32))
    own.fs.materials = C.cast(own.mats, C.POINTER(FlatMaterial))
    return FlatScene(C.pointer(own.fs), own)
