"""actinon_b200 — B200-native (sm_100a) implementation of Actinon's per-sample ray/path-tracing loop.

The product is the C-ABI shared library ``libactinon_b200.so`` (``include/actinon_b200.h``): a
hand-written CUDA wavefront tracer behind the reference's own seam
``lum_machine_s_run(scene, lum_arr)`` (reference ``src/scene.c:1017-1028``), plus the host-side
scene-description API, flattener, pass controller and ``.pnm`` writer.  This package is a thin
ctypes binding of that library; torch is used only for device memory, streams and
``torch.distributed``.

There is no CPU rendering path: without the built extension, or without a CUDA device, the
tracer raises.
"""
from .api import (  # noqa: F401
    AcnError,
    FlatScene,
    Image,
    DeviceImage,
    Group,
    pixel_owner,
    PIX_SCALE,
    render_image_device,
    Options,
    Scene,
    Stats,
    Tracer,
    device_count,
    library_path,
    load_library,
    lum_machine_run,
    measure_fp32_peak_tflops,
    render_image,
    SEED_POSITION_HASH,
    SEED_INDEX_KEYED,
    PRECISION_F32,
    PRECISION_F64,
    CSG_AUTO,
    CSG_INTERVALS,
    CSG_MARCH,
    SPECIALIZE_AUTO,
    SPECIALIZE_ON,
    SPECIALIZE_OFF,
    spec_probe,
)
from . import scenes  # noqa: F401

__all__ = [
    "AcnError", "FlatScene", "Image", "DeviceImage", "Group", "pixel_owner", "PIX_SCALE", "render_image_device", "Options", "Scene", "Stats", "Tracer", "device_count",
    "library_path", "load_library", "lum_machine_run", "measure_fp32_peak_tflops", "render_image",
    "scenes", "SEED_POSITION_HASH", "SEED_INDEX_KEYED", "PRECISION_F32", "PRECISION_F64",
    "CSG_AUTO", "CSG_INTERVALS", "CSG_MARCH", "SPECIALIZE_AUTO", "SPECIALIZE_ON", "SPECIALIZE_OFF", "spec_probe",
]
