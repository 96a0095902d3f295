"""The reference's ten exported functions: Itanium-mangled names and ctypes signatures
(ref: include/GPUSolver.h:6-10, include/GPUImageProcessing.h:4-10, include/GPUDepthEffect.h:4-9).

No native dependency: this file is also loaded BY PATH (without importing the package, which would dlopen librtdd.so)
by oracle/binding.py to bind the reference's own library, so that bench.py --impl reference runs without our library
in the process.
"""
import ctypes as C

vp, sz, i32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float

# the reference-named C++ shims (Itanium-mangled), same ten functions as include/GPU*.h
SHIM_SYMBOLS = {
    "GPUAllocateDeviceMemory": "_Z23GPUAllocateDeviceMemoryiii",
    "GPUFreeDeviceMemory": "_Z19GPUFreeDeviceMemoryi",
    "GPULoadWeights": "_Z14GPULoadWeightsf",
    "GPUMatrixFreeSolver": "_Z19GPUMatrixFreeSolverPfmPhmS0_miififi",
    "GPUConvertToFloat": "_Z17GPUConvertToFloatPhmPfmS_mii",
    "GPUPyrDownAnnotation": "_Z20GPUPyrDownAnnotationPhmS_miiS_mS_mii",
    "GPUPaintImage": "_Z13GPUPaintImageiiiiPhmS_mii",
    "GPUSimulateDefocus": "_Z18GPUSimulateDefocusPhmPfmS_mii",
    "GPUSimulateDesaturation": "_Z23GPUSimulateDesaturationPhmS_mPfmS_mii",
    "GPUSimulateHaze": "_Z15GPUSimulateHazePhmPfmS_mii",
}

SHIM_SIGNATURES = {
    "GPUAllocateDeviceMemory": [i32, i32, i32],
    "GPUFreeDeviceMemory": [i32],
    "GPULoadWeights": [f32],
    "GPUMatrixFreeSolver": [vp, sz, vp, sz, vp, sz, i32, i32, f32, i32, f32, i32],
    "GPUConvertToFloat": [vp, sz, vp, sz, vp, sz, i32, i32],
    "GPUPyrDownAnnotation": [vp, sz, vp, sz, i32, i32, vp, sz, vp, sz, i32, i32],
    "GPUPaintImage": [i32, i32, i32, i32, vp, sz, vp, sz, i32, i32],
    "GPUSimulateDefocus": [vp, sz, vp, sz, vp, sz, i32, i32],
    "GPUSimulateDesaturation": [vp, sz, vp, sz, vp, sz, vp, sz, i32, i32],
    "GPUSimulateHaze": [vp, sz, vp, sz, vp, sz, i32, i32],
}


def bind_reference_api(cdll):
    """Return {name: callable} for the ten reference-named functions of `cdll`
    (works for librtdd.so's shims and for oracle/_ref/libref.so alike)."""
    out = {}
    for name, sym in SHIM_SYMBOLS.items():
        fn = getattr(cdll, sym)
        fn.restype = None
        fn.argtypes = SHIM_SIGNATURES[name]
        out[name] = fn
    return out
