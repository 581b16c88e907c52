"""ctypes mirror of include/monopsr_b200_net.h."""
import ctypes

c_i = ctypes.c_int
c_p = ctypes.c_void_p

TC_FWD, TC_DGRAD, TC_WGRAD = 0, 1, 2


class TcGemmParams(ctypes.Structure):
    _fields_ = [
        ("op", c_i), ("H", c_i), ("W", c_i), ("kh", c_i), ("kw", c_i), ("dil", c_i), ("M", c_i),
        ("Cin", c_i), ("Cout", c_i),
        ("X", c_p), ("ldx", c_i),
        ("Y", c_p), ("ldy", c_i),
        ("Wt", c_p), ("ldw", c_i),
        ("out", c_p), ("ldo", c_i),
        ("tapmask", c_p),
        ("scale", c_p), ("shift", c_p),
        ("res", c_p), ("ldr", c_i),
        ("mask", c_p), ("ldm", c_i),
        ("scale2", c_p), ("colsum", c_p),
        ("relu", c_i), ("round_tf32", c_i), ("atomic", c_i), ("ksplit", c_i),
    ]


def declare(lib):
    lib.mpb_tc_gemm.argtypes = [ctypes.POINTER(TcGemmParams), c_i, c_p]
    lib.mpb_tc_gemm.restype = c_i
    lib.mpb_build_tapmask.argtypes = [c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]
    lib.mpb_build_tapmask.restype = c_i
