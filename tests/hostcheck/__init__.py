"""Builds tests/hostcheck/hostcheck.cpp: the PRODUCT's host/device headers compiled for the host, so
CPU tests can check their algorithms against the oracle.  Test-only; never loaded by the package."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "libhostcheck.so")
CSRC = os.path.join(ROOT, "pokerl_b200", "csrc")


def load():
    src = [os.path.join(HERE, "hostcheck.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-fPIC", "-shared", "-x", "c++", "-std=c++17", "-I", CSRC, "-o", SO, src[0]],
                       check=True, capture_output=True)
    L = C.CDLL(SO)
    L.hc_eval_mask.restype = C.c_uint32
    L.hc_eval_mask.argtypes = [C.c_uint64]
    L.hc_eval_mask_tbl.restype = C.c_uint32
    L.hc_eval_mask_tbl.argtypes = [C.c_uint64]
    L.hc_eval_tbl_mismatches.restype = C.c_uint64
    L.hc_eval_tbl_mismatches.argtypes = [C.c_int, C.c_uint64]
    L.hc_eval_cards.restype = C.c_uint32
    L.hc_eval_cards.argtypes = [C.c_void_p, C.c_int]
    L.hc_eval_cards_literal.restype = C.c_uint32
    L.hc_eval_cards_literal.argtypes = [C.c_void_p, C.c_int]
    L.hc_eval_masks.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.hc_eval_exhaustive.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.hc_deal.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    L.hc_sample.restype = C.c_int
    L.hc_sample.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    L.hc_compare.restype = C.c_int
    L.hc_compare.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return L
