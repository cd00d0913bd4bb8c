"""ctypes binding of ``libtemfpy_b200.so`` (the C ABI declared in ``include/temfpy_b200.h``).

The package has no CPU fallback: :func:`load` raises if the CUDA library has not been built or no
CUDA device is visible.  (The CPU test-suite injects the kernel *simulator* build through
``bind()`` from ``tests/hostsim``; nothing in this package ever loads it.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtemfpy_b200.so")

TMF_MAX_MODES = 64
SIDE_L, SIDE_R = 0, 1
OPT_SNAP, OPT_NESTED, OPT_DEVICE_PLAN, OPT_COMPLEX, OPT_PEER_OUT = 1, 2, 3, 4, 5

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_i64_p = C.POINTER(C.c_int64)
c_u64_p = C.POINTER(C.c_uint64)


class SitePlan(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "mode", "physical", "n_bra", "n_ket", "k_bra", "k_ket", "f_bra", "f_ket", "k_always",
        "s_bra", "s_ket", "n_rows", "chi_bra", "chi_ket", "n_blocks", "qtotal", "ka_bra", "ka_ket")]


class GemmJob(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p),
                ("a_idx", C.c_void_p), ("b_idx", C.c_void_p),
                ("row_scale", C.c_void_p), ("col_scale", C.c_void_p),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("lda", C.c_int), ("ldb", C.c_int), ("ldc", C.c_int),
                ("transA", C.c_int), ("transB", C.c_int),
                ("a_row_off", C.c_int), ("b_row_off", C.c_int),
                ("alpha", C.c_double), ("beta", C.c_double), ("pad_", C.c_int * 4)]


class SiteJob(C.Structure):
    _fields_ = [("Vb", C.c_void_p), ("Vk", C.c_void_p), ("bra_cols", C.c_void_p),
                ("ket_cols", C.c_void_p), ("bra_sign", C.c_void_p), ("ket_sign", C.c_void_p),
                ("O", C.c_void_p), ("S", C.c_void_p), ("det", C.c_void_p),
                ("ldb", C.c_int), ("ldk", C.c_int), ("n_bra", C.c_int), ("n_ket", C.c_int),
                ("mode", C.c_int), ("physical", C.c_int), ("ka_bra", C.c_int), ("ka_ket", C.c_int),
                ("sb", C.c_int), ("sk", C.c_int), ("emb", C.c_int), ("pad_", C.c_int * 3)]


class NestedJob(C.Structure):
    _fields_ = [("e_bra", C.c_void_p), ("e_ket", C.c_void_p), ("a_col", C.c_void_p), ("c_edge", C.c_void_p),
                ("k_bra", C.c_int), ("k_ket", C.c_int), ("df", C.c_int), ("pad_", C.c_int * 5)]


class MinorBlock(C.Structure):
    _fields_ = [("S", C.c_void_p), ("det", C.c_void_p), ("bra_masks", C.c_void_p),
                ("ket_masks", C.c_void_p), ("out", C.c_void_p),
                ("s_bra", C.c_int), ("s_ket", C.c_int), ("n_bra", C.c_int), ("n_ket", C.c_int),
                ("minor", C.c_int), ("pad_", C.c_int)]


class PairJob(C.Structure):
    _fields_ = [("V", C.c_void_p), ("tmp", C.c_void_p), ("e_raw", C.c_void_p), ("e_out", C.c_void_p),
                ("status", C.c_void_p), ("kh_out", C.c_void_p), ("rows", C.c_int), ("ld", C.c_int),
                ("k4", C.c_int), ("side", C.c_int)]


class PfBlock(C.Structure):
    _fields_ = [("N", C.c_void_p), ("bra_masks", C.c_void_p), ("ket_masks", C.c_void_p),
                ("out", C.c_void_p), ("scale", C.c_double),
                ("m", C.c_int), ("n_bra", C.c_int), ("n_ket", C.c_int), ("n1", C.c_int),
                ("n2", C.c_int), ("pad_", C.c_int)]


class PfSiteJob(C.Structure):
    _fields_ = [("S", C.c_void_p), ("rot_up", C.c_void_p), ("rot_lo", C.c_void_p), ("N", C.c_void_p),
                ("out", C.c_void_p), ("work_", C.c_void_p), ("idx1_mask", C.c_uint32), ("idx2_mask", C.c_uint32),
                ("sb", C.c_int), ("sk", C.c_int), ("sur_b", C.c_int), ("sur_k", C.c_int), ("mode", C.c_int),
                ("k1", C.c_int), ("k2", C.c_int), ("fix", C.c_int), ("want_n", C.c_int), ("no_phys", C.c_int),
                ("u_p", C.c_double), ("ket_sign", C.c_double), ("pad_", C.c_int * 4)]


class ProcrustesJob(C.Structure):
    _fields_ = [("C", C.c_void_p), ("sk", C.c_void_p), ("R", C.c_void_p), ("metrics", C.c_void_p),
                ("ldc", C.c_int64), ("ldr", C.c_int64), ("m", C.c_int), ("n", C.c_int), ("pad_", C.c_int * 2)]


class GutzJob(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("out", C.c_void_p),
                ("k_scale", C.c_void_p), ("row_scale", C.c_void_p), ("col_scale", C.c_void_p),
                ("sa_i", C.c_int64), ("sa_k", C.c_int64), ("sb_k", C.c_int64), ("sb_n", C.c_int64),
                ("so_i", C.c_int64), ("m", C.c_int), ("k", C.c_int), ("n", C.c_int), ("pad_", C.c_int * 7)]


assert C.sizeof(GemmJob) == 128 and C.sizeof(SiteJob) == 128 and C.sizeof(MinorBlock) == 64
assert C.sizeof(PairJob) == 64 and C.sizeof(PfBlock) == 64 and C.sizeof(NestedJob) == 64
assert C.sizeof(GutzJob) == 128 and C.sizeof(PfSiteJob) == 128 and C.sizeof(ProcrustesJob) == 64

# name -> (restype, argtypes); this table is also what tests check against include/temfpy_b200.h
SIGNATURES = {
    "tmf_version": (C.c_int, []),
    "tmf_last_error": (C.c_char_p, []),
    "tmf_is_cuda": (C.c_int, []),
    "tmf_device_count": (C.c_int, []),
    "tmf_corr_build": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "tmf_gemm_desc_bytes": (C.c_int64, [C.c_int]),
    "tmf_gemm_grouped": (C.c_int, [C.POINTER(GemmJob), C.c_int, C.c_void_p, C.c_void_p]),
    "tmf_projector_defect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tmf_slater_modes_workspace": (C.c_int64, [C.c_int, C.c_int, c_int_p, c_int_p, C.c_int]),
    "tmf_slater_modes_batched": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p,
                                           C.c_double, C.c_int, c_i64_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tmf_slater_modes_slot_cols": (C.c_int64, [C.c_int] * 5),
    "tmf_slater_modes_nested": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p,
                                          C.c_double, C.c_int, c_i64_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tmf_slater_modes_nested_emb": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p,
                                              C.c_double, C.c_int, c_i64_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tmf_slater_pair_bond_c_workspace": (C.c_int64, [C.c_int, C.c_int]),
    "tmf_slater_pair_bond_c": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, C.c_double,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tmf_slater_pair_bond_workspace": (C.c_int64, [C.c_int, C.c_int]),
    "tmf_slater_pair_bond": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, C.c_double,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tmf_lowest_sums": (C.c_int, [c_double_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double,
                                  c_int_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_u64_p,
                                  c_int_p, c_int_p]),
    "tmf_bond_vectors_batched": (C.c_int, [C.c_int, c_double_p, c_int_p, c_int_p, C.c_int, C.c_double,
                                           C.c_double, c_int_p, C.c_int, C.c_int, c_u64_p, c_double_p,
                                           c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, C.c_int]),
    "tmf_slater_site_plan": (C.c_int, [C.c_int] * 7 + [c_u64_p, c_int_p] + [C.c_int] * 4 +
                             [c_u64_p, c_int_p, C.POINTER(SitePlan), c_int_p, c_double_p, c_int_p,
                              c_double_p, c_u64_p, c_u64_p, c_int_p, c_int_p, c_int_p]),
    "tmf_site_desc_bytes": (C.c_int64, [C.c_int]),
    "tmf_site_overlap_schur_batched": (C.c_int, [C.POINTER(SiteJob), C.c_int, C.c_void_p, C.c_void_p]),
    "tmf_site_nested_batched": (C.c_int, [C.POINTER(SiteJob), C.POINTER(NestedJob), C.c_int, C.c_void_p,
                                          C.c_void_p]),
    "tmf_site_nested_c_batched": (C.c_int, [C.POINTER(SiteJob), C.POINTER(NestedJob), C.c_int, C.c_void_p,
                                            C.c_void_p]),
    "tmf_minors_blocks_c": (C.c_int, [C.POINTER(MinorBlock), C.c_int, C.c_void_p, C.c_void_p]),
    "tmf_minor_desc_bytes": (C.c_int64, [C.c_int]),
    "tmf_minors_blocks": (C.c_int, [C.POINTER(MinorBlock), C.c_int, C.c_void_p, C.c_void_p]),
    "tmf_chain_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                      c_int_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "tmf_chain_destroy": (None, [C.c_void_p]),
    "tmf_chain_modes_sizes": (C.c_int, [C.c_void_p, c_i64_p]),
    "tmf_chain_modes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tmf_chain_modes_enqueue": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tmf_chain_modes_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tmf_chain_enumerate": (C.c_int, [C.c_void_p]),
    "tmf_chain_enum_workspace": (C.c_int64, [C.c_void_p]),
    "tmf_chain_enumerate_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tmf_chain_tensor_sizes": (C.c_int, [C.c_void_p, c_i64_p]),
    "tmf_chain_tensors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "tmf_chain_bond": (C.c_int, [C.c_void_p, C.c_int, c_int_p, C.POINTER(c_double_p),
                                 C.POINTER(c_int_p), C.POINTER(c_u64_p), C.POINTER(c_int_p),
                                 C.POINTER(c_int_p), C.POINTER(c_double_p)]),
    "tmf_chain_site": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(SitePlan), C.POINTER(c_int_p),
                                 C.POINTER(c_i64_p), C.POINTER(c_int_p), C.POINTER(c_int_p), c_i64_p]),
    "tmf_chain_bonds_sizes": (C.c_int, [C.c_void_p, c_i64_p]),
    "tmf_chain_bonds_export": (C.c_int, [C.c_void_p] * 10),
    "tmf_chain_sites_sizes": (C.c_int, [C.c_void_p, c_i64_p]),
    "tmf_chain_sites_export": (C.c_int, [C.c_void_p] * 8),
    "tmf_chain_job_voff": (C.c_int64, [C.c_void_p, C.c_int]),
    "tmf_chain_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "tmf_chain_flops": (C.c_int, [C.c_void_p, c_double_p]),
    "tmf_launch_count": (C.c_longlong, [C.c_int]),
    "tmf_prof_enable": (C.c_int, [C.c_int]),
    "tmf_prof_report": (C.c_int, [C.c_char_p, C.c_int]),
    "tmf_prof_timeline": (C.c_int, [C.c_char_p, C.c_int]),
    "tmf_pair_tmp_doubles": (C.c_int64, [C.c_int]),
    "tmf_pfaffian_pair_modes": (C.c_int, [C.POINTER(PairJob), C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "tmf_pf_desc_bytes": (C.c_int64, [C.c_int]),
    "tmf_pfaffians_blocks": (C.c_int, [C.POINTER(PfBlock), C.c_int, C.c_void_p, C.c_void_p]),
    "tmf_pfaffian_site_finish": (C.c_int, [C.POINTER(PfSiteJob), C.c_int, C.c_void_p, C.c_void_p]),
    "tmf_gutz_desc_bytes": (C.c_int64, [C.POINTER(GutzJob), C.c_int]),
    "tmf_gutzwiller_project": (C.c_int, [C.POINTER(GutzJob), C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_void_p]),
    "tmf_canon_create": (C.c_void_p, [C.c_int, c_int_p, c_int_p, c_int_p]),
    "tmf_canon_destroy": (None, [C.c_void_p]),
    "tmf_canon_sizes": (C.c_int, [C.c_void_p, c_i64_p]),
    "tmf_canon_dims": (C.c_int, [C.c_void_p, c_int_p, c_int_p]),
    "tmf_canon_run": (C.c_int, [C.c_void_p, C.c_void_p, c_i64_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    "tmf_procrustes_workspace": (C.c_int64, [C.POINTER(ProcrustesJob), C.c_int]),
    "tmf_procrustes_blocks": (C.c_int, [C.POINTER(ProcrustesJob), C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "tmf_ipc_export": (C.c_int, [C.c_void_p, C.c_char_p, c_i64_p]),
    "tmf_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "tmf_ipc_close": (C.c_int, [C.c_void_p]),
    "tmf_host_register": (C.c_int, [C.c_void_p, C.c_int64]),
    "tmf_host_unregister": (C.c_int, [C.c_void_p]),
    "tmf_copy_d2h_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tmf_fp64_peak_probe": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_float), c_double_p, C.c_void_p]),
}

_ERRORS = {-1: ValueError, -2: AssertionError, -3: RuntimeError}


def bind(path: str) -> C.CDLL:
    """Loads a build of the C ABI and attaches the signatures."""
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = symbol missing from the build
        fn.restype = res
        fn.argtypes = args
    return lib


def check(lib, rc: int):
    """Maps a C status code to the exception type the reference raises."""
    if rc == 0:
        return
    msg = lib.tmf_last_error()
    msg = msg.decode() if msg else "unknown error"
    raise _ERRORS.get(rc, RuntimeError)(msg)


_lib = None


def load() -> C.CDLL:
    """The CUDA build of the library; fails loudly when it (or a GPU) is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"temfpy_b200: CUDA extension {LIB_PATH} has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C temfpy_b200/csrc`). "
            "There is no CPU fallback.")
    lib = bind(LIB_PATH)
    if not lib.tmf_is_cuda():
        raise RuntimeError("temfpy_b200: library at csrc/ is not a CUDA build")
    if lib.tmf_device_count() < 1:
        raise RuntimeError("temfpy_b200: no CUDA device visible; there is no CPU fallback")
    _lib = lib
    return lib
