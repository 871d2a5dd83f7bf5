"""mfem_b200/capi.py -- ctypes binding of libb200fem.so (include/b200fem.h).

The Python layer is plumbing for tests and bench.py (the product is the C-ABI library and
the C++ adapter in adapter/): it loads the library built in-tree by
`make -C mfem_b200/csrc`, declares every prototype of include/b200fem.h, and turns
non-zero status codes into exceptions.  There is no fallback: if the library is missing
the import of any compute entry point fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B2F_LIB: an alternative build of the same library (kernel tuning experiments, tools/opbench.py)
LIB_PATH = os.environ.get("B2F_LIB") or os.path.join(HERE, "lib", "libb200fem.so")

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)
vp = C.c_void_p
i64 = C.c_int64


class B2FError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libb200fem error {code}: {text}")
        self.code = code


class PcgOptions(C.Structure):
    _fields_ = [("rel_tol", C.c_double), ("abs_tol", C.c_double), ("max_iter", C.c_int),
                ("check_every", C.c_int)]


class PcgResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int),
                ("initial_norm", C.c_double), ("final_norm", C.c_double)]


# name -> (restype, argtypes); every symbol include/b200fem.h declares
PROTOTYPES = {
    "b2f_version": (C.c_int, []),
    "b2f_last_error": (C.c_char_p, []),
    "b2f_ctx_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "b2f_ctx_destroy": (None, [vp]),
    "b2f_ctx_sync": (C.c_int, [vp]),
    "b2f_ctx_device": (C.c_int, [vp]),
    "b2f_ctx_sm_count": (C.c_int, [vp]),
    "b2f_ctx_launch_count": (i64, [vp]),
    "b2f_ctx_reset_launch_count": (None, [vp]),
    "b2f_ctx_kernel_timing": (C.c_int, [vp, C.c_int]),
    "b2f_ctx_kernel_time": (C.c_int, [vp, c_double_p, c_int_p]),
    "b2f_malloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "b2f_free": (C.c_int, [vp, vp]),
    "b2f_malloc_host": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "b2f_free_host": (C.c_int, [vp, vp]),
    "b2f_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "b2f_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "b2f_memcpy_d2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "b2f_memset": (C.c_int, [vp, vp, C.c_int, C.c_size_t]),
    "b2f_gauss_legendre_h": (C.c_int, [C.c_int, vp, vp]),
    "b2f_gauss_lobatto_h": (C.c_int, [C.c_int, vp, vp]),
    "b2f_h1_dof_to_quad_h": (C.c_int, [C.c_int, C.c_int, vp, vp, vp]),
    "b2f_tensor_weights_h": (C.c_int, [C.c_int, C.c_int, vp, vp]),
    "b2f_vector_randomize_h": (C.c_int, [vp, i64, C.c_int]),
    "b2f_halo_tables_from_groups_h": (C.c_int, [C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp,
                                                vp, vp, vp, vp, vp, vp, vp]),
    "b2f_cartesian_h1_space_h": (C.c_int, [C.c_int] * 6 + [c_int_p, c_int_p, vp, vp, c_int_p, vp, vp]),
    "b2f_restriction_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                         C.POINTER(vp)]),
    "b2f_restriction_create_from_elem_dof": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                       vp, vp, C.POINTER(vp)]),
    "b2f_restriction_destroy": (None, [vp]),
    "b2f_restriction_tables": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "b2f_restriction_mult": (C.c_int, [vp, vp, vp]),
    "b2f_restriction_abs_mult": (C.c_int, [vp, vp, vp]),
    "b2f_restriction_mult_transpose": (C.c_int, [vp, vp, vp, C.c_int]),
    "b2f_restriction_abs_mult_transpose": (C.c_int, [vp, vp, vp]),
    "b2f_diffusion_setup": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, i64, vp, vp, vp, vp, vp, C.c_int,
                                      i64, C.POINTER(vp)]),
    "b2f_mass_setup": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, i64, vp, vp, vp, vp, i64,
                                 C.POINTER(vp)]),
    "b2f_pa_setup_from_nodes": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, i64, vp, vp, vp,
                                          C.c_int, vp, vp, vp, C.c_double, C.POINTER(vp)]),
    "b2f_pa_destroy": (None, [vp]),
    "b2f_pa_data": (C.c_int, [vp, C.POINTER(vp), C.POINTER(i64)]),
    "b2f_pa_add_mult": (C.c_int, [vp, vp, vp]),
    "b2f_pa_add_abs_mult": (C.c_int, [vp, vp, vp]),
    "b2f_pa_add_mult_transpose": (C.c_int, [vp, vp, vp]),
    "b2f_pa_add_mult_vec": (C.c_int, [vp, C.c_int, vp, vp]),
    "b2f_pa_assemble_diagonal": (C.c_int, [vp, vp]),
    "b2f_domain_lf_assemble": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, i64, vp, vp, vp, vp, vp, i64,
                                         vp]),
    "b2f_lor_vertex_coordinates": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, i64, vp, vp, vp]),
    "b2f_lor_assemble_h1": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, i64, vp, i64, vp, C.c_int, C.c_int,
                                      C.POINTER(vp)]),
    "b2f_lor_csr": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(i64), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "b2f_lor_sparse_ij": (C.c_int, [vp, C.POINTER(vp), C.POINTER(i64)]),
    "b2f_lor_kernel_ms": (C.c_int, [vp, C.POINTER(C.c_double)]),
    "b2f_lor_mult": (C.c_int, [vp, vp, vp]),
    "b2f_lor_destroy": (None, [vp]),
    "b2f_operator_create": (C.c_int, [vp, vp, C.POINTER(vp), C.c_int, vp, C.c_int, C.POINTER(vp)]),
    "b2f_operator_destroy": (None, [vp]),
    "b2f_operator_mult": (C.c_int, [vp, vp, vp]),
    "b2f_operator_mult_unfused": (C.c_int, [vp, vp, vp]),
    "b2f_operator_mult_constrained": (C.c_int, [vp, vp, vp]),
    "b2f_operator_abs_mult": (C.c_int, [vp, vp, vp]),
    "b2f_operator_abs_mult_constrained": (C.c_int, [vp, vp, vp]),
    "b2f_operator_set_diag_policy": (C.c_int, [vp, C.c_int]),
    "b2f_operator_assemble_diagonal_constrained": (C.c_int, [vp, vp]),
    "b2f_operator_assemble_diagonal": (C.c_int, [vp, vp]),
    "b2f_operator_eliminate_rhs": (C.c_int, [vp, vp, vp]),
    "b2f_operator_set_halo": (C.c_int, [vp, vp]),
    "b2f_operator_height": (C.c_int, [vp]),
    "b2f_operator_set_interior": (C.c_int, [vp, i64]),
    "b2f_operator_set_markers": (C.c_int, [vp, C.c_int, vp, vp, C.c_int]),
    "b2f_jacobi_setup": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, C.c_double, vp]),
    "b2f_jacobi_mult": (C.c_int, [vp, C.c_int, vp, vp, vp]),
    "b2f_vector_dot": (C.c_int, [vp, i64, vp, vp, c_double_p]),
    "b2f_vector_add": (C.c_int, [vp, i64, vp, C.c_double, vp, vp]),
    "b2f_pcg": (C.c_int, [vp, vp, vp, vp, C.POINTER(PcgOptions), C.POINTER(PcgResult), vp, C.c_int]),
    "b2f_chebyshev_coefficients_h": (C.c_int, [C.c_int, C.c_double, c_double_p]),
    "b2f_power_method": (C.c_int, [vp, vp, vp, C.c_int, C.c_double, c_double_p]),
    "b2f_chebyshev_create": (C.c_int, [vp, vp, C.c_int, C.c_double, C.POINTER(vp)]),
    "b2f_chebyshev_destroy": (None, [vp]),
    "b2f_chebyshev_mult": (C.c_int, [vp, vp, vp]),
    "b2f_pcg_chebyshev": (C.c_int, [vp, vp, vp, vp, C.POINTER(PcgOptions), C.POINTER(PcgResult), vp,
                                    C.c_int]),
    "b2f_nccl_unique_id_h": (C.c_int, [vp]),
    "b2f_halo_create": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp,
                                  vp, vp, C.POINTER(vp)]),
    "b2f_halo_destroy": (None, [vp]),
    "b2f_halo_p2p_blob_bytes": (C.c_int, []),
    "b2f_halo_p2p_export": (C.c_int, [vp, vp]),
    "b2f_halo_p2p_connect": (C.c_int, [vp, vp, C.c_int]),
    "b2f_halo_mult": (C.c_int, [vp, vp, vp]),
    "b2f_halo_mult_transpose": (C.c_int, [vp, vp, vp]),
    "b2f_halo_dot": (C.c_int, [vp, i64, vp, vp, c_double_p]),
    "b2f_cartesian_partition_h": (C.c_int, [C.c_int] * 10 + [c_int_p] + [vp] * 11),
}

_lib = None


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C mfem_b200/csrc -j8` "
                              "(or __graft_entry__.build()); there is no fallback path")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise B2FError(rc, lib().b2f_last_error().decode(errors="replace"))


def call(name, *args):
    check(getattr(lib(), name)(*args))
