"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref/ref_driver, built by `make -C oracle ref` from /root/reference).
Run in the build container (needs /root/reference to have been compiled):

    python oracle/make_golden.py

The fixtures are the parity anchor for both the C restatement (oracle_pa.c) and the
CUDA path; the GPU box has no /root/reference, so they are committed.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refio  # noqa: E402

OUT = os.path.join(os.path.dirname(refio.HERE), "tests", "golden")

# name -> ref_driver arguments.  Sizes are kept tiny: every array on the path is stored.
DATA = "/root/reference/data/"
CASES = {
    # BP3-style diffusion, Gauss-Legendre Q = p+2 (bench_ceed.cpp:121)
    "diff3d_p1_n3": dict(dim=3, n=3, p=1, kind="diffusion"),
    "diff3d_p2_n3": dict(dim=3, n=3, p=2, kind="diffusion"),
    "diff3d_p3_n2": dict(dim=3, n=2, p=3, kind="diffusion"),
    "diff3d_p4_n3": dict(dim=3, n=3, p=4, kind="diffusion"),
    "diff3d_p4_n234": dict(dim=3, n=2, ny=3, nz=4, p=4, kind="diffusion"),
    "diff3d_p5_n2": dict(dim=3, n=2, p=5, kind="diffusion"),
    "diff3d_p6_n2": dict(dim=3, n=2, p=6, kind="diffusion"),
    "diff3d_p7_n2": dict(dim=3, n=2, p=7, kind="diffusion"),
    "diff3d_p8_n2": dict(dim=3, n=2, p=8, kind="diffusion"),
    "diff3d_p3_n2_curved": dict(dim=3, n=2, p=3, kind="diffusion", curved=1),
    # ex1 -o 3 default rule (order 2p+dim-1 => Q = p+2 as well) on a lexicographic mesh
    "diff3d_p3_n3_nosfc": dict(dim=3, n=3, p=3, kind="diffusion", sfc=0),
    # BP1-style mass
    "mass3d_p1_n3": dict(dim=3, n=3, p=1, kind="mass"),
    "mass3d_p2_n2": dict(dim=3, n=2, p=2, kind="mass"),
    "mass3d_p4_n2": dict(dim=3, n=2, p=4, kind="mass"),
    "mass3d_p6_n2": dict(dim=3, n=2, p=6, kind="mass"),
    "mass3d_p8_n2": dict(dim=3, n=2, p=8, kind="mass"),
    "mass3d_p3_n2_curved": dict(dim=3, n=2, p=3, kind="mass", curved=1),
    # BP5-style collocated GLL diffusion (Q = p+1, B = I)
    "bp5_3d_p1_n4": dict(dim=3, n=4, p=1, kind="diffusion", quad="gll"),
    "bp5_3d_p1_n3_curved": dict(dim=3, n=3, p=1, kind="diffusion", quad="gll", curved=1),
    "bp5_3d_p2_n3": dict(dim=3, n=3, p=2, kind="diffusion", quad="gll"),
    "bp5_3d_p4_n2": dict(dim=3, n=2, p=4, kind="diffusion", quad="gll"),
    "bp5_3d_p7_n2": dict(dim=3, n=2, p=7, kind="diffusion", quad="gll"),
    "bp5_3d_p7_n2_curved": dict(dim=3, n=2, p=7, kind="diffusion", quad="gll", curved=1),
    # 2D
    "diff2d_p1_n4": dict(dim=2, n=4, p=1, kind="diffusion"),
    "diff2d_p3_n3": dict(dim=2, n=3, p=3, kind="diffusion"),
    "diff2d_p4_n3_curved": dict(dim=2, n=3, p=4, kind="diffusion", curved=1),
    "diff2d_p8_n2": dict(dim=2, n=2, p=8, kind="diffusion"),
    "mass2d_p2_n3": dict(dim=2, n=3, p=2, kind="mass"),
    "mass2d_p5_n2_curved": dict(dim=2, n=2, p=5, kind="mass", curved=1),
    "bp5_2d_p6_n2": dict(dim=2, n=2, p=6, kind="diffusion", quad="gll"),
    # vector problems, CEED BP2 / BP4 / BP6 (vdim = 3; bench_ceed.cpp:131 uses byVDIM)
    "vec_diff3d_p2_n2_byvdim": dict(dim=3, n=2, p=2, kind="diffusion", vdim=3, byvdim=1),
    "vec_diff3d_p4_n2_byvdim": dict(dim=3, n=2, p=4, kind="diffusion", vdim=3, byvdim=1),
    "vec_diff3d_p3_n2_bynodes": dict(dim=3, n=2, p=3, kind="diffusion", vdim=3, byvdim=0),
    "vec_mass3d_p3_n2_bynodes": dict(dim=3, n=2, p=3, kind="mass", vdim=3, byvdim=0),
    "vec_mass3d_p4_n2_byvdim": dict(dim=3, n=2, p=4, kind="mass", vdim=3, byvdim=1),
    "vec_bp6_3d_p3_n3_byvdim": dict(dim=3, n=3, p=3, kind="diffusion", quad="gll", vdim=3, byvdim=1),
    "vec_diff3d_p5_n2_curved": dict(dim=3, n=2, p=5, kind="diffusion", vdim=3, byvdim=1, curved=1),
    # 2D vector spaces (run-time-sized kernels on the CUDA side)
    "vec_diff2d_p3_n3_byvdim": dict(dim=2, n=3, p=3, kind="diffusion", vdim=2, byvdim=1),
    "vec_mass2d_p4_n2_bynodes": dict(dim=2, n=2, p=4, kind="mass", vdim=2, byvdim=0, curved=1),
    # non-symmetric matrix coefficient (dim*dim quadrature-data components)
    "nonsym_diff3d_p2_n2": dict(dim=3, n=2, p=2, kind="diffusion", nonsym=1),
    "nonsym_diff3d_p4_n2_curved": dict(dim=3, n=2, p=4, kind="diffusion", nonsym=1, curved=1),
    "nonsym_diff3d_p7_n2": dict(dim=3, n=2, p=7, kind="diffusion", nonsym=1),
    "nonsym_diff2d_p3_n3_curved": dict(dim=2, n=3, p=3, kind="diffusion", nonsym=1, curved=1),
    "nonsym_bp5_3d_p3_n2": dict(dim=3, n=2, p=3, kind="diffusion", quad="gll", nonsym=1),
    # OperatorChebyshevSmoother: coefficients/action for orders 1..5, power method, CG + Chebyshev
    "cheb_diff3d_p2_n3": dict(dim=3, n=3, p=2, kind="diffusion", cheb=1, order=3),
    "cheb_diff3d_p4_n2_curved": dict(dim=3, n=2, p=4, kind="diffusion", cheb=1, curved=1, order=2),
    "cheb_diff2d_p3_n4": dict(dim=2, n=4, p=3, kind="diffusion", cheb=1, order=4),
    "cheb_mass3d_p3_n2": dict(dim=3, n=2, p=3, kind="mass", cheb=1, order=5),
    "cheb_bp5_3d_p5_n2": dict(dim=3, n=2, p=5, kind="diffusion", quad="gll", cheb=1, order=3),
    # integrators restricted by element attribute (AddDomainIntegrator(bfi, elem_marker))
    "mark_3d_p2_n3": dict(dim=3, n=3, p=2, kind="diffusion", markers=1, nattr=3, maxit=25),
    "mark_3d_p4_n2_curved": dict(dim=3, n=2, p=4, kind="diffusion", markers=1, nattr=3, curved=1, maxit=25),
    "mark_3d_p3_n3_plain_first": dict(dim=3, n=3, p=3, kind="diffusion", markers=1, nattr=2, marker_order=1,
                                      jacobi=0, maxit=25),
    "mark_2d_p3_n5": dict(dim=2, n=5, p=3, kind="diffusion", markers=1, nattr=3, maxit=25),
    # meshes from the reference's data directory (unstructured numbering, twisted orientations,
    # curved Q2/Q3 nodes); inline-hex -o 3 is BASELINE config C1 (here refined once, not 3 times)
    "mesh_inline_hex_r1_p3": dict(mesh=DATA + "inline-hex.mesh", refine=1, p=3, kind="diffusion", lean=1),
    "mesh_toroid_hex_r1_p2": dict(mesh=DATA + "toroid-hex.mesh", refine=1, p=2, kind="diffusion", lean=1),
    "mesh_toroid_hex_r0_p5_mass": dict(mesh=DATA + "toroid-hex.mesh", refine=0, p=5, kind="mass", lean=1),
    "mesh_fichera_q2_r1_p4": dict(mesh=DATA + "fichera-q2.mesh", refine=1, p=4, kind="diffusion", lean=1),
    "mesh_star_q3_r1_p3": dict(mesh=DATA + "star-q3.mesh", refine=1, p=3, kind="diffusion", lean=1),
    "mesh_fichera_q3_r0_p7_bp5": dict(mesh=DATA + "fichera-q3.mesh", refine=0, p=7, kind="diffusion", quad="gll",
                                      lean=1),
    # DomainLFIntegrator / VectorDomainLFIntegrator device assembly (LinearForm::UseFastAssembly)
    "lf_3d_p2_n3": dict(dim=3, n=3, p=2, kind="mass", lf=1, nattr=2),
    "lf_3d_p4_n2_curved": dict(dim=3, n=2, p=4, kind="mass", lf=1, nattr=3, curved=1),
    "lf_3d_p7_n2_gll": dict(dim=3, n=2, p=7, kind="mass", quad="gll", lf=1, nattr=2),
    "lf_2d_p3_n4_curved": dict(dim=2, n=4, p=3, kind="mass", lf=1, nattr=2, curved=1),
    "lf_3d_p3_n2_vec_byvdim": dict(dim=3, n=2, p=3, kind="mass", lf=1, nattr=2, vdim=3, byvdim=1),
    "lf_3d_p2_n2_vec_bynodes": dict(dim=3, n=2, p=2, kind="mass", lf=1, nattr=2, vdim=3, byvdim=0, curved=1),
    # LOR batched assembly, H1 (fem/lor/lor_batched.cpp, lor_h1_impl.hpp): macro-element stencil array,
    # CSR without / with essential dofs eliminated, and the legacy refined-mesh assembly of the same form
    "lor_3d_p1_n3": dict(dim=3, n=3, p=1, lor=1, lorkind="diffusion"),
    "lor_3d_p2_n3": dict(dim=3, n=3, p=2, lor=1, lorkind="both", varcoef=1),
    "lor_3d_p3_n2_curved": dict(dim=3, n=2, p=3, lor=1, lorkind="both", varcoef=1, curved=1),
    "lor_3d_p4_n2": dict(dim=3, n=2, p=4, lor=1, lorkind="diffusion"),
    "lor_3d_p4_n2_mass": dict(dim=3, n=2, p=4, lor=1, lorkind="mass", mass_const=2.5),
    "lor_3d_p6_n2_curved": dict(dim=3, n=2, p=6, lor=1, lorkind="both", curved=1, mass_const=0.5, diff_const=3.0),
    "lor_3d_p8_n1": dict(dim=3, n=1, ny=2, nz=1, p=8, lor=1, lorkind="diffusion", varcoef=1),
    "lor_mesh_fichera_q2_p3": dict(mesh=DATA + "fichera-q2.mesh", refine=0, p=3, lor=1, lorkind="both", varcoef=1),
    "lor_mesh_toroid_hex_p2": dict(mesh=DATA + "toroid-hex.mesh", refine=0, p=2, lor=1, lorkind="diffusion"),
    # 2D (BatchedLOR_H1::Assemble2D, 9-point stencil)
    "lor_2d_p1_n4": dict(dim=2, n=4, p=1, lor=1, lorkind="diffusion"),
    "lor_2d_p3_n3_curved": dict(dim=2, n=3, p=3, lor=1, lorkind="both", varcoef=1, curved=1),
    "lor_2d_p5_n2": dict(dim=2, n=2, ny=3, p=5, lor=1, lorkind="both", mass_const=0.5, diff_const=2.0),
    "lor_2d_p8_n2_curved": dict(dim=2, n=2, p=8, lor=1, lorkind="diffusion", varcoef=1, curved=1),
    "lor_mesh_star_q3_p4": dict(mesh=DATA + "star-q3.mesh", refine=0, p=4, lor=1, lorkind="both", varcoef=1),
}


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    os.makedirs(OUT, exist_ok=True)
    if not refio.have_ref_driver():
        raise SystemExit("oracle/_ref/ref_driver missing: run `make -C oracle ref` first")
    total = 0
    for name, kw in CASES.items():
        if only and not name.startswith(only):
            continue
        d = refio.run_dump(rtol=1e-8, **kw)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        sz = os.path.getsize(path)
        total += sz
        its = int(d["pcg_iters"][0]) if "pcg_iters" in d else -1
        print(f"{name:28s} NE={int(d['NE'][0]):4d} ndofs={int(d['ndofs'][0]):6d} "
              f"pcg_iters={its:4d} {sz/1024:8.1f} KiB")
    print(f"total {total/1e6:.2f} MB")


if __name__ == "__main__":
    main()
