// adapter/ex1_b200.cpp -- examples/ex1.cpp (-pa) with the PA hot path routed to libb200fem.so.
//
// Same driver flow as the reference's ex1 (examples/ex1.cpp:122-267): mesh, H1 space, essential
// boundary, RHS (1, phi_i), BilinearForm with AssemblyLevel::PARTIAL and a DiffusionIntegrator,
// FormLinearSystem, Jacobi-PCG.  It runs the problem TWICE in one process -- MFEM's own CPU PA
// path and the B200 path behind the same interfaces (adapter/b200_mfem.hpp) -- and checks
//   seam 3: MFEM's PABilinearFormExtension::Mult with the b200 integrator == MFEM's   (1e-12)
//   seam 1: fused b200::PAOperator::Mult == MFEM's a.Mult, AssembleDiagonal          (1e-12)
//   seam 2: the ex1 driver itself (examples/ex1.cpp:203-268) with `b200::BilinearForm` as the only
//           changed type: STOCK integrators, a.Assemble(), a.FormLinearSystem(), MFEM's own
//           OperatorJacobiSmoother M(a, ess), PCG(*A, M, B, X, ...), a.RecoverFEMSolution()
//   solve : b200::CGSolver iteration count within +-1 of mfem::CGSolver, same solution.
// -d cuda (an MFEM built with MFEM_USE_CUDA, oracle/_ref_cuda): device-resident mode, every call
// takes MFEM's own device pointers (x.Read() / y.Write()); -d cpu: host-MFEM mode.
// usage: ex1_b200 [-d cpu|cuda] [-o order] [-n elems_per_side | -m mesh -r refinements] [-k mass|diffusion]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "b200_mfem.hpp"

using namespace mfem;

static double relerr(const Vector &a, const Vector &b)
{
   Vector d(a);
   d -= b;
   return d.Norml2() / b.Norml2();
}

int main(int argc, char *argv[])
{
   int order = 3, n = 6, ref = 0, vdim = 1;
   const char *mesh_file = "";
   const char *kind = "diffusion";
   const char *device_config = "cpu";
   OptionsParser args(argc, argv);
   args.AddOption(&device_config, "-d", "--device", "MFEM device configuration string (examples/ex1.cpp:76)");
   args.AddOption(&order, "-o", "--order", "H1 order");
   args.AddOption(&n, "-n", "--elems", "elements per side of the Cartesian hex mesh");
   args.AddOption(&mesh_file, "-m", "--mesh", "mesh file (instead of the Cartesian mesh)");
   args.AddOption(&ref, "-r", "--refine", "uniform refinements of the mesh file");
   args.AddOption(&kind, "-k", "--kind", "diffusion | mass");
   args.AddOption(&vdim, "-vd", "--vdim", "vector dimension (3: CEED BP2/BP4 style vector problem, byVDIM)");
   args.Parse();
   if (!args.Good()) { args.PrintUsage(std::cout); return 1; }
   Device device(device_config); // "cpu": host-MFEM mode; "cuda": device-resident mode
   device.Print();
   std::printf("b200 memory mode: %s\n", b200::DeviceMode() ? "device-resident (MFEM device pointers)" : "host-MFEM (adapter-owned copies)");

   Mesh mesh = strlen(mesh_file) ? Mesh(mesh_file, 1, 1)
               : Mesh::MakeCartesian3D(n, n, n, Element::HEXAHEDRON);
   for (int l = 0; l < ref; l++) { mesh.UniformRefinement(); }
   const int dim = mesh.Dimension();
   H1_FECollection fec(order, dim);
   FiniteElementSpace fes(&mesh, &fec, vdim, vdim > 1 ? Ordering::byVDIM : Ordering::byNODES);
   Array<int> ess_bdr(mesh.bdr_attributes.Max()), ess;
   ess_bdr = 1;
   fes.GetEssentialTrueDofs(ess_bdr, ess);
   ConstantCoefficient one(1.0);
   const bool mass = !strcmp(kind, "mass");
   std::printf("ex1_b200: %s, order %d, %d elements, %d unknowns\n", kind, order, mesh.GetNE(),
               fes.GetTrueVSize());

   auto make_form = [&](bool b200) -> BilinearForm *
   {
      BilinearForm *a = new BilinearForm(&fes);
      a->SetAssemblyLevel(AssemblyLevel::PARTIAL);
      if (vdim > 1)
      {
         if (mass) { a->AddDomainIntegrator(b200 ? (BilinearFormIntegrator *)new b200::VectorMassIntegrator(one) : new VectorMassIntegrator(one)); }
         else { a->AddDomainIntegrator(b200 ? (BilinearFormIntegrator *)new b200::VectorDiffusionIntegrator(one) : new VectorDiffusionIntegrator(one)); }
      }
      else if (mass) { a->AddDomainIntegrator(b200 ? (BilinearFormIntegrator *)new b200::MassIntegrator(one) : new MassIntegrator(one)); }
      else { a->AddDomainIntegrator(b200 ? (BilinearFormIntegrator *)new b200::DiffusionIntegrator(one) : new DiffusionIntegrator(one)); }
      a->Assemble();
      return a;
   };
   std::unique_ptr<BilinearForm> a_ref(make_form(false)), a_b2(make_form(true));

   int fails = 0;
   auto check = [&](const char *what, double err, double tol)
   {
      std::printf("  %-58s %.3e  %s\n", what, err, err <= tol ? "ok" : "FAIL");
      fails += !(err <= tol);
   };
   const int N = fes.GetVSize();
   Vector x(N), y_ref(N), y(N);
   x.Randomize(1);
   a_ref->Mult(x, y_ref);
   a_b2->Mult(x, y); // MFEM's extension + restriction, b200 element kernel (seam 3)
   check("seam 3: PABilinearFormExtension::Mult with b200 integrator", relerr(y, y_ref), 1e-12);
   Vector d_ref(N), d(N);
   a_ref->AssembleDiagonal(d_ref);
   a_b2->AssembleDiagonal(d);
   check("seam 3: AssembleDiagonal with b200 integrator", relerr(d, d_ref), 1e-12);

   Array<int> none;
   b200::PAOperator A_fused(*a_b2, none);
   A_fused.Mult(x, y);
   check("seam 1: fused b200::PAOperator::Mult", relerr(y, y_ref), 1e-12);
   A_fused.AssembleDiagonal(d);
   check("seam 1: b200::PAOperator::AssembleDiagonal", relerr(d, d_ref), 1e-12);

   // the linear system of ex1: b = (1, phi_i), x = 0 on the boundary
   Vector b(N);
   if (vdim == 1)
   {
      LinearForm lf(&fes);
      lf.AddDomainIntegrator(new DomainLFIntegrator(one));
      lf.Assemble();
      b = lf;
   }
   else { b.Randomize(2); }
   GridFunction u(&fes);
   u = 0.0;
   OperatorPtr A;
   Vector B, X;
   a_ref->FormLinearSystem(ess, u, b, A, X, B);
   Vector yc_ref(N), yc(N);
   A->Mult(x, yc_ref);
   b200::ConstrainedOperator A_b2(*a_b2, ess);
   A_b2.Mult(x, yc);
   check("seam 1: b200::ConstrainedOperator::Mult", relerr(yc, yc_ref), 1e-12);
   Vector B2(b);
   Vector zero(N);
   zero = 0.0;
   A_b2.EliminateRHS(zero, B2);
   check("EliminateRHS", relerr(B2, B), 1e-12);

   OperatorJacobiSmoother M(*a_ref, ess);
   CGSolver cg;
   cg.SetRelTol(1e-6); cg.SetAbsTol(0.0); cg.SetMaxIter(2000); cg.SetPrintLevel(-1);
   cg.SetOperator(*A);
   cg.SetPreconditioner(M);
   X = 0.0;
   cg.Mult(B, X);

   b200::CGSolver cg2;
   cg2.SetRelTol(1e-6); cg2.SetAbsTol(0.0); cg2.SetMaxIter(2000);
   cg2.SetOperator(A_b2);
   cg2.UseJacobi(ess);
   Vector X2(N);
   X2 = 0.0; // an IterativeSolver starts from x (iterative_mode is true by default, linalg/solvers.hpp:102)
   cg2.Mult(B2, X2);
   std::printf("  PCG iterations: MFEM %d, b200 %d\n", cg.GetNumIterations(), cg2.GetNumIterations());
   check("PCG iteration count |difference|", std::abs(cg.GetNumIterations() - cg2.GetNumIterations()), 1.0);
   check("PCG solution", relerr(X2, X), 1e-5);

   // seam 2: the ex1 driver, unchanged except for the form's type (stock integrator, stock
   // Jacobi smoother, FormLinearSystem / RecoverFEMSolution of the form itself)
   if (vdim == 1)
   {
      b200::BilinearForm a(&fes);
      a.SetAssemblyLevel(AssemblyLevel::PARTIAL);
      if (mass) { a.AddDomainIntegrator(new MassIntegrator(one)); }
      else { a.AddDomainIntegrator(new DiffusionIntegrator(one)); }
      a.Assemble();
      check("seam 2: the form took the fused path (1 = yes)", a.Fused() ? 0.0 : 1.0, 0.0);
      a.Mult(x, y);
      check("seam 2: b200::BilinearForm::Mult (stock integrator)", relerr(y, y_ref), 1e-12);
      a.AssembleDiagonal(d);
      check("seam 2: b200::BilinearForm::AssembleDiagonal", relerr(d, d_ref), 1e-12);
      GridFunction u2(&fes);
      u2 = 0.0;
      Vector b2(b);
      OperatorPtr A2;
      Vector Bs, Xs;
      a.FormLinearSystem(ess, u2, b2, A2, Xs, Bs);
      A2->Mult(x, yc);
      check("seam 2: FormLinearSystem operator (fused, constrained)", relerr(yc, yc_ref), 1e-12);
      check("seam 2: FormLinearSystem right-hand side", relerr(Bs, B), 1e-12);
      OperatorJacobiSmoother M2(a, ess); // MFEM's own smoother over the form's (fused) diagonal
      b200::PCG(*A2, M2, Bs, Xs, 0, 2000, 1e-12, 0.0);
      a.RecoverFEMSolution(Xs, b2, u2);
      check("seam 2: PCG(*A, M, B, X) + RecoverFEMSolution", relerr(u2, X), 1e-5);
      // the same through mfem::CGSolver: MFEM's own solver driving the fused operator
      CGSolver cgm;
      cgm.SetRelTol(1e-6); cgm.SetAbsTol(0.0); cgm.SetMaxIter(2000); cgm.SetPrintLevel(-1);
      cgm.SetOperator(*A2);
      cgm.SetPreconditioner(M2);
      Vector X4(N);
      X4.UseDevice(true);
      X4 = 0.0;
      cgm.Mult(Bs, X4);
      check("seam 2: mfem::CGSolver over the fused operator, iteration count |difference|",
            std::abs(cg.GetNumIterations() - cgm.GetNumIterations()), 1.0);
      check("seam 2: mfem::CGSolver over the fused operator, solution", relerr(X4, X), 1e-5);
      // MultTranspose: the same operator for symmetric coefficients (bilininteg_diffusion_pa.cpp:95)
      a.MultTranspose(x, y);
      check("seam 2: MultTranspose (symmetric coefficient)", relerr(y, y_ref), 1e-12);
   }

   // spatially varying scalar and full-matrix coefficients through the reference's own
   // CoefficientVector projection (seam 3 + the fused operator)
   if (vdim == 1 && dim == 3)
   {
      FunctionCoefficient kf([](const Vector &xx) { return 1.0 + 0.5 * xx(0) * xx(1) + xx(2) * xx(2); });
      MatrixFunctionCoefficient mf(3, [](const Vector &xx, DenseMatrix &K)
      {
         K.SetSize(3);
         K = 0.0;
         for (int i = 0; i < 3; i++) { K(i, i) = 2.0 + xx(i); }
         K(0, 1) = 0.3 * xx(2); K(1, 0) = -0.1; K(1, 2) = 0.2 * xx(0); K(2, 0) = 0.15;
      });
      for (int variant = 0; variant < 3; variant++)
      {
         BilinearForm ar(&fes), ab(&fes);
         ar.SetAssemblyLevel(AssemblyLevel::PARTIAL);
         ab.SetAssemblyLevel(AssemblyLevel::PARTIAL);
         if (variant == 0) { ar.AddDomainIntegrator(new DiffusionIntegrator(kf)); ab.AddDomainIntegrator(new b200::DiffusionIntegrator(kf)); }
         if (variant == 1) { ar.AddDomainIntegrator(new DiffusionIntegrator(mf)); ab.AddDomainIntegrator(new b200::DiffusionIntegrator(mf)); }
         if (variant == 2) { ar.AddDomainIntegrator(new MassIntegrator(kf)); ab.AddDomainIntegrator(new b200::MassIntegrator(kf)); }
         ar.Assemble();
         ab.Assemble();
         const char *nm[3] = {"function-coefficient diffusion", "full-matrix-coefficient diffusion",
                              "function-coefficient mass"};
         char msg[128];
         ar.Mult(x, y_ref);
         ab.Mult(x, y);
         std::snprintf(msg, sizeof msg, "seam 3: %s, Mult", nm[variant]);
         check(msg, relerr(y, y_ref), 1e-12);
         b200::PAOperator Af(ab, none);
         Af.Mult(x, y);
         std::snprintf(msg, sizeof msg, "seam 1: %s, fused Mult", nm[variant]);
         check(msg, relerr(y, y_ref), 1e-12);
         ar.AssembleDiagonal(d_ref);
         Af.AssembleDiagonal(d);
         std::snprintf(msg, sizeof msg, "seam 1: %s, diagonal", nm[variant]);
         check(msg, relerr(d, d_ref), 1e-12);
      }
   }

   // LOR matrix (what LORSolver hands to an AMG): mfem::BatchedLORAssembly vs b200::BatchedLORAssembly
   if (vdim == 1 && dim == 3 && order <= 8)
   {
      FunctionCoefficient mcf([](const Vector &xx) { return 1.0 + xx(0) + 2.0 * xx(1) * xx(2); });
      FunctionCoefficient dcf([](const Vector &xx) { return 2.0 + sin(3.0 * xx(0)) * cos(2.0 * xx(1)); });
      BilinearForm al(&fes);
      al.AddDomainIntegrator(new MassIntegrator(mcf));
      al.AddDomainIntegrator(new DiffusionIntegrator(dcf));
      OperatorHandle Ar, Ab;
      BatchedLORAssembly lr(fes);
      lr.Assemble(al, ess, Ar);
      check("LOR: the form is supported (0 = yes)", b200::BatchedLORAssembly::FormIsSupported(al) ? 0.0 : 1.0, 0.0);
      b200::BatchedLORAssembly lb(fes);
      lb.Assemble(al, ess, Ab);
      SparseMatrix &Sr = *Ar.As<SparseMatrix>(), &Sb = *Ab.As<SparseMatrix>();
      check("LOR: number of nonzeros |difference|", std::abs(Sr.NumNonZeroElems() - Sb.NumNonZeroElems()), 0.0);
      Sr.Mult(x, y_ref);
      Sb.Mult(x, y);
      check("LOR: SparseMatrix::Mult of the assembled matrix", relerr(y, y_ref), 1e-12);
      std::unique_ptr<SparseMatrix> diff(Add(1.0, Sr, -1.0, Sb));
      check("LOR: max |A_mfem - A_b200| / max |A_mfem|", diff->MaxNorm() / Sr.MaxNorm(), 1e-12);
      if (!b200::DeviceMode())
      {
         // on the host MFEM fills rows serially: the CSR arrays themselves must coincide
         bool same = true;
         const int *Ir = Sr.HostReadI(), *Ib = Sb.HostReadI(), *Jr = Sr.HostReadJ(), *Jb = Sb.HostReadJ();
         for (int i = 0; i <= N && same; i++) { same = Ir[i] == Ib[i]; }
         for (int k = 0; k < Sr.NumNonZeroElems() && same; k++) { same = Jr[k] == Jb[k]; }
         check("LOR: I and J identical to MFEM's (0 = yes)", same ? 0.0 : 1.0, 0.0);
      }
   }

   // right-hand side with a spatially varying coefficient, fast (device) assembly
   if (vdim == 1)
   {
      FunctionCoefficient f([](const Vector &xx) { return 1.0 + xx(0) + 2.0 * xx(1) * xx(1) + (xx.Size() > 2 ? sin(3.0 * xx(2)) : 0.0); });
      LinearForm l_ref(&fes), l_b2(&fes);
      l_ref.AddDomainIntegrator(new DomainLFIntegrator(f));
      l_ref.UseFastAssembly(true);
      l_ref.Assemble();
      l_b2.AddDomainIntegrator(new b200::DomainLFIntegrator(f));
      l_b2.UseFastAssembly(true);
      l_b2.Assemble();
      check("LinearForm fast assembly with b200::DomainLFIntegrator", relerr(l_b2, l_ref), 1e-12);
   }

   // Chebyshev smoother (power-method constructor, MFEM's own start vector) as preconditioner
   {
      Vector dg(N);
      a_ref->AssembleDiagonal(dg);
      OperatorChebyshevSmoother S(*A, dg, ess, 3, 10, 1e-8, 12345);
      b200::OperatorChebyshevSmoother S2(A_b2, dg, ess, 3, 10, 1e-8, 12345);
      Vector t(N), s_ref(N), s2(N);
      t.Randomize(11);
      S.Mult(t, s_ref);
      S2.Mult(t, s2);
      check("OperatorChebyshevSmoother::Mult (order 3, power method)", relerr(s2, s_ref), 1e-10);
      cg.SetPreconditioner(S);
      X = 0.0;
      cg.Mult(B, X);
      b200::CGSolver cg3;
      cg3.SetRelTol(1e-6); cg3.SetAbsTol(0.0); cg3.SetMaxIter(2000);
      cg3.SetOperator(A_b2);
      cg3.SetPreconditioner(S2);
      Vector X3(N);
      X3 = 0.0;
      cg3.Mult(B2, X3);
      std::printf("  Chebyshev-PCG iterations: MFEM %d, b200 %d\n", cg.GetNumIterations(), cg3.GetNumIterations());
      check("Chebyshev-PCG iteration count |difference|", std::abs(cg.GetNumIterations() - cg3.GetNumIterations()), 1.0);
      check("Chebyshev-PCG solution", relerr(X3, X), 1e-5);
   }
   std::printf(fails ? "ex1_b200: %d check(s) FAILED\n" : "ex1_b200: all checks passed\n", fails);
   return fails ? 1 : 0;
}
