// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Driver that runs the UNMODIFIED reference (MFEM, compiled by oracle/Makefile into
// oracle/_ref/libmfem_ref.a) through its own public API on the partial-assembly hot
// path and (a) dumps every table / vector on that path as a parity fixture, or
// (b) times the reference's CPU path (cpu_baseline and `bench.py --impl reference`).
//
// It is the generator of tests/golden/*.npz (via oracle/make_golden.py) and the
// "reference" arm of bench.py.  Nothing in mfem_b200/ links or calls it.
//
// Reference entry points exercised (file:line under /root/reference):
//   Mesh::MakeCartesian2D/3D                 mesh/mesh.cpp:4776,4786
//   H1_FECollection / FiniteElementSpace     fem/fe_coll.hpp, fem/fespace.cpp:356
//   ElementRestriction (ctor, Mult, MultT)   fem/restriction.cpp:26,116,159
//   DofToQuad (TENSOR)                       fem/fe/fe_base.cpp:2697
//   GeometricFactors                         mesh/mesh.cpp:15420
//   internal::PADiffusionSetup               fem/integ/bilininteg_diffusion_kernels.cpp:107
//   Diffusion/MassIntegrator::AssemblePA,
//     AddMultPA, AssembleDiagonalPA          fem/integ/bilininteg_{diffusion,mass}_pa.cpp
//   BilinearForm (PARTIAL) Mult/AssembleDiagonal/FormLinearSystem
//                                            fem/bilinearform_ext.cpp:332,370,476,487
//   OperatorJacobiSmoother, CGSolver         linalg/solvers.cpp:324,869
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
// The PA data of the integrators is private in the reference; the dump needs to read it.
#define private public
#define protected public
#include "mfem.hpp"
#include "fem/integ/bilininteg_diffusion_kernels.hpp"
#include "fem/lor/lor_batched.hpp"
#undef private
#undef protected

using namespace mfem;
using std::string;

// ---- tiny binary container: [name:32][dtype:1 ('i','d','l')][count:int64][payload] ----
struct Dump
{
   FILE *f;
   explicit Dump(const char *path) : f(fopen(path, "wb"))
   {
      if (!f) { perror(path); exit(2); }
   }
   ~Dump() { fclose(f); }
   void put(const char *name, char dt, int64_t n, const void *p, size_t esz)
   {
      char nm[32]; memset(nm, 0, 32); strncpy(nm, name, 31);
      fwrite(nm, 1, 32, f); fwrite(&dt, 1, 1, f); fwrite(&n, 8, 1, f);
      if (n) { fwrite(p, esz, n, f); }
   }
   void ints(const char *n, const int *p, int64_t c) { put(n, 'i', c, p, 4); }
   void dbls(const char *n, const double *p, int64_t c) { put(n, 'd', c, p, 8); }
   void ints(const char *n, const Array<int> &a) { a.HostRead(); ints(n, a.GetData(), a.Size()); }
   void dbls(const char *n, const Vector &v) { v.HostRead(); dbls(n, v.GetData(), v.Size()); }
   void dbls(const char *n, const Array<real_t> &v) { v.HostRead(); dbls(n, v.GetData(), v.Size()); }
   void scalar(const char *n, int v) { ints(n, &v, 1); }
   void scalar(const char *n, double v) { dbls(n, &v, 1); }
};

struct Args
{
   std::map<string, string> kv;
   Args(int argc, char **argv, int first)
   {
      for (int i = first; i + 1 < argc; i += 2)
      {
         string k = argv[i];
         if (k.rfind("--", 0) == 0) { k = k.substr(2); }
         kv[k] = argv[i + 1];
      }
   }
   int geti(const string &k, int d) const
   { auto it = kv.find(k); return it == kv.end() ? d : atoi(it->second.c_str()); }
   double getd(const string &k, double d) const
   { auto it = kv.find(k); return it == kv.end() ? d : atof(it->second.c_str()); }
   string gets(const string &k, const string &d) const
   { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
};

// C5's curved mesh: Q3 nodes displaced by a smooth map.
static void curve_map(const Vector &x, Vector &y)
{
   const double a = 0.05, tp = 2 * M_PI;
   y = x;
   if (x.Size() == 3)
   {
      y(0) += a * sin(tp * x(1)) * sin(tp * x(2));
      y(1) += a * sin(tp * x(0)) * sin(tp * x(2));
      y(2) += a * sin(tp * x(0)) * sin(tp * x(1));
   }
   else
   {
      y(0) += a * sin(tp * x(1));
      y(1) += a * sin(tp * x(0));
   }
}

struct Problem
{
   int dim, n, p, nz, vdim, byvdim;
   bool mass, gll, curved, sfc;
   Mesh mesh;
   H1_FECollection *fec = nullptr;
   FiniteElementSpace *fes = nullptr;
   const IntegrationRule *ir = nullptr;
   IntegrationRules gll_rules{0, Quadrature1D::GaussLobatto};
   ConstantCoefficient one{1.0};
   BilinearForm *a = nullptr;
   BilinearFormIntegrator *bfi = nullptr;

   Problem(const Args &A)
   {
      dim = A.geti("dim", 3); n = A.geti("n", 2); p = A.geti("p", 2);
      nz = A.geti("nz", n);
      mass = A.gets("kind", "diffusion") == "mass";
      gll = A.gets("quad", "gl") == "gll";
      curved = A.geti("curved", 0) != 0;
      sfc = A.geti("sfc", 1) != 0;
      vdim = A.geti("vdim", 1);
      byvdim = A.geti("byvdim", vdim == 3 ? 1 : 0); // bench_ceed.cpp:131: byVDIM for VDIM == 3
      const int ny = A.geti("ny", n);
      string meshfile = A.gets("mesh", "");
      if (!meshfile.empty())
      {
         mesh = Mesh(meshfile.c_str(), 1, 1);
         int ref = A.geti("refine", 0);
         for (int l = 0; l < ref; l++) { mesh.UniformRefinement(); }
         dim = mesh.Dimension();
      }
      else if (dim == 3)
      {
         mesh = Mesh::MakeCartesian3D(n, ny, nz, Element::HEXAHEDRON, 1.0, 1.0, 1.0, sfc);
      }
      else
      {
         mesh = Mesh::MakeCartesian2D(n, ny, Element::QUADRILATERAL, false, 1.0, 1.0, sfc);
      }
      if (curved) { mesh.SetCurvature(A.geti("node_order", 3)); mesh.Transform(curve_map); }
      if (const int nattr = A.geti("nattr", 0))
      {
         // element attributes 1..nattr in short runs (for the marked-integrator dump)
         for (int e = 0; e < mesh.GetNE(); e++) { mesh.SetAttribute(e, 1 + (e / 3 + e / 7) % nattr); }
         mesh.SetAttributes();
      }
      fec = new H1_FECollection(p, dim);
      fes = vdim == 1 ? new FiniteElementSpace(&mesh, fec)
            : new FiniteElementSpace(&mesh, fec, vdim, byvdim ? Ordering::byVDIM : Ordering::byNODES);
      const Geometry::Type g = dim == 3 ? Geometry::CUBE : Geometry::SQUARE;
      // bench_ceed.cpp:121,133-137 : GL order 2p+3 (Q=p+2) or GLL order 2p-1 (Q=p+1)
      const int q_extra = A.geti("qorder", -1);
      if (gll) { ir = &gll_rules.Get(g, q_extra >= 0 ? q_extra : 2 * p - 1); }
      else { ir = &IntRules.Get(g, q_extra >= 0 ? q_extra : 2 * p + 3); }
      a = new BilinearForm(fes);
      a->SetAssemblyLevel(AssemblyLevel::PARTIAL);
      if (vdim > 1)
      {
         // CEED BP2 / BP4 / BP6 (bench_ceed.cpp: VectorMassIntegrator / VectorDiffusionIntegrator)
         if (mass) { auto v = new VectorMassIntegrator(one); v->SetIntRule(ir); bfi = v; }
         else { bfi = new VectorDiffusionIntegrator(one, ir); }
      }
      else if (mass) { bfi = new MassIntegrator(one, ir); }
      else { bfi = new DiffusionIntegrator(one, ir); }
      a->AddDomainIntegrator(bfi);
   }
   ~Problem() { delete a; delete fes; delete fec; }
};

static double now()
{
   return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int do_dump(const Args &A, const char *out)
{
   Problem P(A);
   Dump D(out);
   P.a->Assemble();
   FiniteElementSpace &fes = *P.fes;
   const FiniteElement &fe = *fes.GetTypicalFE();
   const DofToQuad &maps = fe.GetDofToQuad(*P.ir, DofToQuad::TENSOR);
   const int D1D = maps.ndof, Q1D = maps.nqpt, NE = fes.GetNE(), ndofs = fes.GetNDofs();
   const int dim = P.dim;
   D.scalar("dim", dim); D.scalar("n", P.n); D.scalar("p", P.p);
   D.scalar("D1D", D1D); D.scalar("Q1D", Q1D); D.scalar("NE", NE); D.scalar("ndofs", ndofs);
   D.scalar("is_mass", (int)P.mass); D.scalar("is_gll", (int)P.gll); D.scalar("curved", (int)P.curved);

   // --- a1: restriction tables
   auto R = dynamic_cast<const ElementRestriction *>(
               fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
   const Table &e2d = fes.GetElementToDofTable();
   D.ints("elem_dof", e2d.GetJ(), (int64_t)NE * fe.GetDof());
   auto tfe = dynamic_cast<const TensorBasisElement *>(&fe);
   D.ints("dof_map", tfe->GetDofMap());
   D.ints("gather_map", R->GatherMap());
   D.ints("offsets", R->Offsets());
   D.ints("indices", R->Indices());

   // --- a12: basis tables
   D.dbls("B", maps.B); D.dbls("G", maps.G); D.dbls("Bt", maps.Bt); D.dbls("Gt", maps.Gt);
   D.dbls("W", P.ir->GetWeights());
   {
      // 1D points of the dof basis and the quadrature rule (for checking our generators)
      const IntegrationRule &ir1 = (P.gll ? P.gll_rules : IntRules).Get(Geometry::SEGMENT, P.ir->GetOrder());
      Vector qp(ir1.GetNPoints()), qw(ir1.GetNPoints());
      for (int i = 0; i < ir1.GetNPoints(); i++) { qp(i) = ir1.IntPoint(i).x; qw(i) = ir1.IntPoint(i).weight; }
      D.dbls("qpts1d", qp); D.dbls("qwts1d", qw);
      const double *gl = poly1d.ClosedPoints(P.p, BasisType::GaussLobatto);
      D.dbls("dofpts1d", gl, P.p + 1);
   }

   // --- a21: mesh nodes and geometric factors
   P.mesh.EnsureNodes();
   const GridFunction *nodes = P.mesh.GetNodes();
   const FiniteElementSpace *nfes = nodes->FESpace();
   const Operator *nR = nfes->GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC);
   Vector nodesE(nR->Height());
   nR->Mult(*nodes, nodesE);
   const int nodeD1D = nfes->GetTypicalFE()->GetOrder() + 1;
   D.scalar("node_D1D", nodeD1D);
   D.dbls("nodesE", nodesE); // (nodeD1D^dim, dim, NE)
   {
      const DofToQuad &nm = nfes->GetTypicalFE()->GetDofToQuad(*P.ir, DofToQuad::TENSOR);
      D.dbls("Bn", nm.B); D.dbls("Gn", nm.G);
   }
   const GeometricFactors *geom = P.mesh.GetGeometricFactors(
                                     *P.ir, GeometricFactors::JACOBIANS | GeometricFactors::DETERMINANTS);
   D.dbls("J", geom->J); D.dbls("detJ", geom->detJ);

   // --- a6/a9: quadrature data
   Vector pa_data;
   if (P.mass) { pa_data = static_cast<MassIntegrator *>(P.bfi)->pa_data; }
   else { pa_data = static_cast<DiffusionIntegrator *>(P.bfi)->pa_data; }
   D.dbls("pa_data", pa_data);
   if (!P.mass)
   {
      // the seam itself, with a constant and a spatially varying matrix coefficient
      const int NQ = P.ir->GetNPoints();
      Vector c1(1); c1 = 1.0;
      Vector d2(NQ * (dim == 3 ? 6 : 3) * NE);
      internal::PADiffusionSetup(dim, dim, D1D, Q1D, 1, NE, P.ir->GetWeights(), geom->J, c1, d2);
      D.dbls("pa_data_seam", d2);
      const int cd = dim == 3 ? 6 : 3;
      Vector cm(cd * NQ * NE);
      srand(7);
      for (int i = 0; i < NQ * NE; i++)
      {
         // SPD-ish symmetric matrix coefficient, storage (11,12,13,22,23,33) / (11,12,22)
         double r[6]; for (int k = 0; k < 6; k++) { r[k] = rand() / (RAND_MAX + 1.0); }
         if (dim == 3)
         {
            cm(0 + cd * i) = 2 + r[0]; cm(1 + cd * i) = 0.3 * r[1]; cm(2 + cd * i) = 0.3 * r[2];
            cm(3 + cd * i) = 2 + r[3]; cm(4 + cd * i) = 0.3 * r[4]; cm(5 + cd * i) = 2 + r[5];
         }
         else
         {
            cm(0 + cd * i) = 2 + r[0]; cm(1 + cd * i) = 0.3 * r[1]; cm(2 + cd * i) = 2 + r[3];
         }
      }
      Vector d3(NQ * cd * NE);
      internal::PADiffusionSetup(dim, dim, D1D, Q1D, cd, NE, P.ir->GetWeights(), geom->J, cm, d3);
      D.dbls("coeff_matrix", cm); D.dbls("pa_data_matrix", d3);
      // vector (diagonal) coefficient
      Vector cv(dim * NQ * NE);
      for (int i = 0; i < cv.Size(); i++) { cv(i) = 1 + rand() / (RAND_MAX + 1.0); }
      Vector d4(NQ * cd * NE);
      internal::PADiffusionSetup(dim, dim, D1D, Q1D, dim, NE, P.ir->GetWeights(), geom->J, cv, d4);
      D.dbls("coeff_vector", cv); D.dbls("pa_data_vector", d4);
   }

   // --- a2/a3/a7/a10: operator action pieces on x = Randomize(1)
   Vector x(ndofs); x.Randomize(1);
   D.dbls("x", x);
   Vector xE(R->Height()), yE(R->Height());
   R->Mult(x, xE);
   D.dbls("xE", xE);
   yE = 0.0;
   P.bfi->AddMultPA(xE, yE);
   D.dbls("yE", yE);
   Vector yRt(ndofs);
   R->MultTranspose(yE, yRt);
   D.dbls("y_RtyE", yRt);
   Vector y(ndofs);
   P.a->Mult(x, y);
   D.dbls("y", y);
   D.scalar("y_norml2", y.Norml2());
   {
      Vector t(ndofs);
      R->AbsMultTranspose(xE, t);
      D.dbls("absmultT_xE", t);
      Vector t2(ndofs); t2.Randomize(3);
      R->AddMultTranspose(xE, t2);
      D.dbls("addmultT_xE_seed3", t2);
   }

   // --- a8/a11: diagonal
   Vector diagE(R->Height()); diagE = 0.0;
   P.bfi->AssembleDiagonalPA(diagE);
   D.dbls("diagE", diagE);
   Vector diag(ndofs);
   P.a->AssembleDiagonal(diag);
   D.dbls("diag", diag);

   // --- a14-a17: constrained system + PCG/Jacobi  (ex1.cpp:165-267)
   Array<int> ess_bdr(P.mesh.bdr_attributes.Max()); ess_bdr = 1;
   Array<int> ess;
   fes.GetEssentialTrueDofs(ess_bdr, ess);
   D.ints("ess_tdof_list", ess);
   LinearForm b(&fes);
   b.AddDomainIntegrator(new DomainLFIntegrator(P.one));
   b.Assemble();
   D.dbls("b", b);
   GridFunction u(&fes); u = 0.0;
   OperatorPtr Aop; Vector B, X;
   P.a->FormLinearSystem(ess, u, b, Aop, X, B);
   D.dbls("B_rhs", B);
   {
      Vector cy(ndofs);
      Aop->Mult(x, cy); // ConstrainedOperator::Mult on the random vector
      D.dbls("y_constrained", cy);
   }
   const double rtol = A.getd("rtol", 1e-8);
   const int maxit = A.geti("maxit", 2000);
   const int use_prec = A.geti("jacobi", 1);
   OperatorJacobiSmoother M(*P.a, ess);
   {
      Vector t(ndofs), jr(ndofs);
      t.Randomize(5);
      M.Mult(t, jr);
      D.dbls("jacobi_in_seed5", t); D.dbls("jacobi_out", jr);
   }
   CGSolver cg;
   cg.SetRelTol(rtol); cg.SetAbsTol(0.0); cg.SetMaxIter(maxit);
   cg.SetOperator(*Aop);
   if (use_prec) { cg.SetPreconditioner(M); }
   // capture the (B r, r) history through the print stream
   std::stringstream hist;
   std::streambuf *old = mfem::out.rdbuf(hist.rdbuf());
   cg.SetPrintLevel(1);
   X = 0.0;
   cg.Mult(B, X);
   mfem::out.rdbuf(old);
   std::vector<double> h;
   {
      string line;
      while (std::getline(hist, line))
      {
         size_t pos = line.find("(B r, r) =");
         if (pos != string::npos) { h.push_back(atof(line.c_str() + pos + 10)); }
      }
   }
   D.dbls("pcg_history", h.data(), h.size());
   D.scalar("pcg_iters", cg.GetNumIterations());
   D.scalar("pcg_converged", (int)cg.GetConverged());
   D.scalar("pcg_final_norm", (double)cg.GetFinalNorm());
   D.scalar("pcg_rtol", rtol);
   D.scalar("pcg_jacobi", use_prec);
   D.dbls("X_sol", X);
   std::printf("dump %s: dim=%d NE=%d p=%d D1D=%d Q1D=%d ndofs=%d |y|=%.17g pcg_iters=%d\n",
               out, dim, NE, P.p, D1D, Q1D, ndofs, y.Norml2(), cg.GetNumIterations());
   return 0;
}

// Vector problems (vdim > 1): only what differs from the scalar dump -- the vdof layout, the
// operator action, the diagonal and the constrained solve.
static int do_dump_vector(const Args &A, const char *out)
{
   Problem P(A);
   Dump D(out);
   P.a->Assemble();
   FiniteElementSpace &fes = *P.fes;
   const FiniteElement &fe = *fes.GetTypicalFE();
   const DofToQuad &maps = fe.GetDofToQuad(*P.ir, DofToQuad::TENSOR);
   const int D1D = maps.ndof, Q1D = maps.nqpt, NE = fes.GetNE(), ndofs = fes.GetNDofs();
   const int nv = fes.GetVSize();
   D.scalar("dim", P.dim); D.scalar("n", P.n); D.scalar("p", P.p);
   D.scalar("D1D", D1D); D.scalar("Q1D", Q1D); D.scalar("NE", NE); D.scalar("ndofs", ndofs);
   D.scalar("vdim", P.vdim); D.scalar("byvdim", P.byvdim);
   D.scalar("is_mass", (int)P.mass); D.scalar("is_gll", (int)P.gll); D.scalar("curved", (int)P.curved);
   auto R = dynamic_cast<const ElementRestriction *>(
               fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
   D.ints("gather_map", R->GatherMap());
   D.ints("offsets", R->Offsets());
   D.ints("indices", R->Indices());
   D.dbls("B", maps.B); D.dbls("G", maps.G);
   D.dbls("W", P.ir->GetWeights());
   const GeometricFactors *geom = P.mesh.GetGeometricFactors(
                                     *P.ir, GeometricFactors::JACOBIANS | GeometricFactors::DETERMINANTS);
   D.dbls("J", geom->J); D.dbls("detJ", geom->detJ);
   Vector x(nv); x.Randomize(1);
   D.dbls("x", x);
   Vector xE(R->Height()), yE(R->Height());
   R->Mult(x, xE);
   D.dbls("xE", xE);
   yE = 0.0;
   P.bfi->AddMultPA(xE, yE);
   D.dbls("yE", yE);
   Vector y(nv);
   P.a->Mult(x, y);
   D.dbls("y", y);
   Vector diag(nv);
   P.a->AssembleDiagonal(diag);
   D.dbls("diag", diag);
   Array<int> ess_bdr(P.mesh.bdr_attributes.Max()); ess_bdr = 1;
   Array<int> ess;
   fes.GetEssentialTrueDofs(ess_bdr, ess);
   D.ints("ess_tdof_list", ess);
   Vector b(nv); b.Randomize(2); // bench_ceed uses a vector load; any RHS pins the solver
   GridFunction u(&fes); u = 0.0;
   OperatorPtr Aop; Vector B, X;
   D.dbls("b", b);
   P.a->FormLinearSystem(ess, u, b, Aop, X, B);
   D.dbls("B_rhs", B);
   {
      Vector cy(nv);
      Aop->Mult(x, cy);
      D.dbls("y_constrained", cy);
   }
   const double rtol = A.getd("rtol", 1e-8);
   OperatorJacobiSmoother M(*P.a, ess);
   CGSolver cg;
   cg.SetRelTol(rtol); cg.SetAbsTol(0.0); cg.SetMaxIter(2000); cg.SetPrintLevel(-1);
   cg.SetOperator(*Aop);
   cg.SetPreconditioner(M);
   X = 0.0;
   cg.Mult(B, X);
   D.scalar("pcg_iters", cg.GetNumIterations());
   D.scalar("pcg_converged", (int)cg.GetConverged());
   D.scalar("pcg_rtol", rtol);
   D.dbls("X_sol", X);
   std::printf("dump %s: vdim=%d byvdim=%d NE=%d p=%d ndofs=%d |y|=%.17g pcg_iters=%d\n", out, P.vdim,
               P.byvdim, NE, P.p, ndofs, y.Norml2(), cg.GetNumIterations());
   return 0;
}

// Non-symmetric matrix coefficient (DiffusionIntegrator(MatrixCoefficient&), symmetric = false:
// dim*dim quadrature-data components, fem/integ/bilininteg_diffusion_pa.cpp:165).
static void nonsym_coeff(const Vector &x, DenseMatrix &K)
{
   const int d = x.Size();
   K.SetSize(d);
   for (int i = 0; i < d; i++)
      for (int j = 0; j < d; j++)
      {
         K(i, j) = (i == j ? 2.0 + 0.5 * x(i) : 0.0) + 0.3 * (i + 1) * x((i + j) % d) - 0.2 * (j + 1) * x(j);
      }
}

static int do_dump_nonsym(const Args &A, const char *out)
{
   Problem P(A);
   Dump D(out);
   const int dim = P.dim;
   MatrixFunctionCoefficient MQ(dim, nonsym_coeff);
   BilinearForm a(P.fes);
   a.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   auto *bfi = new DiffusionIntegrator(MQ, P.ir);
   a.AddDomainIntegrator(bfi);
   a.Assemble();
   FiniteElementSpace &fes = *P.fes;
   const FiniteElement &fe = *fes.GetTypicalFE();
   const DofToQuad &maps = fe.GetDofToQuad(*P.ir, DofToQuad::TENSOR);
   const int D1D = maps.ndof, Q1D = maps.nqpt, NE = fes.GetNE(), ndofs = fes.GetNDofs();
   D.scalar("dim", dim); D.scalar("n", P.n); D.scalar("p", P.p);
   D.scalar("D1D", D1D); D.scalar("Q1D", Q1D); D.scalar("NE", NE); D.scalar("ndofs", ndofs);
   D.scalar("curved", (int)P.curved);
   auto R = dynamic_cast<const ElementRestriction *>(
               fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
   D.ints("gather_map", R->GatherMap());
   D.ints("offsets", R->Offsets());
   D.ints("indices", R->Indices());
   D.dbls("B", maps.B); D.dbls("G", maps.G);
   D.dbls("W", P.ir->GetWeights());
   const GeometricFactors *geom = P.mesh.GetGeometricFactors(*P.ir, GeometricFactors::JACOBIANS);
   D.dbls("J", geom->J);
   // the coefficient at the quadrature points, as AssemblePA evaluates it (dim*dim, NQ, NE)
   QuadratureSpace qs(P.mesh, *P.ir);
   CoefficientVector cv(qs, CoefficientStorage::FULL);
   cv.ProjectTranspose(MQ); // what AssemblePA does (bilininteg_diffusion_pa.cpp:159): row-major entries
   D.dbls("coeff_full", cv);
   D.scalar("pa_symmetric", (int)bfi->symmetric);
   D.dbls("pa_data", bfi->pa_data);
   Vector x(ndofs); x.Randomize(1);
   D.dbls("x", x);
   Vector xE(R->Height()), yE(R->Height());
   R->Mult(x, xE);
   D.dbls("xE", xE);
   yE = 0.0;
   bfi->AddMultPA(xE, yE);
   D.dbls("yE", yE);
   Vector y(ndofs);
   a.Mult(x, y);
   D.dbls("y", y);
   Vector diagE(R->Height()); diagE = 0.0;
   bfi->AssembleDiagonalPA(diagE);
   D.dbls("diagE", diagE);
   Vector diag(ndofs);
   a.AssembleDiagonal(diag);
   D.dbls("diag", diag);
   std::printf("dump %s: nonsymmetric coefficient dim=%d NE=%d p=%d symmetric=%d |y|=%.17g\n", out, dim, NE,
               P.p, (int)bfi->symmetric, y.Norml2());
   return 0;
}

// Unstructured / curved meshes from the reference's data directory (inline-hex = BASELINE
// config C1, toroid-hex, fichera-q2, star-q3): a lean dump -- the reference's own tables, the
// nodes E-vector and the results of the path, without the bulky geometric factors.
static int do_dump_mesh(const Args &A, const char *out)
{
   Problem P(A);
   Dump D(out);
   P.a->Assemble();
   FiniteElementSpace &fes = *P.fes;
   const FiniteElement &fe = *fes.GetTypicalFE();
   const DofToQuad &maps = fe.GetDofToQuad(*P.ir, DofToQuad::TENSOR);
   const int D1D = maps.ndof, Q1D = maps.nqpt, NE = fes.GetNE(), ndofs = fes.GetNDofs();
   const int dim = P.dim;
   D.scalar("dim", dim); D.scalar("p", P.p);
   D.scalar("D1D", D1D); D.scalar("Q1D", Q1D); D.scalar("NE", NE); D.scalar("ndofs", ndofs);
   D.scalar("is_mass", (int)P.mass);
   auto R = dynamic_cast<const ElementRestriction *>(
               fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
   D.ints("gather_map", R->GatherMap());
   D.ints("offsets", R->Offsets());
   D.ints("indices", R->Indices());
   D.dbls("B", maps.B); D.dbls("G", maps.G);
   D.dbls("W", P.ir->GetWeights());
   P.mesh.EnsureNodes();
   const GridFunction *nodes = P.mesh.GetNodes();
   const FiniteElementSpace *nfes = nodes->FESpace();
   const Operator *nR = nfes->GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC);
   Vector nodesE(nR->Height());
   nR->Mult(*nodes, nodesE);
   D.scalar("node_D1D", nfes->GetTypicalFE()->GetOrder() + 1);
   D.dbls("nodesE", nodesE);
   {
      const DofToQuad &nm = nfes->GetTypicalFE()->GetDofToQuad(*P.ir, DofToQuad::TENSOR);
      D.dbls("Bn", nm.B); D.dbls("Gn", nm.G);
   }
   Vector x(ndofs); x.Randomize(1);
   D.dbls("x", x);
   Vector y(ndofs);
   P.a->Mult(x, y);
   D.dbls("y", y);
   Vector diag(ndofs);
   P.a->AssembleDiagonal(diag);
   D.dbls("diag", diag);
   Array<int> ess_bdr(P.mesh.bdr_attributes.Max()); ess_bdr = 1;
   Array<int> ess;
   fes.GetEssentialTrueDofs(ess_bdr, ess);
   D.ints("ess_tdof_list", ess);
   LinearForm b(&fes);
   b.AddDomainIntegrator(new DomainLFIntegrator(P.one));
   b.Assemble();
   D.dbls("b", b);
   GridFunction u(&fes); u = 0.0;
   OperatorPtr Aop; Vector B, X;
   P.a->FormLinearSystem(ess, u, b, Aop, X, B);
   D.dbls("B_rhs", B);
   Vector cy(ndofs);
   Aop->Mult(x, cy);
   D.dbls("y_constrained", cy);
   {
      // Operator::AbsMult: the extension's MultInternal(useAbs) behind ConstrainedOperator::AbsMult
      // (with and without constraints), and the integrator's AddAbsMultPA on the E-vector
      Vector xs(x);
      for (int i = 0; i < ndofs; i++) { xs(i) -= 0.5; } // signed input
      D.dbls("x_signed", xs);
      Vector ya(ndofs);
      Aop->AbsMult(xs, ya);
      D.dbls("y_abs_constrained", ya);
      Array<int> no_ess;
      OperatorPtr A0; Vector B0, X0, b0(b);
      GridFunction u0(&fes); u0 = 0.0;
      P.a->FormLinearSystem(no_ess, u0, b0, A0, X0, B0);
      A0->AbsMult(xs, ya);
      D.dbls("y_abs", ya);
      Vector xE(R->Height()), yE(R->Height());
      R->Mult(xs, xE);
      yE = 0.0;
      P.bfi->AddAbsMultPA(xE, yE);
      D.dbls("yE_abs", yE); // input: R applied to x_signed (bit-exact gather, not dumped)
      // the other diagonal policy of the ConstrainedOperator, and its AssembleDiagonal
      ConstrainedOperator Cz(P.a, ess, false, Operator::DIAG_ZERO);
      Cz.Mult(x, ya);
      D.dbls("y_constrained_diag_zero", ya);
      Cz.AssembleDiagonal(ya);
      D.dbls("diag_constrained_diag_zero", ya);
      ConstrainedOperator C1(P.a, ess, false, Operator::DIAG_ONE);
      C1.AssembleDiagonal(ya);
      D.dbls("diag_constrained_diag_one", ya);
   }
   // ex1.cpp:266-267: PCG(*A, M, B, X, 1, 400, 1e-12, 0.0) -- rel_tol^2 = 1e-12
   const double rtol = A.getd("rtol", 1e-6);
   OperatorJacobiSmoother M(*P.a, ess);
   CGSolver cg;
   cg.SetRelTol(rtol); cg.SetAbsTol(0.0); cg.SetMaxIter(A.geti("maxit", 400));
   cg.SetOperator(*Aop);
   cg.SetPreconditioner(M);
   std::stringstream hist;
   std::streambuf *old = mfem::out.rdbuf(hist.rdbuf());
   cg.SetPrintLevel(1);
   X = 0.0;
   cg.Mult(B, X);
   mfem::out.rdbuf(old);
   std::vector<double> h;
   {
      string line;
      while (std::getline(hist, line))
      {
         size_t pos = line.find("(B r, r) =");
         if (pos != string::npos) { h.push_back(atof(line.c_str() + pos + 10)); }
      }
   }
   D.dbls("pcg_history", h.data(), h.size());
   D.scalar("pcg_iters", cg.GetNumIterations());
   D.scalar("pcg_converged", (int)cg.GetConverged());
   D.scalar("pcg_rtol", rtol);
   D.dbls("X_sol", X);
   std::printf("dump %s: mesh file dim=%d NE=%d p=%d ndofs=%d node order %d |y|=%.17g pcg_iters=%d\n", out, dim, NE,
               P.p, ndofs, nfes->GetTypicalFE()->GetOrder(), y.Norml2(), cg.GetNumIterations());
   return 0;
}

// DomainLFIntegrator / VectorDomainLFIntegrator fast assembly (LinearForm::UseFastAssembly,
// fem/linearform_ext.cpp:20-66; fem/integ/lininteg_domain.cpp, lininteg_domain_kernels.hpp)
static double lf_coeff(const Vector &x)
{
   double s = 1.0 + x(0) + 2.0 * x(1) * x(1);
   if (x.Size() > 2) { s += std::sin(3.0 * x(2)); }
   return s;
}
static void lf_vcoeff(const Vector &x, Vector &v)
{
   for (int c = 0; c < v.Size(); c++) { v(c) = lf_coeff(x) * (c + 1) + 0.25 * c * x(c % x.Size()); }
}
static int do_dump_lf(const Args &A, const char *out)
{
   Problem P(A);
   Dump D(out);
   const int dim = P.dim, vdim = P.vdim;
   FiniteElementSpace &fes = *P.fes;
   const FiniteElement &fe = *fes.GetTypicalFE();
   const DofToQuad &maps = fe.GetDofToQuad(*P.ir, DofToQuad::TENSOR);
   const int D1D = maps.ndof, Q1D = maps.nqpt, NE = fes.GetNE(), ndofs = fes.GetNDofs();
   D.scalar("dim", dim); D.scalar("n", P.n); D.scalar("p", P.p); D.scalar("vdim", vdim);
   D.scalar("byvdim", P.byvdim);
   D.scalar("D1D", D1D); D.scalar("Q1D", Q1D); D.scalar("NE", NE); D.scalar("ndofs", ndofs);
   auto R = dynamic_cast<const ElementRestriction *>(
               fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
   D.ints("gather_map", R->GatherMap());
   D.ints("offsets", R->Offsets());
   D.ints("indices", R->Indices());
   D.dbls("B", maps.B);
   D.dbls("W", P.ir->GetWeights());
   const GeometricFactors *geom = P.mesh.GetGeometricFactors(*P.ir, GeometricFactors::DETERMINANTS);
   D.dbls("detJ", geom->detJ);
   FunctionCoefficient f(lf_coeff);
   VectorFunctionCoefficient vf(vdim, lf_vcoeff);
   QuadratureSpace qs(P.mesh, *P.ir);
   std::unique_ptr<LinearFormIntegrator> integ, integ2;
   if (vdim == 1)
   {
      CoefficientVector cv(f, qs, CoefficientStorage::COMPRESSED);
      D.dbls("coeff", cv);
      integ.reset(new DomainLFIntegrator(f, P.ir));
      integ2.reset(new DomainLFIntegrator(f, P.ir));
   }
   else
   {
      CoefficientVector cv(vf, qs, CoefficientStorage::COMPRESSED);
      D.dbls("coeff", cv);
      integ.reset(new VectorDomainLFIntegrator(vf, P.ir));
      integ2.reset(new VectorDomainLFIntegrator(vf, P.ir));
   }
   // the integrator's device kernel on an E-vector, with element markers
   const Array<int> &attr = P.mesh.GetElementAttributes();
   D.ints("elem_attributes", attr);
   Array<int> markers(NE);
   for (int e = 0; e < NE; e++) { markers[e] = (attr[e] % 2 == 1); }
   D.ints("markers", markers);
   Vector be(R->Height());
   be = 0.0;
   integ->AssembleDevice(fes, markers, be);
   D.dbls("b_e_marked", be);
   markers = 1;
   be = 0.0;
   integ->AssembleDevice(fes, markers, be);
   D.dbls("b_e", be);
   // constant coefficient (CoefficientVector of size vdim)
   {
      ConstantCoefficient c3(3.0);
      Vector cvec(vdim);
      for (int c = 0; c < vdim; c++) { cvec(c) = 1.5 + c; }
      VectorConstantCoefficient vc3(cvec);
      std::unique_ptr<LinearFormIntegrator> ic;
      if (vdim == 1) { ic.reset(new DomainLFIntegrator(c3, P.ir)); }
      else { ic.reset(new VectorDomainLFIntegrator(vc3, P.ir)); }
      be = 0.0;
      ic->AssembleDevice(fes, markers, be);
      D.dbls("b_e_const", be);
   }
   // the whole LinearForm: fast assembly and the legacy element loop
   LinearForm lf(&fes);
   lf.AddDomainIntegrator(integ.release());
   lf.UseFastAssembly(true);
   lf.Assemble();
   D.dbls("b_fast", lf);
   LinearForm lf2(&fes);
   lf2.AddDomainIntegrator(integ2.release());
   lf2.Assemble();
   D.dbls("b_legacy", lf2);
   Vector diff(lf); diff -= lf2;
   std::printf("dump %s: linear form dim=%d vdim=%d NE=%d p=%d |b|=%.17g fast-legacy=%.3e\n", out, dim, vdim, NE,
               P.p, lf.Norml2(), diff.Norml2() / lf2.Norml2());
   return 0;
}

// BilinearForm::AddDomainIntegrator(bfi, elem_marker): integrators restricted by element
// attribute (PABilinearFormExtension::AddMultWithMarkers / AssembleDiagonal,
// fem/bilinearform_ext.cpp:370-454,753-847).  Three integrators on a mesh with 3 attributes:
// diffusion on attributes {1,3}, mass on {2}, mass (coefficient 0.5) unrestricted.
static int do_dump_markers(const Args &A, const char *out)
{
   Problem P(A);
   Dump D(out);
   const int dim = P.dim;
   FiniteElementSpace &fes = *P.fes;
   const FiniteElement &fe = *fes.GetTypicalFE();
   const DofToQuad &maps = fe.GetDofToQuad(*P.ir, DofToQuad::TENSOR);
   const int D1D = maps.ndof, Q1D = maps.nqpt, NE = fes.GetNE(), ndofs = fes.GetNDofs();
   const int nattr = P.mesh.attributes.Max();
   Array<int> m_diff(nattr), m_mass(nattr);
   for (int k = 0; k < nattr; k++) { m_diff[k] = (k % 2 == 0); m_mass[k] = (k % 2 == 1); }
   ConstantCoefficient half(0.5);
   BilinearForm a(&fes);
   a.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   auto *bd = new DiffusionIntegrator(P.one, P.ir);
   auto *bm = new MassIntegrator(P.one, P.ir);
   auto *bh = new MassIntegrator(half, P.ir);
   const int order = A.geti("marker_order", 0); // 0: marked, marked, plain; 1: plain first
   if (order == 1) { a.AddDomainIntegrator(bh); }
   a.AddDomainIntegrator(bd, m_diff);
   a.AddDomainIntegrator(bm, m_mass);
   if (order == 0) { a.AddDomainIntegrator(bh); }
   a.Assemble();
   D.scalar("dim", dim); D.scalar("n", P.n); D.scalar("p", P.p);
   D.scalar("D1D", D1D); D.scalar("Q1D", Q1D); D.scalar("NE", NE); D.scalar("ndofs", ndofs);
   D.scalar("marker_order", order);
   auto R = dynamic_cast<const ElementRestriction *>(
               fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
   D.ints("gather_map", R->GatherMap());
   D.ints("offsets", R->Offsets());
   D.ints("indices", R->Indices());
   D.dbls("B", maps.B); D.dbls("G", maps.G);
   D.dbls("W", P.ir->GetWeights());
   {
      const GeometricFactors *geom = P.mesh.GetGeometricFactors(
                                        *P.ir, GeometricFactors::JACOBIANS | GeometricFactors::DETERMINANTS);
      D.dbls("J", geom->J); D.dbls("detJ", geom->detJ);
   }
   D.ints("elem_attributes", P.mesh.GetElementAttributes());
   D.ints("marker_diffusion", m_diff); D.ints("marker_mass", m_mass);
   D.dbls("pa_data_diffusion", bd->pa_data);
   D.dbls("pa_data_mass", bm->pa_data);
   D.dbls("pa_data_mass_half", bh->pa_data);
   Vector x(ndofs); x.Randomize(1);
   D.dbls("x", x);
   Vector y(ndofs);
   a.Mult(x, y);
   D.dbls("y", y);
   Vector diag(ndofs);
   a.AssembleDiagonal(diag);
   D.dbls("diag", diag);
   Array<int> ess_bdr(P.mesh.bdr_attributes.Max()); ess_bdr = 1;
   Array<int> ess;
   fes.GetEssentialTrueDofs(ess_bdr, ess);
   D.ints("ess_tdof_list", ess);
   LinearForm b(&fes);
   b.AddDomainIntegrator(new DomainLFIntegrator(P.one));
   b.Assemble();
   D.dbls("b", b);
   GridFunction u(&fes); u = 0.0;
   OperatorPtr Aop; Vector B, X;
   a.FormLinearSystem(ess, u, b, Aop, X, B);
   D.dbls("B_rhs", B);
   Vector cy(ndofs);
   Aop->Mult(x, cy);
   D.dbls("y_constrained", cy);
   // (with the plain integrator first the reference's marked diagonal has zero entries -- the
   // later integrators reset what was accumulated on their unmarked elements -- and
   // OperatorJacobiSmoother aborts; that case is solved without a preconditioner)
   const int use_prec = A.geti("jacobi", 1);
   std::unique_ptr<OperatorJacobiSmoother> M;
   if (use_prec) { M.reset(new OperatorJacobiSmoother(a, ess)); }
   const double rtol = A.getd("rtol", 1e-8);
   CGSolver cg;
   cg.SetRelTol(rtol); cg.SetAbsTol(0.0); cg.SetMaxIter(A.geti("maxit", 2000));
   cg.SetOperator(*Aop);
   if (use_prec) { cg.SetPreconditioner(*M); }
   cg.SetPrintLevel(-1);
   X = 0.0;
   cg.Mult(B, X);
   D.scalar("pcg_iters", cg.GetNumIterations());
   D.scalar("pcg_converged", (int)cg.GetConverged());
   D.scalar("pcg_rtol", rtol);
   D.scalar("pcg_jacobi", use_prec);
   D.dbls("X_sol", X);
   std::printf("dump %s: markers dim=%d NE=%d p=%d nattr=%d |y|=%.17g pcg_iters=%d\n", out, dim, NE, P.p, nattr,
               y.Norml2(), cg.GetNumIterations());
   return 0;
}

// OperatorChebyshevSmoother (linalg/solvers.cpp:453-658) on the constrained PA operator:
// coefficients and action for every order, the power-method eigenvalue estimate, and CG
// preconditioned with it.
static int do_dump_cheb(const Args &A, const char *out)
{
   Problem P(A);
   Dump D(out);
   const int dim = P.dim;
   P.a->Assemble();
   FiniteElementSpace &fes = *P.fes;
   const FiniteElement &fe = *fes.GetTypicalFE();
   const DofToQuad &maps = fe.GetDofToQuad(*P.ir, DofToQuad::TENSOR);
   const int D1D = maps.ndof, Q1D = maps.nqpt, NE = fes.GetNE(), ndofs = fes.GetNDofs();
   D.scalar("dim", dim); D.scalar("n", P.n); D.scalar("p", P.p);
   D.scalar("D1D", D1D); D.scalar("Q1D", Q1D); D.scalar("NE", NE); D.scalar("ndofs", ndofs);
   D.scalar("is_mass", (int)P.mass);
   auto R = dynamic_cast<const ElementRestriction *>(
               fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
   D.ints("gather_map", R->GatherMap());
   D.ints("offsets", R->Offsets());
   D.ints("indices", R->Indices());
   D.dbls("B", maps.B); D.dbls("G", maps.G);
   Vector pa_data;
   if (P.mass) { pa_data = static_cast<MassIntegrator *>(P.bfi)->pa_data; }
   else { pa_data = static_cast<DiffusionIntegrator *>(P.bfi)->pa_data; }
   D.dbls("pa_data", pa_data);
   D.dbls("W", P.ir->GetWeights());
   {
      const GeometricFactors *geom = P.mesh.GetGeometricFactors(
                                        *P.ir, GeometricFactors::JACOBIANS | GeometricFactors::DETERMINANTS);
      D.dbls("J", geom->J); D.dbls("detJ", geom->detJ);
   }
   Array<int> ess_bdr(P.mesh.bdr_attributes.Max()); ess_bdr = 1;
   Array<int> ess;
   fes.GetEssentialTrueDofs(ess_bdr, ess);
   D.ints("ess_tdof_list", ess);
   LinearForm b(&fes);
   b.AddDomainIntegrator(new DomainLFIntegrator(P.one));
   b.Assemble();
   D.dbls("b", b);
   GridFunction u(&fes); u = 0.0;
   OperatorPtr Aop; Vector B, X;
   P.a->FormLinearSystem(ess, u, b, Aop, X, B);
   D.dbls("B_rhs", B);
   Vector diag(ndofs);
   P.a->AssembleDiagonal(diag);
   D.dbls("diag", diag);
   // power method, as the smoother's second constructor runs it (solvers.cpp:496-513)
   const int seed = A.geti("power_seed", 12345), pit = A.geti("power_iterations", 10);
   const double ptol = A.getd("power_tolerance", 1e-8);
   double lam;
   {
      OperatorJacobiSmoother invDiag(diag, ess, 1.0);
      ProductOperator dp(&invDiag, Aop.Ptr(), false, false);
      PowerMethod pm;
      Vector ev(ndofs);
      ev.Randomize(seed);
      D.dbls("power_start", ev);
      lam = pm.EstimateLargestEigenvalue(dp, ev, pit, ptol, seed);
      D.dbls("power_vector", ev);
   }
   D.scalar("power_seed", seed); D.scalar("power_iterations", pit); D.scalar("power_tolerance", ptol);
   D.scalar("max_eig_estimate", lam);
   Vector t(ndofs); t.Randomize(11);
   D.dbls("cheb_in_seed11", t);
   for (int order = 1; order <= 5; order++)
   {
      OperatorChebyshevSmoother S(*Aop, diag, ess, order, lam);
      Vector y(ndofs);
      S.Mult(t, y);
      char nm[64];
      std::snprintf(nm, sizeof nm, "cheb_out_order%d", order);
      D.dbls(nm, y);
   }
   // CG preconditioned with the order-`order` smoother built by the power-method constructor
   const int order = A.geti("order", 3);
   const double rtol = A.getd("rtol", 1e-8);
   OperatorChebyshevSmoother S(*Aop, diag, ess, order, pit, ptol, seed);
   CGSolver cg;
   cg.SetRelTol(rtol); cg.SetAbsTol(0.0); cg.SetMaxIter(A.geti("maxit", 2000));
   cg.SetOperator(*Aop);
   cg.SetPreconditioner(S);
   std::stringstream hist;
   std::streambuf *old = mfem::out.rdbuf(hist.rdbuf());
   cg.SetPrintLevel(1);
   X = 0.0;
   cg.Mult(B, X);
   mfem::out.rdbuf(old);
   std::vector<double> h;
   {
      string line;
      while (std::getline(hist, line))
      {
         size_t pos = line.find("(B r, r) =");
         if (pos != string::npos) { h.push_back(atof(line.c_str() + pos + 10)); }
      }
   }
   D.dbls("pcg_history", h.data(), h.size());
   D.scalar("pcg_order", order);
   D.scalar("pcg_iters", cg.GetNumIterations());
   D.scalar("pcg_converged", (int)cg.GetConverged());
   D.scalar("pcg_final_norm", (double)cg.GetFinalNorm());
   D.scalar("pcg_rtol", rtol);
   D.dbls("X_sol", X);
   std::printf("dump %s: chebyshev dim=%d NE=%d p=%d lambda=%.17g pcg_iters=%d\n", out, dim, NE, P.p, lam,
               cg.GetNumIterations());
   return 0;
}

// ---- LOR (low-order-refined) batched assembly, H1: fem/lor/lor_batched.cpp, lor_h1_impl.hpp ------
// The macro-element "sparse IJ" array, its CSR form without and with essential dofs eliminated, and
// the legacy path (full assembly on the refined mesh) of the same form as an independent check.
static double lor_mass_coeff(const Vector &x) { return 1.0 + x(0) + 2.0 * x(1) * x(x.Size() - 1); }
static double lor_diff_coeff(const Vector &x) { return 2.0 + std::sin(3.0 * x(0)) * std::cos(2.0 * x(1)); }

static int do_dump_lor(const Args &A, const char *out)
{
   Problem P(A);
   Dump D(out);
   FiniteElementSpace &fes = *P.fes;
   const int dim = P.dim, NE = fes.GetNE(), ndofs = fes.GetNDofs();
   const string kind = A.gets("lorkind", "both"); // mass | diffusion | both
   const bool var = A.geti("varcoef", 0) != 0;
   FunctionCoefficient mc(lor_mass_coeff), dc(lor_diff_coeff);
   ConstantCoefficient mconst(A.getd("mass_const", 1.0)), dconst(A.getd("diff_const", 1.0));
   BilinearForm a(&fes);
   const bool has_m = kind != "diffusion", has_d = kind != "mass";
   if (has_m) { a.AddDomainIntegrator(var ? new MassIntegrator(mc) : new MassIntegrator(mconst)); }
   if (has_d) { a.AddDomainIntegrator(var ? new DiffusionIntegrator(dc) : new DiffusionIntegrator(dconst)); }
   D.scalar("dim", dim); D.scalar("n", P.n); D.scalar("p", P.p); D.scalar("NE", NE); D.scalar("ndofs", ndofs);
   D.scalar("has_mass", (int)has_m); D.scalar("has_diffusion", (int)has_d); D.scalar("varcoef", (int)var);
   auto R = dynamic_cast<const ElementRestriction *>(
               fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
   D.ints("gather_map", R->GatherMap());
   D.ints("offsets", R->Offsets());
   D.ints("indices", R->Indices());
   {
      // the nodes as an E-vector (input of FormLORVertexCoordinates)
      P.mesh.EnsureNodes();
      const GridFunction *nodes = P.mesh.GetNodes();
      const Operator *nr = nodes->FESpace()->GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC);
      Vector ne_(nr->Height());
      nr->Mult(*nodes, ne_);
      D.dbls("nodes_e", ne_);
      D.scalar("node_order", nodes->FESpace()->GetMaxElementOrder());
   }
   // the projected coefficients, as BatchedLORKernel's constructor forms them (lor_batched.hpp:180-216)
   {
      IntegrationRule ir = GetCollocatedIntRule(fes);
      QuadratureSpace qs(*fes.GetMesh(), ir);
      CoefficientVector c1(qs, CoefficientStorage::COMPRESSED), c2(qs, CoefficientStorage::COMPRESSED);
      if (!has_m) { c1.SetConstant(0.0); }
      else if (var) { c1.Project(mc); }
      else { c1.Project(mconst); }
      if (!has_d) { c2.SetConstant(0.0); }
      else if (var) { c2.Project(dc); }
      else { c2.Project(dconst); }
      D.dbls("c_mass", c1); D.dbls("c_diff", c2);
   }
   Array<int> ess;
   {
      Array<int> ess_bdr(P.mesh.bdr_attributes.Max()); ess_bdr = 1;
      fes.GetEssentialTrueDofs(ess_bdr, ess);
   }
   D.ints("ess_tdof_list", ess);
   BatchedLORAssembly batched(fes);
   D.dbls("X_vert", batched.X_vert);
   {
      OperatorHandle A0;
      batched.AssembleWithoutBC(a, A0);
      D.dbls("sparse_ij", batched.sparse_ij);
      D.ints("sparse_mapping", batched.sparse_mapping);
      SparseMatrix *M = A0.As<SparseMatrix>();
      D.ints("I", M->HostReadI(), ndofs + 1);
      D.ints("J", M->HostReadJ(), M->HostReadI()[ndofs]);
      D.dbls("data", M->HostReadData(), M->HostReadI()[ndofs]);
   }
   {
      OperatorHandle A1;
      batched.Assemble(a, ess, A1);
      SparseMatrix *M = A1.As<SparseMatrix>();
      D.ints("I_bc", M->HostReadI(), ndofs + 1);
      D.ints("J_bc", M->HostReadJ(), M->HostReadI()[ndofs]);
      D.dbls("data_bc", M->HostReadData(), M->HostReadI()[ndofs]);
      Vector x(ndofs), y(ndofs);
      x.Randomize(1);
      M->Mult(x, y);
      D.dbls("x", x); D.dbls("y_bc", y);
   }
   double legacy_diff = -1.0;
   {
      // legacy path: refined mesh, lowest-order space, full assembly (fem/lor/lor.cpp, LegacyAssembleSystem)
      LORDiscretization lor(fes);
      lor.LegacyAssembleSystem(a, ess);
      SparseMatrix L = lor.GetAssembledMatrix();
      L.SortColumnIndices();
      D.ints("I_legacy", L.HostReadI(), ndofs + 1);
      D.ints("J_legacy", L.HostReadJ(), L.HostReadI()[ndofs]);
      D.dbls("data_legacy", L.HostReadData(), L.HostReadI()[ndofs]);
      lor.AssembleSystem(a, ess);
      SparseMatrix &B = lor.GetAssembledMatrix();
      std::unique_ptr<SparseMatrix> diff(Add(1.0, L, -1.0, B));
      legacy_diff = diff->MaxNorm();
   }
   D.scalar("legacy_minus_batched_maxnorm", legacy_diff);
   std::printf("dump %s: LOR dim=%d NE=%d p=%d ndofs=%d kind=%s varcoef=%d legacy-batched=%.3e\n", out, dim, NE, P.p,
               ndofs, kind.c_str(), (int)var, legacy_diff);
   return 0;
}

// Timing of the reference CPU path.  Prints one JSON object.
//   mode=mult : y = A x (BilinearForm::Mult, PA) repeated `reps` times
//   mode=kernel : AddMultPA on E-vectors (bench_ceed.cpp BK style), `reps` times
//   mode=all  : kernel, mult and pcg on one problem set-up (mult_seconds / kernel_seconds / seconds)
//   mode=pcg  : bench_ceed.cpp BP style: CG (optionally Jacobi), fixed `iters` iterations
// LOR batched assembly of the reference (host or its own CUDA backend): vertex coordinates once, then
// Assemble (stencil array, CSR, boundary elimination) `reps` times; best and mean seconds.
static int do_bench_lor(const Args &A)
{
   const int reps = A.geti("reps", 3);
   Problem P(A);
   FiniteElementSpace &fes = *P.fes;
   const int ndofs = fes.GetVSize();
   const string kind = A.gets("lorkind", "diffusion");
   ConstantCoefficient one(1.0);
   BilinearForm a(&fes);
   if (kind != "diffusion") { a.AddDomainIntegrator(new MassIntegrator(one)); }
   if (kind != "mass") { a.AddDomainIntegrator(new DiffusionIntegrator(one)); }
   Array<int> ess;
   {
      Array<int> ess_bdr(P.mesh.bdr_attributes.Max()); ess_bdr = 1;
      fes.GetEssentialTrueDofs(ess_bdr, ess);
   }
   fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC); // built outside the timed region
   MFEM_DEVICE_SYNC;
   double t0 = now();
   BatchedLORAssembly batched(fes);
   MFEM_DEVICE_SYNC;
   const double t_vert = now() - t0;
   OperatorHandle Ah;
   double best = 1e300, sum = 0;
   for (int r = 0; r < reps + 1; r++)
   {
      MFEM_DEVICE_SYNC;
      double t1 = now();
      batched.Assemble(a, ess, Ah);
      MFEM_DEVICE_SYNC;
      const double dt = now() - t1;
      if (r == 0) { continue; } // warm-up (allocations)
      best = std::min(best, dt); sum += dt;
   }
   SparseMatrix *M = Ah.As<SparseMatrix>();
   const long long nnz = M->NumNonZeroElems();
   Vector x(ndofs), y(ndofs);
   x.UseDevice(true); y.UseDevice(true);
   x.Randomize(1);
   M->Mult(x, y);
   std::printf("{\"mode\": \"lor\", \"ndofs\": %d, \"ne\": %d, \"p\": %d, \"nnz\": %lld, \"assemble_seconds\": %.6e, "
               "\"assemble_seconds_mean\": %.6e, \"vertex_seconds\": %.6e, \"ynorm\": %.17g, \"device\": \"%s\"}\n",
               ndofs, fes.GetNE(), P.p, nnz, best, sum / reps, t_vert, (double)y.Norml2(),
               Device::IsEnabled() ? "cuda" : "cpu");
   return 0;
}

static int do_bench(const Args &A)
{
   const string mode = A.gets("mode", "pcg");
   if (mode == "lor") { return do_bench_lor(A); }
   const int reps = A.geti("reps", 3), iters = A.geti("iters", 32);
   const int skip = std::min(A.geti("skip", 0), reps - 1); // pcg: leading solves left out of the average
   const int use_prec = A.geti("jacobi", 0);
   double t0 = now();
   Problem P(A);
   P.a->Assemble();
   FiniteElementSpace &fes = *P.fes;
   const int ndofs = fes.GetVSize();
   double t_setup = now() - t0;
   Vector x(ndofs), y(ndofs);
   x.UseDevice(true); y.UseDevice(true);
   x.Randomize(1);
   double sec = 0, ynorm = 0, final_norm = 0, initial_norm = 0; int its = 0;
   double sec_mult = 0, sec_kernel = 0; // mode=all: the three timings of one problem set-up
   const bool all = mode == "all";
   if (mode == "mult" || all)
   {
      P.a->Mult(x, y); // warm-up
      MFEM_DEVICE_SYNC;
      double t1 = now();
      for (int i = 0; i < reps; i++) { P.a->Mult(x, y); }
      MFEM_DEVICE_SYNC;
      sec = (now() - t1) / reps; its = 1;
      sec_mult = sec;
      ynorm = y.Norml2();
   }
   if (mode == "kernel" || all)
   {
      // BK style (bench_ceed.cpp:258-298): AddMultPA on E-vectors only, i.e. the reference's
      // SmemPADiffusionApply3D / SmemPAMassApply3D (or its fallback) without the restriction.
      const Operator *R = fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC);
      Vector xe(R->Height()), ye(R->Height());
      xe.UseDevice(true); ye.UseDevice(true);
      R->Mult(x, xe);
      ye = 0.0;
      P.bfi->AddMultPA(xe, ye); // warm-up
      ye = 0.0;
      MFEM_DEVICE_SYNC;
      double t1 = now();
      for (int i = 0; i < reps; i++) { P.bfi->AddMultPA(xe, ye); }
      MFEM_DEVICE_SYNC;
      sec = (now() - t1) / reps; its = 1;
      sec_kernel = sec;
      ye *= 1.0 / reps;
      if (!all) { ynorm = ye.Norml2(); }
   }
   if (mode == "pcg" || all)
   {
      sec = 0;
      Array<int> ess_bdr(P.mesh.bdr_attributes.Max()); ess_bdr = 1;
      Array<int> ess;
      fes.GetEssentialTrueDofs(ess_bdr, ess);
      Vector b(fes.GetVSize());
      if (P.vdim == 1)
      {
         LinearForm lf(&fes);
         lf.AddDomainIntegrator(new DomainLFIntegrator(P.one));
         lf.Assemble();
         b = lf;
      }
      else { b.Randomize(2); }
      GridFunction u(&fes); u = 0.0;
      OperatorPtr Aop; Vector B, X;
      P.a->FormLinearSystem(ess, u, b, Aop, X, B);
      OperatorJacobiSmoother M(*P.a, ess);
      CGSolver cg;
      cg.SetRelTol(0.0); cg.SetAbsTol(0.0); cg.SetMaxIter(iters); cg.SetPrintLevel(-1);
      cg.SetOperator(*Aop);
      if (use_prec) { cg.SetPreconditioner(M); }
      double best = 1e300;
      for (int r = 0; r < reps; r++)
      {
         X = 0.0;
         MFEM_DEVICE_SYNC;
         double t1 = now();
         cg.Mult(B, X);
         MFEM_DEVICE_SYNC;
         double t = now() - t1;
         best = std::min(best, t);
         if (r >= skip) { sec += t; }
      }
      sec /= (reps - skip); its = cg.GetNumIterations();
      ynorm = X.Norml2();
      final_norm = cg.GetFinalNorm(); initial_norm = cg.GetInitialNorm();
   }
   const char *dev = getenv("REF_DRIVER_DEVICE");
   if (!dev) { dev = getenv("MFEM_DEVICE"); }
   std::printf("{\"mode\": \"%s\", \"ndofs\": %d, \"ne\": %d, \"p\": %d, \"iters\": %d, "
               "\"seconds\": %.6e, \"setup_seconds\": %.3f, \"dofs_iters_per_s\": %.6e, "
               "\"norm\": %.17g, \"final_norm\": %.17g, \"initial_norm\": %.17g, "
               "\"mult_seconds\": %.6e, \"kernel_seconds\": %.6e, \"device\": \"%s\"}\n",
               mode.c_str(), ndofs, fes.GetNE(), P.p, its, sec, t_setup,
               (double)ndofs * its / sec, ynorm, final_norm, initial_norm, sec_mult, sec_kernel,
               dev ? dev : "cpu");
   return 0;
}

int main(int argc, char **argv)
{
   if (argc < 2)
   {
      std::fprintf(stderr, "usage: ref_driver dump <out.bin> [--key val ...] | bench [--key val ...]\n");
      return 1;
   }
   // MFEM_DEVICE (read by the reference itself, device.cpp:145): "omp" uses all host threads.
   // REF_DRIVER_DEVICE: configured here, in main, as examples/ex1.cpp does with -d (the CUDA
   // backend must not be brought up from the singleton's static initialiser).
   const char *dev = getenv("REF_DRIVER_DEVICE");
   if (!dev) { dev = getenv("MFEM_DEVICE"); }
   Device device(dev ? dev : "cpu");
   string cmd = argv[1];
   if (cmd == "dump")
   {
      Args A(argc, argv, 3);
      if (A.geti("nonsym", 0)) { return do_dump_nonsym(A, argv[2]); }
      if (A.geti("cheb", 0)) { return do_dump_cheb(A, argv[2]); }
      if (A.geti("lor", 0)) { return do_dump_lor(A, argv[2]); }
      if (A.geti("markers", 0)) { return do_dump_markers(A, argv[2]); }
      if (A.geti("lf", 0)) { return do_dump_lf(A, argv[2]); }
      if (A.geti("lean", 0)) { return do_dump_mesh(A, argv[2]); }
      return A.geti("vdim", 1) > 1 ? do_dump_vector(A, argv[2]) : do_dump(A, argv[2]);
   }
   if (cmd == "bench") { Args A(argc, argv, 2); return do_bench(A); }
   std::fprintf(stderr, "unknown command %s\n", cmd.c_str());
   return 1;
}
