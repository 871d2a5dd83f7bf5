// adapter/b200_mfem.hpp -- MFEM-side binding of libb200fem.so (include/b200fem.h).
//
// Compiled against MFEM's own headers (v4.9.1-dev); this is the code a maintainer adds to an
// MFEM application to route the partial-assembly hot path to the B200 library.  It honours
// MFEM's three nested seams:
//
//   seam 3  b200::DiffusionIntegrator / b200::MassIntegrator : mfem::{Diffusion,Mass}Integrator
//           override AssemblePA / AddMultPA / AssembleDiagonalPA (fem/bilininteg.hpp:49-95), so
//           an unchanged `BilinearForm a; a.SetAssemblyLevel(PARTIAL); a.AddDomainIntegrator(
//           new b200::DiffusionIntegrator(one))` runs MFEM's own PABilinearFormExtension with the
//           B200 element kernels (E-vector in, E-vector out).
//   seam 1  b200::PAOperator : mfem::Operator (linalg/operator.hpp:25-313) -- the FUSED operator
//           y = R^T A_E R x on L-vectors (one kernel, no E-vector in memory), with
//           AssembleDiagonal; b200::ConstrainedOperator wraps it the way
//           Operator::FormLinearSystem does (linalg/operator.cpp:130-146, 526-660).
//           b200::CGSolver : mfem::IterativeSolver runs the whole PCG (linalg/solvers.cpp:
//           869-1050) on the device with device-resident scalars.
//
// Memory: "host-MFEM" mode: MFEM is built CPU-only, Vectors live
// on the host, the adapter owns the device copies and moves x / y across the boundary of each
// call (the CG adapter moves b in and x out once per solve).  With an MFEM built with
// MFEM_USE_CUDA the same calls take x.Read() / y.Write() device pointers and the copies vanish.
// Errors: a failing library call becomes mfem::mfem_error (MFEM_ABORT), the reference's policy.
#ifndef B200_MFEM_HPP
#define B200_MFEM_HPP

#include <memory>
#include <vector>

#include "mfem.hpp"
#include "../include/b200fem.h"

namespace b200
{

#define B200_CHECK(call)                                                               \
   do                                                                                  \
   {                                                                                   \
      if ((call) != B2F_OK) { MFEM_ABORT("libb200fem: " << b2f_last_error()); }        \
   } while (0)

/// The process-wide library context: the role of mfem::Device("cuda") (general/device.hpp:128).
class Context
{
   b2f_ctx *ctx = nullptr;
   Context(int dev) { B200_CHECK(b2f_ctx_create(dev, nullptr, &ctx)); }
public:
   ~Context() { b2f_ctx_destroy(ctx); }
   static Context &Get(int dev = 0)
   {
      static Context c(dev);
      return c;
   }
   b2f_ctx *get() const { return ctx; }
};

/// A device buffer of doubles (what mfem::Vector's device pointer would be).
class DeviceVector
{
   double *p = nullptr;
   size_t n = 0;
public:
   DeviceVector() = default;
   explicit DeviceVector(size_t n_) { Resize(n_); }
   DeviceVector(const DeviceVector &) = delete;
   ~DeviceVector() { if (p) { b2f_free(Context::Get().get(), p); } }
   void Resize(size_t n_)
   {
      if (n_ == n) { return; }
      if (p) { b2f_free(Context::Get().get(), p); p = nullptr; }
      n = n_;
      B200_CHECK(b2f_malloc(Context::Get().get(), sizeof(double) * (n ? n : 1), (void **)&p));
   }
   double *Ptr() const { return p; }
   size_t Size() const { return n; }
   void Upload(const mfem::Vector &v)
   {
      Resize(v.Size());
      B200_CHECK(b2f_memcpy_h2d(Context::Get().get(), p, v.HostRead(), sizeof(double) * n));
   }
   void Download(mfem::Vector &v) const
   {
      MFEM_VERIFY((size_t)v.Size() == n, "size mismatch");
      B200_CHECK(b2f_memcpy_d2h(Context::Get().get(), v.HostWrite(), p, sizeof(double) * n));
      B200_CHECK(b2f_ctx_sync(Context::Get().get()));
   }
};

/// Shared implementation of the two integrators: builds b2f_pa from MFEM's own tables.
class PAData
{
public:
   b2f_pa *pa = nullptr;
   int ne = 0, dof = 0, vdim = 1;
   mutable DeviceVector xe, ye;
   ~PAData() { b2f_pa_destroy(pa); }

   /** AssemblePA: B, G from MFEM's DofToQuad (fem/fe/fe_base.cpp:2697), weights from the
       IntegrationRule, geometry from the mesh nodes E-vector (mesh/mesh.cpp:15445); the
       Jacobians are evaluated inside the setup kernel and never stored. */
   void Assemble(int kind, const mfem::FiniteElementSpace &fes, const mfem::IntegrationRule &ir,
                 double coeff)
   {
      using namespace mfem;
      Mesh &mesh = *fes.GetMesh();
      const FiniteElement &fe = *fes.GetTypicalFE();
      const DofToQuad &maps = fe.GetDofToQuad(ir, DofToQuad::TENSOR);
      mesh.EnsureNodes();
      const GridFunction *nodes = mesh.GetNodes();
      const FiniteElementSpace *nfes = nodes->FESpace();
      const Operator *nR = nfes->GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC);
      Vector nodesE(nR->Height());
      nR->Mult(*nodes, nodesE);
      const FiniteElement &nfe = *nfes->GetTypicalFE();
      const DofToQuad &nmaps = nfe.GetDofToQuad(ir, DofToQuad::TENSOR);
      DeviceVector d_nodes;
      d_nodes.Upload(nodesE);
      ne = fes.GetNE();
      dof = fe.GetDof();
      vdim = fes.GetVDim();
      b2f_pa_destroy(pa);
      pa = nullptr;
      B200_CHECK(b2f_pa_setup_from_nodes(Context::Get().get(), kind, mesh.Dimension(), maps.ndof,
                                         maps.nqpt, ne, maps.B.HostRead(), maps.G.HostRead(),
                                         ir.GetWeights().HostRead(), nmaps.ndof, nmaps.B.HostRead(),
                                         nmaps.G.HostRead(), d_nodes.Ptr(), coeff, &pa));
   }
   /** AssemblePA with a general coefficient, exactly the reference seam
       (bilininteg_diffusion_pa.cpp:108-171, bilininteg_mass_pa.cpp:25-91): the
       CoefficientVector MFEM projects at the quadrature points (scalar, vector, symmetric or
       full matrix) and GeometricFactors J / detJ go to b2f_diffusion_setup / b2f_mass_setup. */
   void AssembleGeneral(int kind, const mfem::FiniteElementSpace &fes, const mfem::IntegrationRule &ir,
                        const mfem::CoefficientVector &coeff)
   {
      using namespace mfem;
      Mesh &mesh = *fes.GetMesh();
      const FiniteElement &fe = *fes.GetTypicalFE();
      const DofToQuad &maps = fe.GetDofToQuad(ir, DofToQuad::TENSOR);
      ne = fes.GetNE();
      dof = fe.GetDof();
      vdim = fes.GetVDim();
      b2f_pa_destroy(pa);
      pa = nullptr;
      DeviceVector d_geo, d_c;
      d_c.Upload(coeff);
      b2f_ctx *ctx = Context::Get().get();
      if (kind == 0)
      {
         const GeometricFactors *geom = mesh.GetGeometricFactors(ir, GeometricFactors::JACOBIANS);
         d_geo.Upload(geom->J);
         B200_CHECK(b2f_diffusion_setup(ctx, mesh.Dimension(), maps.ndof, maps.nqpt, ne, maps.B.HostRead(),
                                        maps.G.HostRead(), ir.GetWeights().HostRead(), d_geo.Ptr(), d_c.Ptr(),
                                        coeff.GetVDim(), coeff.Size(), &pa));
      }
      else
      {
         const GeometricFactors *geom = mesh.GetGeometricFactors(ir, GeometricFactors::DETERMINANTS);
         d_geo.Upload(geom->detJ);
         B200_CHECK(b2f_mass_setup(ctx, mesh.Dimension(), maps.ndof, maps.nqpt, ne, maps.B.HostRead(),
                                   ir.GetWeights().HostRead(), d_geo.Ptr(), d_c.Ptr(), coeff.Size(), &pa));
      }
      B200_CHECK(b2f_ctx_sync(ctx));
   }
   void AddMult(const mfem::Vector &x, mfem::Vector &y) const
   {
      xe.Upload(x);
      ye.Upload(y); // AddMultPA accumulates into y (fem/bilininteg.hpp:78)
      B200_CHECK(b2f_pa_add_mult_vec(pa, vdim, xe.Ptr(), ye.Ptr())); // E-vector (dofs, vdim, NE)
      ye.Download(y);
   }
   void AddDiagonal(mfem::Vector &diag) const
   {
      if (vdim == 1)
      {
         ye.Upload(diag);
         B200_CHECK(b2f_pa_assemble_diagonal(pa, ye.Ptr()));
         ye.Download(diag);
         return;
      }
      // vector integrators: the scalar element diagonal added to every component
      mfem::Vector d(ne * dof);
      d = 0.0;
      ye.Upload(d);
      B200_CHECK(b2f_pa_assemble_diagonal(pa, ye.Ptr()));
      ye.Download(d);
      for (int e = 0; e < ne; e++)
         for (int c = 0; c < vdim; c++)
            for (int i = 0; i < dof; i++) { diag(i + dof * (c + vdim * e)) += d(i + dof * e); }
   }
};

inline double ConstantOf(mfem::Coefficient *Q)
{
   if (!Q) { return 1.0; }
   auto c = dynamic_cast<mfem::ConstantCoefficient *>(Q);
   MFEM_VERIFY(c, "b200 vector integrators take a ConstantCoefficient (or none)");
   return c->constant;
}
inline bool IsConstant(mfem::Coefficient *Q) { return !Q || dynamic_cast<mfem::ConstantCoefficient *>(Q); }

/// seam 3: drop-in for mfem::DiffusionIntegrator on the PA path.
class DiffusionIntegrator : public mfem::DiffusionIntegrator
{
   PAData data;
public:
   using mfem::DiffusionIntegrator::DiffusionIntegrator;
   using mfem::BilinearFormIntegrator::AssemblePA;
   void AssemblePA(const mfem::FiniteElementSpace &fes) override
   {
      const mfem::FiniteElement &fe = *fes.GetTypicalFE();
      const mfem::IntegrationRule *ir = GetIntegrationRule();
      if (!ir) { ir = &GetRule(fe, fe); }
      if (!VQ && !MQ && IsConstant(Q)) { data.Assemble(0, fes, *ir, ConstantOf(Q)); return; }
      // spatially varying scalar, vector (diagonal) or matrix coefficient: the reference's own
      // projection at the quadrature points (bilininteg_diffusion_pa.cpp:155-163)
      mfem::QuadratureSpace qs(*fes.GetMesh(), *ir);
      mfem::CoefficientVector coeff(qs, mfem::CoefficientStorage::COMPRESSED);
      if (MQ) { coeff.ProjectTranspose(*MQ); }
      else if (VQ) { coeff.Project(*VQ); }
      else { coeff.Project(*Q); }
      data.AssembleGeneral(0, fes, *ir, coeff);
   }
   void AddMultPA(const mfem::Vector &x, mfem::Vector &y) const override { data.AddMult(x, y); }
   void AssembleDiagonalPA(mfem::Vector &diag) override { data.AddDiagonal(diag); }
   b2f_pa *Handle() const { return data.pa; }
};

/// seam 3: drop-in for mfem::MassIntegrator on the PA path.
class MassIntegrator : public mfem::MassIntegrator
{
   PAData data;
public:
   using mfem::MassIntegrator::MassIntegrator;
   using mfem::BilinearFormIntegrator::AssemblePA;
   void AssemblePA(const mfem::FiniteElementSpace &fes) override
   {
      const mfem::FiniteElement &fe = *fes.GetTypicalFE();
      const mfem::IntegrationRule *ir = GetIntegrationRule();
      if (!ir) { ir = &GetRule(fe, fe, *fes.GetMesh()->GetTypicalElementTransformation()); }
      if (IsConstant(Q)) { data.Assemble(1, fes, *ir, ConstantOf(Q)); return; }
      mfem::QuadratureSpace qs(*fes.GetMesh(), *ir);
      mfem::CoefficientVector coeff(Q, qs, mfem::CoefficientStorage::COMPRESSED); // bilininteg_mass_pa.cpp:46
      data.AssembleGeneral(1, fes, *ir, coeff);
   }
   void AddMultPA(const mfem::Vector &x, mfem::Vector &y) const override { data.AddMult(x, y); }
   void AssembleDiagonalPA(mfem::Vector &diag) override { data.AddDiagonal(diag); }
   b2f_pa *Handle() const { return data.pa; }
};

/// seam 3 for vector spaces (CEED BP4/BP6): drop-in for mfem::VectorDiffusionIntegrator.
class VectorDiffusionIntegrator : public mfem::VectorDiffusionIntegrator
{
   PAData data;
public:
   using mfem::VectorDiffusionIntegrator::VectorDiffusionIntegrator;
   using mfem::BilinearFormIntegrator::AssemblePA;
   void AssemblePA(const mfem::FiniteElementSpace &fes) override
   {
      const mfem::FiniteElement &fe = *fes.GetTypicalFE();
      const mfem::IntegrationRule *ir = GetIntegrationRule();
      if (!ir) { ir = &mfem::DiffusionIntegrator::GetRule(fe, fe); }
      data.Assemble(0, fes, *ir, ConstantOf(Q));
   }
   void AddMultPA(const mfem::Vector &x, mfem::Vector &y) const override { data.AddMult(x, y); }
   void AssembleDiagonalPA(mfem::Vector &diag) override { data.AddDiagonal(diag); }
   b2f_pa *Handle() const { return data.pa; }
};

/// seam 3 for vector spaces (CEED BP2): drop-in for mfem::VectorMassIntegrator.
class VectorMassIntegrator : public mfem::VectorMassIntegrator
{
   PAData data;
public:
   using mfem::VectorMassIntegrator::VectorMassIntegrator;
   using mfem::BilinearFormIntegrator::AssemblePA;
   void AssemblePA(const mfem::FiniteElementSpace &fes) override
   {
      const mfem::FiniteElement &fe = *fes.GetTypicalFE();
      const mfem::IntegrationRule *ir = GetIntegrationRule();
      if (!ir) { ir = &mfem::MassIntegrator::GetRule(fe, fe, *fes.GetMesh()->GetTypicalElementTransformation()); }
      data.Assemble(1, fes, *ir, ConstantOf(Q));
   }
   void AddMultPA(const mfem::Vector &x, mfem::Vector &y) const override { data.AddMult(x, y); }
   void AssembleDiagonalPA(mfem::Vector &diag) override { data.AddDiagonal(diag); }
   b2f_pa *Handle() const { return data.pa; }
};

/// seam 1: the fused PA operator of a BilinearForm whose domain integrators are b200 ones.
class PAOperator : public mfem::Operator
{
protected:
   b2f_restriction *restr = nullptr;
   b2f_operator *op = nullptr;
   std::vector<b2f_pa *> integ;
   mutable DeviceVector dx, dy;
   friend class CGSolver;
public:
   /** @param a  a BilinearForm with AssemblyLevel::PARTIAL, already Assemble()d, whose domain
                  integrators are b200::DiffusionIntegrator / b200::MassIntegrator.
       @param ess_tdof_list  essential dofs (FormLinearSystem semantics), may be empty. */
   PAOperator(mfem::BilinearForm &a, const mfem::Array<int> &ess_tdof_list)
      : mfem::Operator(a.FESpace()->GetVSize())
   {
      using namespace mfem;
      const FiniteElementSpace &fes = *a.FESpace();
      auto R = dynamic_cast<const ElementRestriction *>(
                  fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
      MFEM_VERIFY(R, "tensor-product H1 space expected");
      const int dof = fes.GetTypicalFE()->GetDof();
      // MFEM's own tables (fem/restriction.hpp:104-106) are handed over unchanged
      B200_CHECK(b2f_restriction_create(Context::Get().get(), fes.GetNE(), dof, fes.GetNDofs(), fes.GetVDim(),
                                        fes.GetOrdering() == Ordering::byVDIM, R->GatherMap().HostRead(),
                                        R->Offsets().HostRead(), R->Indices().HostRead(), &restr));
      for (BilinearFormIntegrator *bfi : *a.GetDBFI())
      {
         if (auto d = dynamic_cast<b200::DiffusionIntegrator *>(bfi)) { integ.push_back(d->Handle()); }
         else if (auto m = dynamic_cast<b200::MassIntegrator *>(bfi)) { integ.push_back(m->Handle()); }
         else if (auto vd = dynamic_cast<b200::VectorDiffusionIntegrator *>(bfi)) { integ.push_back(vd->Handle()); }
         else if (auto vm = dynamic_cast<b200::VectorMassIntegrator *>(bfi)) { integ.push_back(vm->Handle()); }
         else { MFEM_ABORT("PAOperator: only the b200 integrators are supported"); }
      }
      B200_CHECK(b2f_operator_create(Context::Get().get(), restr, integ.data(), (int)integ.size(),
                                     ess_tdof_list.HostRead(), ess_tdof_list.Size(), &op));
      // AddDomainIntegrator(bfi, elem_marker): hand over MFEM's own marker and attribute arrays
      Array<Array<int>*> &markers = *a.GetDBFI_Marker();
      for (int i = 0; i < markers.Size(); i++)
      {
         if (!markers[i]) { continue; }
         const Array<int> &attr = fes.GetMesh()->GetElementAttributes();
         B200_CHECK(b2f_operator_set_markers(op, i, attr.HostRead(), markers[i]->HostRead(), markers[i]->Size()));
      }
   }
   ~PAOperator() override
   {
      b2f_operator_destroy(op);
      b2f_restriction_destroy(restr);
   }
   /// PABilinearFormExtension::Mult (fem/bilinearform_ext.cpp:487): y = R^T A_E R x, fused.
   void Mult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      dx.Upload(x);
      dy.Resize(height);
      B200_CHECK(b2f_operator_mult(op, dx.Ptr(), dy.Ptr()));
      dy.Download(y);
   }
   void MultTranspose(const mfem::Vector &x, mfem::Vector &y) const override { Mult(x, y); } // symmetric
   /// PABilinearFormExtension::AbsMult (fem/bilinearform_ext.hpp:166)
   void AbsMult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      dx.Upload(x);
      dy.Resize(height);
      B200_CHECK(b2f_operator_abs_mult(op, dx.Ptr(), dy.Ptr()));
      dy.Download(y);
   }
   void AssembleDiagonal(mfem::Vector &diag) const override
   {
      dy.Resize(height);
      B200_CHECK(b2f_operator_assemble_diagonal(op, dy.Ptr()));
      diag.SetSize(height);
      dy.Download(diag);
   }
   b2f_operator *Handle() const { return op; }
};

/// ConstrainedOperator (linalg/operator.cpp:526-660, DIAG_ONE) over the fused operator.
class ConstrainedOperator : public PAOperator
{
public:
   using PAOperator::PAOperator;
   void Mult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      dx.Upload(x);
      dy.Resize(height);
      B200_CHECK(b2f_operator_mult_constrained(op, dx.Ptr(), dy.Ptr()));
      dy.Download(y);
   }
   /// ConstrainedOperator::AbsMult (linalg/operator.cpp:663-735) over the extension's
   /// MultInternal(useAbs) (fem/bilinearform_ext.cpp:487-564).
   void AbsMult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      dx.Upload(x);
      dy.Resize(height);
      B200_CHECK(b2f_operator_abs_mult_constrained(op, dx.Ptr(), dy.Ptr()));
      dy.Download(y);
   }
   /// Operator::DiagonalPolicy of the constrained rows (default DIAG_ONE as FormLinearSystem).
   void SetDiagonalPolicy(mfem::Operator::DiagonalPolicy p)
   {
      B200_CHECK(b2f_operator_set_diag_policy(op, p == mfem::Operator::DIAG_ZERO ? 0 : p == mfem::Operator::DIAG_ONE ? 1 : 2));
   }
   /// EliminateRHS (linalg/operator.cpp:574): b -= A w (w = x on ess); b[ess] = x[ess].
   void EliminateRHS(const mfem::Vector &x, mfem::Vector &b) const
   {
      dx.Upload(x);
      dy.Upload(b);
      B200_CHECK(b2f_operator_eliminate_rhs(op, dx.Ptr(), dy.Ptr()));
      dy.Download(b);
   }
};

/// DomainLFIntegrator whose device assembly (LinearForm::UseFastAssembly(true),
/// fem/linearform_ext.cpp:20-66 -> fem/integ/lininteg_domain.cpp:21-59) runs on the B200:
/// the E-vector b_e += B^T (W f detJ) comes from b2f_domain_lf_assemble; MFEM's
/// LinearFormExtension keeps doing the R^T scatter.
class DomainLFIntegrator : public mfem::DomainLFIntegrator
{
   mfem::Coefficient &Qf;
   int qa, qb;
   const mfem::IntegrationRule *ir_given;
public:
   DomainLFIntegrator(mfem::Coefficient &QF, int a = 2, int b = 0)
      : mfem::DomainLFIntegrator(QF, a, b), Qf(QF), qa(a), qb(b), ir_given(nullptr) { }
   DomainLFIntegrator(mfem::Coefficient &QF, const mfem::IntegrationRule *ir)
      : mfem::DomainLFIntegrator(QF, ir), Qf(QF), qa(2), qb(0), ir_given(ir) { }
   void AssembleDevice(const mfem::FiniteElementSpace &fes, const mfem::Array<int> &markers,
                       mfem::Vector &b) override
   {
      using namespace mfem;
      const FiniteElement &fe = *fes.GetTypicalFE();
      const IntegrationRule *ir = ir_given ? ir_given : &IntRules.Get(fe.GetGeomType(), qa * fe.GetOrder() + qb);
      const DofToQuad &maps = fe.GetDofToQuad(*ir, DofToQuad::TENSOR);
      Mesh *mesh = fes.GetMesh();
      const GeometricFactors *geom = mesh->GetGeometricFactors(*ir, GeometricFactors::DETERMINANTS);
      QuadratureSpace qs(*mesh, *ir);
      CoefficientVector coeff(Qf, qs, CoefficientStorage::COMPRESSED);
      b2f_ctx *ctx = Context::Get().get();
      DeviceVector d_detj, d_coeff, d_b;
      d_detj.Upload(geom->detJ);
      d_coeff.Upload(coeff);
      d_b.Upload(b);
      int *d_mk = nullptr;
      B200_CHECK(b2f_malloc(ctx, sizeof(int) * (markers.Size() + 1), (void **)&d_mk));
      B200_CHECK(b2f_memcpy_h2d(ctx, d_mk, markers.HostRead(), sizeof(int) * markers.Size()));
      B200_CHECK(b2f_domain_lf_assemble(ctx, mesh->Dimension(), maps.ndof, maps.nqpt, fes.GetVDim(), fes.GetNE(),
                                        d_mk, maps.B.HostRead(), ir->GetWeights().HostRead(), d_detj.Ptr(),
                                        d_coeff.Ptr(), coeff.Size(), d_b.Ptr()));
      d_b.Download(b);
      b2f_free(ctx, d_mk);
   }
};

/// OperatorChebyshevSmoother (linalg/solvers.hpp:392-505, solvers.cpp:453-658) over the
/// constrained fused operator; both reference constructors (given eigenvalue estimate /
/// power method with MFEM's own Vector::Randomize(power_seed) start vector).
class OperatorChebyshevSmoother : public mfem::Solver
{
   const ConstrainedOperator &A;
   b2f_chebyshev *h = nullptr;
   double max_eig = 0.0;
   mutable DeviceVector dx, dy;
public:
   OperatorChebyshevSmoother(const ConstrainedOperator &A_, const mfem::Vector &d, int order,
                             double max_eig_estimate)
      : mfem::Solver(d.Size()), A(A_), max_eig(max_eig_estimate)
   {
      DeviceVector dd;
      dd.Upload(d);
      B200_CHECK(b2f_chebyshev_create(A.Handle(), dd.Ptr(), order, max_eig, &h));
      B200_CHECK(b2f_ctx_sync(Context::Get().get()));
   }
   OperatorChebyshevSmoother(const ConstrainedOperator &A_, const mfem::Vector &d,
                             const mfem::Array<int> &ess_tdof_list, int order, int power_iterations = 10,
                             double power_tolerance = 1e-8, int power_seed = 12345)
      : mfem::Solver(d.Size()), A(A_)
   {
      MFEM_VERIFY(power_seed != 0, "invalid seed!");
      b2f_ctx *ctx = Context::Get().get();
      DeviceVector dd, dinv(d.Size()), ev;
      dd.Upload(d);
      int *d_ess = nullptr;
      B200_CHECK(b2f_malloc(ctx, sizeof(int) * (ess_tdof_list.Size() + 1), (void **)&d_ess));
      B200_CHECK(b2f_memcpy_h2d(ctx, d_ess, ess_tdof_list.HostRead(), sizeof(int) * ess_tdof_list.Size()));
      B200_CHECK(b2f_jacobi_setup(ctx, d.Size(), dd.Ptr(), d_ess, ess_tdof_list.Size(), 1.0, dinv.Ptr()));
      mfem::Vector v0(d.Size());
      v0.Randomize(power_seed);
      ev.Upload(v0);
      B200_CHECK(b2f_power_method(A.Handle(), dinv.Ptr(), ev.Ptr(), power_iterations, power_tolerance,
                                  &max_eig));
      b2f_free(ctx, d_ess);
      B200_CHECK(b2f_chebyshev_create(A.Handle(), dd.Ptr(), order, max_eig, &h));
      B200_CHECK(b2f_ctx_sync(ctx));
   }
   ~OperatorChebyshevSmoother() override { b2f_chebyshev_destroy(h); }
   void SetOperator(const mfem::Operator &) override { }
   double MaxEigEstimate() const { return max_eig; }
   b2f_chebyshev *Handle() const { return h; }
   void Mult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      MFEM_VERIFY(!iterative_mode, "Chebyshev smoother not implemented for iterative mode");
      dx.Upload(x);
      dy.Resize(height);
      B200_CHECK(b2f_chebyshev_mult(h, dx.Ptr(), dy.Ptr()));
      dy.Download(y);
   }
};

/// CGSolver (linalg/solvers.cpp:869-1050) with an optional Jacobi preconditioner
/// (OperatorJacobiSmoother, solvers.cpp:324-450), entirely on the device.
class CGSolver : public mfem::IterativeSolver
{
   const ConstrainedOperator *A = nullptr;
   bool jacobi = false;
   const OperatorChebyshevSmoother *cheb = nullptr;
   DeviceVector dinv;
   mutable DeviceVector db, dxs;
public:
   void SetOperator(const mfem::Operator &op_) override
   {
      A = dynamic_cast<const ConstrainedOperator *>(&op_);
      MFEM_VERIFY(A, "b200::CGSolver needs a b200::ConstrainedOperator");
      height = width = A->Height();
      oper = A;
   }
   /// OperatorJacobiSmoother(a, ess): dinv = 1/diag, dinv[ess] = 1.
   void UseJacobi(const mfem::Array<int> &ess_tdof_list)
   {
      MFEM_VERIFY(A, "SetOperator first");
      DeviceVector diag(height);
      B200_CHECK(b2f_operator_assemble_diagonal(A->Handle(), diag.Ptr()));
      dinv.Resize(height);
      int *d_ess = nullptr;
      b2f_ctx *ctx = Context::Get().get();
      B200_CHECK(b2f_malloc(ctx, sizeof(int) * (ess_tdof_list.Size() + 1), (void **)&d_ess));
      B200_CHECK(b2f_memcpy_h2d(ctx, d_ess, ess_tdof_list.HostRead(), sizeof(int) * ess_tdof_list.Size()));
      B200_CHECK(b2f_jacobi_setup(ctx, height, diag.Ptr(), d_ess, ess_tdof_list.Size(), 1.0, dinv.Ptr()));
      B200_CHECK(b2f_ctx_sync(ctx));
      b2f_free(ctx, d_ess);
      jacobi = true;
      cheb = nullptr;
   }
   /// Only the b200 Chebyshev smoother is accepted here (Jacobi: UseJacobi).
   void SetPreconditioner(mfem::Solver &pr) override
   {
      cheb = dynamic_cast<const OperatorChebyshevSmoother *>(&pr);
      MFEM_VERIFY(cheb, "b200::CGSolver: preconditioner must be a b200::OperatorChebyshevSmoother");
      jacobi = false;
   }
   void Mult(const mfem::Vector &b, mfem::Vector &x) const override
   {
      MFEM_VERIFY(A, "operator not set");
      b2f_ctx *ctx = Context::Get().get();
      db.Upload(b);
      dxs.Resize(height);
      DeviceVector x0, ax0;
      if (iterative_mode)
      {
         // r = b - A x0; the library iterates on the correction from a zero start, which
         // generates the same Krylov iterates as CGSolver::Mult does from x0
         x0.Upload(x);
         ax0.Resize(height);
         B200_CHECK(b2f_operator_mult_constrained(A->Handle(), x0.Ptr(), ax0.Ptr()));
         B200_CHECK(b2f_vector_add(ctx, height, db.Ptr(), -1.0, ax0.Ptr(), db.Ptr()));
      }
      b2f_pcg_options opt = {rel_tol, abs_tol, max_iter, 1};
      b2f_pcg_result res;
      if (cheb)
      {
         B200_CHECK(b2f_pcg_chebyshev(A->Handle(), cheb->Handle(), db.Ptr(), dxs.Ptr(), &opt, &res, nullptr, 0));
      }
      else
      {
         B200_CHECK(b2f_pcg(A->Handle(), jacobi ? dinv.Ptr() : nullptr, db.Ptr(), dxs.Ptr(), &opt, &res,
                            nullptr, 0));
      }
      if (iterative_mode) { B200_CHECK(b2f_vector_add(ctx, height, x0.Ptr(), 1.0, dxs.Ptr(), dxs.Ptr())); }
      dxs.Download(x);
      final_iter = res.iterations;
      converged = res.converged != 0;
      initial_norm = res.initial_norm;
      final_norm = res.final_norm;
   }
};

} // namespace b200
#endif
