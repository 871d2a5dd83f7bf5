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
//   seam 2  b200::PABilinearFormExtension : mfem::PABilinearFormExtension (fem/bilinearform_ext.hpp:
//           67-160), installed by b200::BilinearForm::SetAssemblyLevel(PARTIAL) the way
//           BilinearForm::SetAssemblyLevel installs the stock one (fem/bilinearform.cpp:109-136): an
//           otherwise UNCHANGED driver (a.Assemble(); a.FormLinearSystem(...); a.Mult(...);
//           OperatorJacobiSmoother M(a, ess)) runs the fused operator, the fused diagonal and the
//           fused constrained operator.  Stock mfem::DiffusionIntegrator / mfem::MassIntegrator
//           objects are accepted as they are (their coefficient and rule are read through the
//           public getters / a protected-member pointer).
//   seam 1  b200::PAOperator : mfem::Operator (linalg/operator.hpp:25-313) -- the FUSED operator
//           y = R^T A_E R x on L-vectors (one kernel, no E-vector in memory), with
//           AssembleDiagonal; b200::ConstrainedOperator wraps it the way
//           Operator::FormLinearSystem does (linalg/operator.cpp:130-146, 526-660).
//           b200::CGSolver : mfem::IterativeSolver runs the whole PCG (linalg/solvers.cpp:
//           869-1050) on the device with device-resident scalars.
//
// Memory, two modes chosen at run time by mfem::Device:
//   (i)  device-resident: MFEM built with MFEM_USE_CUDA and configured with Device("cuda")
//        (general/device.cpp:554-563).  Every call takes x.Read() / y.Write() / y.ReadWrite() device
//        pointers from MFEM's memory manager (linalg/vector.hpp:520-541): no copies, and
//        GetMemoryClass() reports the device class (fem/bilinearform_ext.hpp:40) so that CGSolver
//        allocates device vectors.  Both libraries work on the legacy default stream.
//   (ii) host-MFEM: Device("cpu") (or a CPU-only MFEM build): Vectors live on the host, the
//        adapter owns device copies and moves x / y across the boundary of each call (the CG
//        adapter moves b in and x out once per solve).
// Errors: a failing library call becomes mfem::mfem_error (MFEM_ABORT), the reference's policy.
#ifndef B200_MFEM_HPP
#define B200_MFEM_HPP

#include <memory>
#include <vector>

#include "mfem.hpp"
#include "fem/lor/lor_batched.hpp"
#include "../include/b200fem.h"

namespace b200
{

#define B200_CHECK(call)                                                               \
   do                                                                                  \
   {                                                                                   \
      if ((call) != B2F_OK) { MFEM_ABORT("libb200fem: " << b2f_last_error()); }        \
   } while (0)

/// true when MFEM's own memory manager hands out CUDA device pointers (mode (i) above)
inline bool DeviceMode() { return mfem::Device::Allows(mfem::Backend::CUDA_MASK); }

/// The process-wide library context: the role of mfem::Device("cuda") (general/device.hpp:128).
class Context
{
   b2f_ctx *ctx = nullptr;
   Context(int dev) { B200_CHECK(b2f_ctx_create(dev, nullptr, &ctx)); }
public:
   ~Context() { b2f_ctx_destroy(ctx); }
   static Context &Get()
   {
      // the CUDA device mfem::Device selected (general/device.hpp:128), or device 0 in host mode
      static Context c(DeviceMode() ? mfem::Device::GetId() : 0);
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

/// One argument of a library call: MFEM's own device pointer in device-resident mode, a device
/// copy owned by the adapter in host mode.
class Bridge
{
   mutable DeviceVector buf;
public:
   const double *Read(const mfem::Vector &v) const
   {
      if (DeviceMode()) { return v.Read(); }
      buf.Upload(v);
      return buf.Ptr();
   }
   double *Write(mfem::Vector &v) const
   {
      if (DeviceMode()) { return v.Write(); }
      buf.Resize(v.Size());
      return buf.Ptr();
   }
   double *ReadWrite(mfem::Vector &v) const
   {
      if (DeviceMode()) { return v.ReadWrite(); }
      buf.Upload(v);
      return buf.Ptr();
   }
   /// after Write / ReadWrite: make the result visible to the host in host mode
   void Commit(mfem::Vector &v) const
   {
      if (!DeviceMode()) { buf.Download(v); }
   }
};

/// Shared implementation of the two integrators: builds b2f_pa from MFEM's own tables.
class PAData
{
public:
   b2f_pa *pa = nullptr;
   int ne = 0, dof = 0, vdim = 1;
   bool symmetric = true; // false: full (non-symmetric) MatrixCoefficient, bilininteg_diffusion_pa.cpp:131
   Bridge bx, by;
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
      Bridge d_nodes;
      const double *nodes_d = d_nodes.Read(nodesE);
      ne = fes.GetNE();
      dof = fe.GetDof();
      vdim = fes.GetVDim();
      b2f_pa_destroy(pa);
      pa = nullptr;
      B200_CHECK(b2f_pa_setup_from_nodes(Context::Get().get(), kind, mesh.Dimension(), maps.ndof,
                                         maps.nqpt, ne, maps.B.HostRead(), maps.G.HostRead(),
                                         ir.GetWeights().HostRead(), nmaps.ndof, nmaps.B.HostRead(),
                                         nmaps.G.HostRead(), nodes_d, coeff, &pa));
      B200_CHECK(b2f_ctx_sync(Context::Get().get())); // nodesE may go out of scope
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
      Bridge d_geo, d_c;
      const double *c_d = d_c.Read(coeff);
      b2f_ctx *ctx = Context::Get().get();
      if (kind == 0)
      {
         const GeometricFactors *geom = mesh.GetGeometricFactors(ir, GeometricFactors::JACOBIANS);
         B200_CHECK(b2f_diffusion_setup(ctx, mesh.Dimension(), maps.ndof, maps.nqpt, ne, maps.B.HostRead(),
                                        maps.G.HostRead(), ir.GetWeights().HostRead(), d_geo.Read(geom->J), c_d,
                                        coeff.GetVDim(), coeff.Size(), &pa));
      }
      else
      {
         const GeometricFactors *geom = mesh.GetGeometricFactors(ir, GeometricFactors::DETERMINANTS);
         B200_CHECK(b2f_mass_setup(ctx, mesh.Dimension(), maps.ndof, maps.nqpt, ne, maps.B.HostRead(),
                                   ir.GetWeights().HostRead(), d_geo.Read(geom->detJ), c_d, coeff.Size(), &pa));
      }
      B200_CHECK(b2f_ctx_sync(ctx));
   }
   void AddMult(const mfem::Vector &x, mfem::Vector &y) const
   {
      const double *xp = bx.Read(x);
      double *yp = by.ReadWrite(y); // AddMultPA accumulates into y (fem/bilininteg.hpp:78)
      B200_CHECK(b2f_pa_add_mult_vec(pa, vdim, xp, yp)); // E-vector (dofs, vdim, NE)
      by.Commit(y);
   }
   /// AddMultTransposePA (bilininteg_diffusion_pa.cpp:95-106): the same kernel for symmetric
   /// coefficients, an abort otherwise -- exactly the reference's behaviour
   void AddMultTranspose(const mfem::Vector &x, mfem::Vector &y) const
   {
      MFEM_VERIFY(symmetric, "DiffusionIntegrator::AddMultTransposePA only implemented in the symmetric case.");
      AddMult(x, y);
   }
   void AddDiagonal(mfem::Vector &diag) const
   {
      if (vdim == 1)
      {
         B200_CHECK(b2f_pa_assemble_diagonal(pa, by.ReadWrite(diag)));
         by.Commit(diag);
         return;
      }
      // vector integrators: the scalar element diagonal added to every component
      mfem::Vector d(ne * dof);
      d.UseDevice(true);
      d = 0.0;
      B200_CHECK(b2f_pa_assemble_diagonal(pa, by.ReadWrite(d)));
      by.Commit(d);
      d.HostRead(); diag.HostReadWrite();
      for (int e = 0; e < ne; e++)
         for (int c = 0; c < vdim; c++)
            for (int i = 0; i < dof; i++) { diag(i + dof * (c + vdim * e)) += d(i + dof * e); }
   }
};

inline double ConstantOf(const mfem::Coefficient *Q)
{
   if (!Q) { return 1.0; }
   auto c = dynamic_cast<const mfem::ConstantCoefficient *>(Q);
   MFEM_VERIFY(c, "b200 vector integrators take a ConstantCoefficient (or none)");
   return c->constant;
}
inline bool IsConstant(const mfem::Coefficient *Q) { return !Q || dynamic_cast<const mfem::ConstantCoefficient *>(Q); }

/// AssemblePA of a diffusion integrator given its three (possibly null) coefficients
/// (bilininteg_diffusion_pa.cpp:108-171)
inline void AssembleDiffusion(PAData &data, const mfem::FiniteElementSpace &fes, const mfem::IntegrationRule &ir,
                              mfem::Coefficient *Q, mfem::VectorCoefficient *VQ, mfem::MatrixCoefficient *MQ)
{
   data.symmetric = (MQ == nullptr);
   if (!VQ && !MQ && IsConstant(Q)) { data.Assemble(0, fes, ir, ConstantOf(Q)); return; }
   // spatially varying scalar, vector (diagonal) or matrix coefficient: the reference's own
   // projection at the quadrature points (bilininteg_diffusion_pa.cpp:155-163)
   mfem::QuadratureSpace qs(*fes.GetMesh(), ir);
   mfem::CoefficientVector coeff(qs, mfem::CoefficientStorage::COMPRESSED);
   if (MQ) { coeff.ProjectTranspose(*MQ); }
   else if (VQ) { coeff.Project(*VQ); }
   else { coeff.Project(*Q); }
   data.AssembleGeneral(0, fes, ir, coeff);
}
inline void AssembleMass(PAData &data, const mfem::FiniteElementSpace &fes, const mfem::IntegrationRule &ir,
                         mfem::Coefficient *Q)
{
   data.symmetric = true;
   if (IsConstant(Q)) { data.Assemble(1, fes, ir, ConstantOf(Q)); return; }
   mfem::QuadratureSpace qs(*fes.GetMesh(), ir);
   mfem::CoefficientVector coeff(Q, qs, mfem::CoefficientStorage::COMPRESSED); // bilininteg_mass_pa.cpp:46
   data.AssembleGeneral(1, fes, ir, coeff);
}

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
      AssembleDiffusion(data, fes, *ir, Q, VQ, MQ);
   }
   void AddMultPA(const mfem::Vector &x, mfem::Vector &y) const override { data.AddMult(x, y); }
   void AddMultTransposePA(const mfem::Vector &x, mfem::Vector &y) const override { data.AddMultTranspose(x, y); }
   void AssembleDiagonalPA(mfem::Vector &diag) override { data.AddDiagonal(diag); }
   b2f_pa *Handle() const { return data.pa; }
   bool Symmetric() const { return data.symmetric; }
   /// the three coefficients of a STOCK mfem::DiffusionIntegrator (protected members read through
   /// pointers to members, which a derived class may form)
   static void Coefficients(const mfem::DiffusionIntegrator &d, mfem::Coefficient *&q,
                            mfem::VectorCoefficient *&vq, mfem::MatrixCoefficient *&mq)
   {
      q = d.*(&DiffusionIntegrator::Q);
      vq = d.*(&DiffusionIntegrator::VQ);
      mq = d.*(&DiffusionIntegrator::MQ);
   }
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
      AssembleMass(data, fes, *ir, Q);
   }
   void AddMultPA(const mfem::Vector &x, mfem::Vector &y) const override { data.AddMult(x, y); }
   void AddMultTransposePA(const mfem::Vector &x, mfem::Vector &y) const override { data.AddMult(x, y); }
   void AssembleDiagonalPA(mfem::Vector &diag) override { data.AddDiagonal(diag); }
   b2f_pa *Handle() const { return data.pa; }
   static mfem::Coefficient *CoefficientOf(const mfem::MassIntegrator &m) { return m.*(&MassIntegrator::Q); }
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

/// MFEM's ElementRestriction tables (fem/restriction.hpp:104-106) on the device, shared by every
/// operator built over the same space.
struct RestrictionHandle
{
   b2f_restriction *r = nullptr;
   explicit RestrictionHandle(const mfem::FiniteElementSpace &fes)
   {
      using namespace mfem;
      auto R = dynamic_cast<const ElementRestriction *>(
                  fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
      MFEM_VERIFY(R, "tensor-product H1 space expected");
      const int dof = fes.GetTypicalFE()->GetDof();
      B200_CHECK(b2f_restriction_create(Context::Get().get(), fes.GetNE(), dof, fes.GetNDofs(), fes.GetVDim(),
                                        fes.GetOrdering() == Ordering::byVDIM, R->GatherMap().HostRead(),
                                        R->Offsets().HostRead(), R->Indices().HostRead(), &r));
   }
   RestrictionHandle(const RestrictionHandle &) = delete;
   ~RestrictionHandle() { b2f_restriction_destroy(r); }
};

/// The library handles of a form's domain integrators.  b200 integrators hand over the handle
/// they built in AssemblePA; STOCK mfem::DiffusionIntegrator / mfem::MassIntegrator objects get a
/// twin built here from their coefficient and rule (owned by `twins`).  Returns false when the
/// form holds anything else (other integrators, boundary or face terms): the caller then leaves
/// the form to MFEM's own extension.
inline bool CollectIntegrators(mfem::BilinearForm &a, std::vector<b2f_pa *> &integ,
                               std::vector<std::unique_ptr<PAData>> &twins, bool &symmetric)
{
   using namespace mfem;
   const FiniteElementSpace &fes = *a.FESpace();
   const FiniteElement &fe = *fes.GetTypicalFE();
   integ.clear();
   symmetric = true;
   if (a.GetBBFI()->Size() || a.GetFBFI()->Size() || a.GetBFBFI()->Size()) { return false; }
   if (!dynamic_cast<const ElementRestriction *>(fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC)))
   {
      return false;
   }
   for (BilinearFormIntegrator *bfi : *a.GetDBFI())
   {
      if (auto d = dynamic_cast<b200::DiffusionIntegrator *>(bfi))
      {
         integ.push_back(d->Handle());
         symmetric = symmetric && d->Symmetric();
      }
      else if (auto m = dynamic_cast<b200::MassIntegrator *>(bfi)) { integ.push_back(m->Handle()); }
      else if (auto vd = dynamic_cast<b200::VectorDiffusionIntegrator *>(bfi)) { integ.push_back(vd->Handle()); }
      else if (auto vm = dynamic_cast<b200::VectorMassIntegrator *>(bfi)) { integ.push_back(vm->Handle()); }
      else if (auto sd = dynamic_cast<mfem::DiffusionIntegrator *>(bfi))
      {
         Coefficient *q; VectorCoefficient *vq; MatrixCoefficient *mq;
         b200::DiffusionIntegrator::Coefficients(*sd, q, vq, mq);
         const IntegrationRule *ir = sd->GetIntegrationRule();
         if (!ir) { ir = &mfem::DiffusionIntegrator::GetRule(fe, fe); }
         twins.emplace_back(new PAData);
         AssembleDiffusion(*twins.back(), fes, *ir, q, vq, mq);
         integ.push_back(twins.back()->pa);
         symmetric = symmetric && twins.back()->symmetric;
      }
      else if (auto sm = dynamic_cast<mfem::MassIntegrator *>(bfi))
      {
         const IntegrationRule *ir = sm->GetIntegrationRule();
         if (!ir) { ir = &mfem::MassIntegrator::GetRule(fe, fe, *fes.GetMesh()->GetTypicalElementTransformation()); }
         twins.emplace_back(new PAData);
         AssembleMass(*twins.back(), fes, *ir, b200::MassIntegrator::CoefficientOf(*sm));
         integ.push_back(twins.back()->pa);
      }
      else { return false; }
   }
   for (b2f_pa *h : integ) { if (!h) { return false; } }
   return !integ.empty();
}

/// seam 1: the fused PA operator of a BilinearForm whose domain integrators are diffusion / mass.
class PAOperator : public mfem::Operator
{
protected:
   std::shared_ptr<RestrictionHandle> restr;
   b2f_operator *op = nullptr;
   std::vector<b2f_pa *> integ;
   std::vector<std::unique_ptr<PAData>> twins;
   bool symmetric = true;
   Bridge bx, by;
   friend class CGSolver;

   void Create(mfem::BilinearForm &a, const mfem::Array<int> &ess_tdof_list)
   {
      using namespace mfem;
      const FiniteElementSpace &fes = *a.FESpace();
      B200_CHECK(b2f_operator_create(Context::Get().get(), restr->r, integ.data(), (int)integ.size(),
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
public:
   /** @param a  a BilinearForm with AssemblyLevel::PARTIAL, already Assemble()d, whose domain
                  integrators are (b200 or stock) diffusion / mass integrators.
       @param ess_tdof_list  essential dofs (FormLinearSystem semantics), may be empty. */
   PAOperator(mfem::BilinearForm &a, const mfem::Array<int> &ess_tdof_list)
      : mfem::Operator(a.FESpace()->GetVSize()), restr(new RestrictionHandle(*a.FESpace()))
   {
      MFEM_VERIFY(CollectIntegrators(a, integ, twins, symmetric),
                  "PAOperator: only domain diffusion / mass integrators on tensor H1 spaces are supported");
      Create(a, ess_tdof_list);
   }
   /// over handles collected before (the extension shares its restriction and integrators)
   PAOperator(mfem::BilinearForm &a, const mfem::Array<int> &ess_tdof_list,
              std::shared_ptr<RestrictionHandle> r, const std::vector<b2f_pa *> &handles, bool symm)
      : mfem::Operator(a.FESpace()->GetVSize()), restr(std::move(r)), integ(handles), symmetric(symm)
   {
      Create(a, ess_tdof_list);
   }
   ~PAOperator() override { b2f_operator_destroy(op); }
   /// vectors of this operator live where MFEM's device lives (fem/bilinearform_ext.hpp:40)
   mfem::MemoryClass GetMemoryClass() const override { return mfem::Device::GetDeviceMemoryClass(); }
   /// PABilinearFormExtension::Mult (fem/bilinearform_ext.cpp:487): y = R^T A_E R x, fused.
   void Mult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      const double *xp = bx.Read(x);
      B200_CHECK(b2f_operator_mult(op, xp, by.Write(y)));
      by.Commit(y);
   }
   /// PABilinearFormExtension::MultTranspose (bilinearform_ext.cpp:566-613) -> AddMultTransposePA:
   /// the same operator for symmetric coefficients; the reference aborts for a full
   /// MatrixCoefficient (bilininteg_diffusion_pa.cpp:95-106) and so does this
   void MultTranspose(const mfem::Vector &x, mfem::Vector &y) const override
   {
      MFEM_VERIFY(symmetric, "DiffusionIntegrator::AddMultTransposePA only implemented in the symmetric case.");
      Mult(x, y);
   }
   /// PABilinearFormExtension::AbsMult (fem/bilinearform_ext.hpp:166)
   void AbsMult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      const double *xp = bx.Read(x);
      B200_CHECK(b2f_operator_abs_mult(op, xp, by.Write(y)));
      by.Commit(y);
   }
   void AssembleDiagonal(mfem::Vector &diag) const override
   {
      diag.SetSize(height);
      B200_CHECK(b2f_operator_assemble_diagonal(op, by.Write(diag)));
      by.Commit(diag);
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
      const double *xp = bx.Read(x);
      B200_CHECK(b2f_operator_mult_constrained(op, xp, by.Write(y)));
      by.Commit(y);
   }
   void MultTranspose(const mfem::Vector &x, mfem::Vector &y) const override
   {
      MFEM_VERIFY(symmetric, "DiffusionIntegrator::AddMultTransposePA only implemented in the symmetric case.");
      Mult(x, y);
   }
   /// ConstrainedOperator::AbsMult (linalg/operator.cpp:663-735) over the extension's
   /// MultInternal(useAbs) (fem/bilinearform_ext.cpp:487-564).
   void AbsMult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      const double *xp = bx.Read(x);
      B200_CHECK(b2f_operator_abs_mult_constrained(op, xp, by.Write(y)));
      by.Commit(y);
   }
   /// Operator::DiagonalPolicy of the constrained rows (default DIAG_ONE as FormLinearSystem).
   void SetDiagonalPolicy(mfem::Operator::DiagonalPolicy p)
   {
      B200_CHECK(b2f_operator_set_diag_policy(op, p == mfem::Operator::DIAG_ZERO ? 0 : p == mfem::Operator::DIAG_ONE ? 1 : 2));
   }
   /// EliminateRHS (linalg/operator.cpp:574): b -= A w (w = x on ess); b[ess] = x[ess].
   void EliminateRHS(const mfem::Vector &x, mfem::Vector &b) const
   {
      const double *xp = bx.Read(x);
      B200_CHECK(b2f_operator_eliminate_rhs(op, xp, by.ReadWrite(b)));
      by.Commit(b);
   }
};

/// seam 2: the extension MFEM's BilinearForm delegates to at AssemblyLevel::PARTIAL
/// (fem/bilinearform_ext.hpp:67-160).  Assemble() lets every integrator run its AssemblePA (stock
/// integrators: MFEM's, b200 integrators: the library's) and then builds the fused operator; Mult,
/// AbsMult, MultTranspose, AssembleDiagonal, FormSystemMatrix and FormLinearSystem go to it.  A
/// form the library does not cover (other integrators, boundary / face terms, non-tensor
/// elements) keeps MFEM's own behaviour: every override falls through to the base class.
class PABilinearFormExtension : public mfem::PABilinearFormExtension
{
   std::shared_ptr<RestrictionHandle> restr;
   std::unique_ptr<PAOperator> fused;
   std::vector<b2f_pa *> integ;
   std::vector<std::unique_ptr<PAData>> twins;
   bool symmetric = true;
   mfem::BilinearForm *form;

   // the conforming serial case FormLinearSystem handles itself (no prolongation)
   bool Plain() const { return fused && !GetProlongation(); }
public:
   explicit PABilinearFormExtension(mfem::BilinearForm *f) : mfem::PABilinearFormExtension(f), form(f) { }
   bool Fused() const { return (bool)fused; }

   void Assemble() override
   {
      fused.reset();
      twins.clear();
      bool all_b200 = form->GetDBFI()->Size() > 0;
      for (mfem::BilinearFormIntegrator *bfi : *form->GetDBFI())
      {
         all_b200 = all_b200 && (dynamic_cast<b200::DiffusionIntegrator *>(bfi) || dynamic_cast<b200::MassIntegrator *>(bfi) ||
                                 dynamic_cast<b200::VectorDiffusionIntegrator *>(bfi) ||
                                 dynamic_cast<b200::VectorMassIntegrator *>(bfi));
      }
      // stock integrators keep their own AssemblePA only when the form stays with MFEM; when the
      // library covers the form their quadrature data is built once, by the library
      if (all_b200) { mfem::PABilinearFormExtension::Assemble(); }
      if (CollectIntegrators(*form, integ, twins, symmetric))
      {
         restr.reset(new RestrictionHandle(*form->FESpace()));
         mfem::Array<int> none;
         fused.reset(new PAOperator(*form, none, restr, integ, symmetric));
      }
      else
      {
         twins.clear();
         if (!all_b200) { mfem::PABilinearFormExtension::Assemble(); }
      }
   }
   void Update() override
   {
      mfem::PABilinearFormExtension::Update();
      fused.reset();
      restr.reset();
      height = width = form->FESpace()->GetVSize();
   }
   void Mult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      if (fused) { fused->Mult(x, y); }
      else { mfem::PABilinearFormExtension::Mult(x, y); }
   }
   void AbsMult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      if (fused) { fused->AbsMult(x, y); }
      else { mfem::PABilinearFormExtension::AbsMult(x, y); }
   }
   void MultTranspose(const mfem::Vector &x, mfem::Vector &y) const override
   {
      if (fused) { fused->MultTranspose(x, y); }
      else { mfem::PABilinearFormExtension::MultTranspose(x, y); }
   }
   void AssembleDiagonal(mfem::Vector &diag) const override
   {
      if (!fused) { mfem::PABilinearFormExtension::AssembleDiagonal(diag); return; }
      const mfem::Operator *P = GetProlongation();
      if (!P) { fused->AssembleDiagonal(diag); return; }
      // bilinearform_ext.cpp:402-431: the local diagonal through |P|^T
      mfem::Vector local(height);
      local.UseDevice(true);
      fused->AssembleDiagonal(local);
      P->AbsMultTranspose(local, diag);
   }
   /// FormSystemMatrix / FormLinearSystem (bilinearform_ext.cpp:460-485 -> linalg/operator.cpp:
   /// 130-146): for a conforming serial space the constrained operator is the library's fused one
   /// (z = x; z[ess] = 0; y = A z; y[ess] = x[ess] in the same kernel); otherwise MFEM's
   /// RAP + ConstrainedOperator wrap this extension's fused Mult.
   void FormSystemMatrix(const mfem::Array<int> &ess_tdof_list, mfem::OperatorHandle &A) override
   {
      if (!Plain()) { mfem::PABilinearFormExtension::FormSystemMatrix(ess_tdof_list, A); return; }
      // (DIAG_ONE: Operator::FormConstrainedSystemOperator's default, linalg/operator.cpp:227-236)
      A.Reset(new ConstrainedOperator(*form, ess_tdof_list, restr, integ, symmetric)); // A owns it
   }
   void FormLinearSystem(const mfem::Array<int> &ess_tdof_list, mfem::Vector &x, mfem::Vector &b,
                         mfem::OperatorHandle &A, mfem::Vector &X, mfem::Vector &B, int copy_interior = 0) override
   {
      if (!Plain())
      {
         mfem::PABilinearFormExtension::FormLinearSystem(ess_tdof_list, x, b, A, X, B, copy_interior);
         return;
      }
      // Operator::InitTVectors without P: X and B alias x and b (linalg/operator.cpp:104-127)
      X.MakeRef(x, 0, x.Size());
      B.MakeRef(b, 0, b.Size());
      if (!copy_interior) { X.SetSubVectorComplement(ess_tdof_list, 0.0); }
      auto *c = new ConstrainedOperator(*form, ess_tdof_list, restr, integ, symmetric);
      c->EliminateRHS(X, B);
      A.Reset(c);
   }
};

/// mfem::BilinearForm whose PARTIAL assembly level installs the extension above -- the three lines
/// a maintainer would otherwise put into BilinearForm::SetAssemblyLevel (fem/bilinearform.cpp:
/// 127-129).  Everything else is the base class: `b200::BilinearForm a(&fespace);` is the only
/// change in examples/ex1.cpp:203-268.
class BilinearForm : public mfem::BilinearForm
{
public:
   using mfem::BilinearForm::BilinearForm;
   void SetAssemblyLevel(mfem::AssemblyLevel level)
   {
      if (level != mfem::AssemblyLevel::PARTIAL) { mfem::BilinearForm::SetAssemblyLevel(level); return; }
      MFEM_VERIFY(!ext, "the assembly level has already been set!");
      assembly = level;
      ext.reset(new b200::PABilinearFormExtension(this));
   }
   bool Fused() const
   {
      auto e = dynamic_cast<const b200::PABilinearFormExtension *>(ext.get());
      return e && e->Fused();
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
      Bridge d_detj, d_coeff, d_b;
      const double *detj = d_detj.Read(geom->detJ), *cf = d_coeff.Read(coeff);
      const int *mk = nullptr;
      int *d_mk = nullptr;
      if (DeviceMode()) { mk = markers.Read(); }
      else
      {
         B200_CHECK(b2f_malloc(ctx, sizeof(int) * (markers.Size() + 1), (void **)&d_mk));
         B200_CHECK(b2f_memcpy_h2d(ctx, d_mk, markers.HostRead(), sizeof(int) * markers.Size()));
         mk = d_mk;
      }
      B200_CHECK(b2f_domain_lf_assemble(ctx, mesh->Dimension(), maps.ndof, maps.nqpt, fes.GetVDim(), fes.GetNE(),
                                        mk, maps.B.HostRead(), ir->GetWeights().HostRead(), detj,
                                        cf, coeff.Size(), d_b.ReadWrite(b)));
      d_b.Commit(b);
      B200_CHECK(b2f_ctx_sync(ctx));
      if (d_mk) { b2f_free(ctx, d_mk); }
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
   Bridge bx, by;
public:
   OperatorChebyshevSmoother(const ConstrainedOperator &A_, const mfem::Vector &d, int order,
                             double max_eig_estimate)
      : mfem::Solver(d.Size()), A(A_), max_eig(max_eig_estimate)
   {
      Bridge dd;
      B200_CHECK(b2f_chebyshev_create(A.Handle(), dd.Read(d), order, max_eig, &h));
      B200_CHECK(b2f_ctx_sync(Context::Get().get()));
   }
   OperatorChebyshevSmoother(const ConstrainedOperator &A_, const mfem::Vector &d,
                             const mfem::Array<int> &ess_tdof_list, int order, int power_iterations = 10,
                             double power_tolerance = 1e-8, int power_seed = 12345)
      : mfem::Solver(d.Size()), A(A_)
   {
      MFEM_VERIFY(power_seed != 0, "invalid seed!");
      b2f_ctx *ctx = Context::Get().get();
      Bridge dd;
      DeviceVector dinv(d.Size()), ev;
      const double *dp = dd.Read(d);
      int *d_ess = nullptr;
      B200_CHECK(b2f_malloc(ctx, sizeof(int) * (ess_tdof_list.Size() + 1), (void **)&d_ess));
      B200_CHECK(b2f_memcpy_h2d(ctx, d_ess, ess_tdof_list.HostRead(), sizeof(int) * ess_tdof_list.Size()));
      B200_CHECK(b2f_jacobi_setup(ctx, d.Size(), dp, d_ess, ess_tdof_list.Size(), 1.0, dinv.Ptr()));
      mfem::Vector v0(d.Size());
      v0.Randomize(power_seed);
      ev.Upload(v0);
      B200_CHECK(b2f_power_method(A.Handle(), dinv.Ptr(), ev.Ptr(), power_iterations, power_tolerance,
                                  &max_eig));
      b2f_free(ctx, d_ess);
      B200_CHECK(b2f_chebyshev_create(A.Handle(), dp, order, max_eig, &h));
      B200_CHECK(b2f_ctx_sync(ctx));
   }
   ~OperatorChebyshevSmoother() override { b2f_chebyshev_destroy(h); }
   void SetOperator(const mfem::Operator &) override { }
   double MaxEigEstimate() const { return max_eig; }
   b2f_chebyshev *Handle() const { return h; }
   void Mult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      MFEM_VERIFY(!iterative_mode, "Chebyshev smoother not implemented for iterative mode");
      const double *xp = bx.Read(x);
      B200_CHECK(b2f_chebyshev_mult(h, xp, by.Write(y)));
      by.Commit(y);
   }
};

/// CGSolver (linalg/solvers.cpp:869-1050) with an optional Jacobi or Chebyshev preconditioner,
/// entirely on the device (device-resident scalars, one host poll per solve or per `check_every`).
/// SetPreconditioner takes MFEM's own OperatorJacobiSmoother (solvers.cpp:324-450) -- the object
/// ex1 builds with `OperatorJacobiSmoother M(a, ess_tdof_list)` -- or a b200 Chebyshev smoother.
class CGSolver : public mfem::IterativeSolver
{
   const ConstrainedOperator *A = nullptr;
   bool jacobi = false;
   const OperatorChebyshevSmoother *cheb = nullptr;
   DeviceVector dinv;
   Bridge bb, bxs;
public:
   void SetOperator(const mfem::Operator &op_) override
   {
      A = dynamic_cast<const ConstrainedOperator *>(&op_);
      MFEM_VERIFY(A, "b200::CGSolver needs a b200::ConstrainedOperator (what FormLinearSystem of a "
                  "b200::BilinearForm returns)");
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
   void SetPreconditioner(mfem::Solver &pr) override
   {
      MFEM_VERIFY(A, "SetOperator first");
      if (auto c = dynamic_cast<const OperatorChebyshevSmoother *>(&pr))
      {
         cheb = c;
         jacobi = false;
         return;
      }
      auto js = dynamic_cast<mfem::OperatorJacobiSmoother *>(&pr);
      MFEM_VERIFY(js, "b200::CGSolver: the preconditioner must be an mfem::OperatorJacobiSmoother or a "
                  "b200::OperatorChebyshevSmoother");
      // the smoother's action on the vector of ones is its scaled inverse diagonal (dinv is private,
      // linalg/solvers.hpp:481; Mult is y = dinv .* x, solvers.cpp:425-450)
      mfem::Vector ones(height), d(height);
      ones.UseDevice(true); d.UseDevice(true);
      ones = 1.0;
      const bool it = js->iterative_mode;
      js->iterative_mode = false;
      js->Mult(ones, d);
      js->iterative_mode = it;
      dinv.Upload(d);
      B200_CHECK(b2f_ctx_sync(Context::Get().get()));
      jacobi = true;
      cheb = nullptr;
      prec = &pr;
   }
   void Mult(const mfem::Vector &b, mfem::Vector &x) const override
   {
      MFEM_VERIFY(A, "operator not set");
      b2f_ctx *ctx = Context::Get().get();
      DeviceVector rhs, x0, ax0;
      const double *bp = bb.Read(b);
      if (iterative_mode)
      {
         // r = b - A x0; the library iterates on the correction from a zero start, which
         // generates the same Krylov iterates as CGSolver::Mult does from x0
         x0.Upload(x);
         ax0.Resize(height);
         rhs.Resize(height);
         B200_CHECK(b2f_operator_mult_constrained(A->Handle(), x0.Ptr(), ax0.Ptr()));
         B200_CHECK(b2f_vector_add(ctx, height, bp, -1.0, ax0.Ptr(), rhs.Ptr()));
         bp = rhs.Ptr();
      }
      double *xp = bxs.Write(x);
      b2f_pcg_options opt = {rel_tol, abs_tol, max_iter, 1};
      b2f_pcg_result res;
      if (cheb) { B200_CHECK(b2f_pcg_chebyshev(A->Handle(), cheb->Handle(), bp, xp, &opt, &res, nullptr, 0)); }
      else { B200_CHECK(b2f_pcg(A->Handle(), jacobi ? dinv.Ptr() : nullptr, bp, xp, &opt, &res, nullptr, 0)); }
      if (iterative_mode) { B200_CHECK(b2f_vector_add(ctx, height, x0.Ptr(), 1.0, xp, xp)); }
      bxs.Commit(x);
      B200_CHECK(b2f_ctx_sync(ctx));
      final_iter = res.iterations;
      converged = res.converged != 0;
      initial_norm = res.initial_norm;
      final_norm = res.final_norm;
   }
};

/// mfem::PCG (linalg/solvers.cpp:1067-1080; tolerances are squared) on the device: the call ex1
/// makes, `PCG(*A, M, B, X, 1, 200, 1e-12, 0.0)`, with A from a b200::BilinearForm.  An operator
/// that is not the library's runs MFEM's own PCG.
inline void PCG(const mfem::Operator &A, mfem::Solver &B, const mfem::Vector &b, mfem::Vector &x,
                int print_iter = 0, int max_num_iter = 1000, mfem::real_t RTOLERANCE = 1e-12,
                mfem::real_t ATOLERANCE = 1e-24)
{
   if (!dynamic_cast<const ConstrainedOperator *>(&A))
   {
      mfem::PCG(A, B, b, x, print_iter, max_num_iter, RTOLERANCE, ATOLERANCE);
      return;
   }
   CGSolver pcg;
   pcg.SetPrintLevel(-1);
   pcg.SetMaxIter(max_num_iter);
   pcg.SetRelTol(std::sqrt(RTOLERANCE));
   pcg.SetAbsTol(std::sqrt(ATOLERANCE));
   pcg.SetOperator(A);
   pcg.SetPreconditioner(B);
   pcg.Mult(b, x);
   if (print_iter)
   {
      mfem::out << "b200::PCG: " << pcg.GetNumIterations() << " iterations, (B r, r)^1/2 = " << pcg.GetFinalNorm()
                << (pcg.GetConverged() ? "" : " (not converged)") << '\n';
   }
}

/// LOR matrix assembly for H1 spaces in 3D: drop-in for mfem::BatchedLORAssembly::Assemble
/// (fem/lor/lor_batched.cpp:712-734; H1 kernel fem/lor/lor_h1_impl.hpp:184-455) -- what
/// LORDiscretization / LORSolver<T> call to obtain the SparseMatrix they hand to T (e.g. an AMG).
/// The vertex coordinates and the projected coefficients are formed by MFEM's own routines
/// (FormLORVertexCoordinates, ProjectLORCoefficient), exactly as BatchedLOR_H1's constructor does;
/// the stencil array, the CSR conversion and the boundary elimination run in the library.
/// The matrix comes back with the rows and the columns inside each row in the order of MFEM's serial
/// host loop, on the device in device mode (no host copy), on the host otherwise.
class BatchedLORAssembly
{
   mfem::FiniteElementSpace &fes_ho;
   mfem::Vector X_vert;
public:
   explicit BatchedLORAssembly(mfem::FiniteElementSpace &fes) : fes_ho(fes)
   {
      mfem::BatchedLORAssembly::FormLORVertexCoordinates(fes_ho, X_vert);
   }
   const mfem::Vector &GetLORVertexCoordinates() const { return X_vert; }

   /// H1, 3D, tensor elements, MassIntegrator / DiffusionIntegrator with scalar coefficients only
   static bool FormIsSupported(mfem::BilinearForm &a)
   {
      using namespace mfem;
      const FiniteElementSpace &fes = *a.FESpace();
      if (!dynamic_cast<const H1_FECollection *>(fes.FEColl()) || fes.GetMesh()->Dimension() != 3 ||
          fes.GetMesh()->SpaceDimension() != 3 || fes.GetVDim() != 1 || !UsesTensorBasis(fes) ||
          fes.GetMaxElementOrder() > 8 || fes.GetConformingProlongation())
      {
         return false;
      }
      if (a.GetBBFI()->Size() || a.GetFBFI()->Size() || a.GetBFBFI()->Size()) { return false; }
      for (BilinearFormIntegrator *bfi : *a.GetDBFI())
      {
         auto d = dynamic_cast<mfem::DiffusionIntegrator *>(bfi);
         if (d)
         {
            Coefficient *q; VectorCoefficient *vq; MatrixCoefficient *mq;
            b200::DiffusionIntegrator::Coefficients(*d, q, vq, mq);
            if (vq || mq) { return false; } // the reference projects the scalar coefficient only
         }
         else if (!dynamic_cast<mfem::MassIntegrator *>(bfi)) { return false; }
      }
      return true;
   }

   void Assemble(mfem::BilinearForm &a, const mfem::Array<int> &ess_dofs, mfem::OperatorHandle &A)
   {
      using namespace mfem;
      MFEM_VERIFY(FormIsSupported(a), "b200::BatchedLORAssembly: H1 mass + diffusion forms in 3D only");
      IntegrationRule ir = GetCollocatedIntRule(fes_ho);
      QuadratureSpace qs(*fes_ho.GetMesh(), ir);
      CoefficientVector c1(qs, CoefficientStorage::COMPRESSED), c2(qs, CoefficientStorage::COMPRESSED);
      ProjectLORCoefficient<mfem::MassIntegrator>(a, c1);
      ProjectLORCoefficient<mfem::DiffusionIntegrator>(a, c2);
      b2f_ctx *ctx = Context::Get().get();
      RestrictionHandle R(fes_ho);
      Bridge bx, b1, b2;
      const int n_ess = ess_dofs.Size();
      const int *ess = nullptr;
      int *d_ess = nullptr;
      if (DeviceMode()) { ess = ess_dofs.Read(); }
      else if (n_ess)
      {
         B200_CHECK(b2f_malloc(ctx, sizeof(int) * n_ess, (void **)&d_ess));
         B200_CHECK(b2f_memcpy_h2d(ctx, d_ess, ess_dofs.HostRead(), sizeof(int) * n_ess));
         ess = d_ess;
      }
      b2f_lor *L = nullptr;
      B200_CHECK(b2f_lor_assemble_h1(ctx, R.r, 3, fes_ho.GetMaxElementOrder(), bx.Read(X_vert), b1.Read(c1), c1.Size(),
                                     b2.Read(c2), c2.Size(), ess, n_ess, 0, &L));
      if (d_ess) { b2f_free(ctx, d_ess); }
      int n = 0;
      int64_t nnz = 0;
      const int *I = nullptr, *J = nullptr;
      const double *V = nullptr;
      B200_CHECK(b2f_lor_csr(L, &n, &nnz, &I, &J, &V));
      SparseMatrix *M = A.Is<SparseMatrix>();
      if (!M) { M = new SparseMatrix; A.Reset(M); }
      M->OverrideSize(n, n);
      EnsureCapacity(M->GetMemoryI(), n + 1);
      EnsureCapacity(M->GetMemoryJ(), (int)nnz);
      EnsureCapacity(M->GetMemoryData(), (int)nnz);
      if (DeviceMode())
      {
         B200_CHECK(b2f_memcpy_d2d(ctx, M->WriteI(), I, sizeof(int) * (n + 1)));
         B200_CHECK(b2f_memcpy_d2d(ctx, M->WriteJ(), J, sizeof(int) * nnz));
         B200_CHECK(b2f_memcpy_d2d(ctx, M->WriteData(), V, sizeof(double) * nnz));
      }
      else
      {
         B200_CHECK(b2f_memcpy_d2h(ctx, M->HostWriteI(), I, sizeof(int) * (n + 1)));
         B200_CHECK(b2f_memcpy_d2h(ctx, M->HostWriteJ(), J, sizeof(int) * nnz));
         B200_CHECK(b2f_memcpy_d2h(ctx, M->HostWriteData(), V, sizeof(double) * nnz));
      }
      B200_CHECK(b2f_ctx_sync(ctx));
      b2f_lor_destroy(L);
   }
};

} // namespace b200
#endif
