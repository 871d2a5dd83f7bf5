// mfem_b200/csrc/host_mesh.cpp -- structured quad/hex meshes and H1 dof numbering with the
// reference's conventions (host C++; produces the integer tables the device path consumes).
//
// What is reproduced (so that L-vectors are numbered exactly as the reference numbers them,
// and x.Randomize(seed) / golden norms / PCG histories are comparable entry by entry):
//   Mesh::Make2D / Make3D vertex + element generation          mesh/mesh.cpp:3841,4388
//   the generalised Hilbert element order (sfc_ordering=true)  mesh/ncmesh.cpp:5377-5640
//   edge numbering = first encounter over elements x local edges   mesh/mesh.cpp:8526-8560
//   face numbering = first encounter over elements x local faces   mesh/mesh.cpp:8996-9060
//   edge / face orientations                                   mesh/mesh.cpp:8044-8075,7586,8741
//   H1 dof layout: vertices | edges | faces | interiors         fem/fespace.cpp:3428-3536
//   lexicographic -> native element dof map                    fem/fe/fe_base.cpp:2497-2680
//
// Formulation: everything is derived geometrically from the reference-element vertex
// coordinates and its local edge / face vertex lists: an entity dof is a lattice point
// p*c(v0) + i*(c(v1)-c(v0)) + j*(c(v3)-c(v0)); the orientation of a shared edge / face is
// applied by expressing that lattice point in the parametrisation of the entity's first
// (owning) element.
#include <algorithm>
#include <exception>
#include <array>
#include <cstdint>
#include <cstdlib>
#include <unordered_map>
#include <vector>

#include "../../include/b200fem.h"
#include "host_parallel.h"

namespace b2f
{
void set_error(const char *fmt, ...);

namespace
{

struct V3
{
   int x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline int extent(V3 a) { return std::abs(a.x + a.y + a.z); }
inline int sgn(int v) { return (v > 0) - (v < 0); }
inline V3 unit(V3 a) { return {sgn(a.x), sgn(a.y), sgn(a.z)}; }
inline V3 half(V3 a) { return {a.x / 2, a.y / 2, a.z / 2}; }

// Generalised Hilbert curve over a w x h (x d) box of arbitrary size (the "gilbert"
// construction): the box is described by an origin and axis vectors a (major), b, c.
struct Gilbert
{
   std::vector<V3> out;

   void line(V3 o, V3 step, int n)
   {
      for (int i = 0; i < n; i++, o = o + step) { out.push_back(o); }
   }

   void run2(V3 o, V3 a, V3 b)
   {
      const int w = extent(a), h = extent(b);
      const V3 da = unit(a), db = unit(b);
      if (h == 1) { line(o, da, w); return; }
      if (w == 1) { line(o, db, h); return; }
      V3 a2 = half(a), b2 = half(b);
      const int w2 = extent(a2), h2 = extent(b2);
      if (2 * w > 3 * h)
      {
         if ((w2 & 1) && w > 2) { a2 = a2 + da; }
         run2(o, a2, b);
         run2(o + a2, a - a2, b);
      }
      else
      {
         if ((h2 & 1) && h > 2) { b2 = b2 + db; }
         run2(o, b2, a2);
         run2(o + b2, a, b - b2);
         run2(o + (a - da) + (b2 - db), -b2, -(a - a2));
      }
   }

   void run3(V3 o, V3 a, V3 b, V3 c)
   {
      const int w = extent(a), h = extent(b), d = extent(c);
      const V3 da = unit(a), db = unit(b), dc = unit(c);
      if (h == 1 && d == 1) { line(o, da, w); return; }
      if (w == 1 && d == 1) { line(o, db, h); return; }
      if (w == 1 && h == 1) { line(o, dc, d); return; }
      V3 a2 = half(a), b2 = half(b), c2 = half(c);
      const int w2 = extent(a2), h2 = extent(b2), d2 = extent(c2);
      if ((w2 & 1) && w > 2) { a2 = a2 + da; }
      if ((h2 & 1) && h > 2) { b2 = b2 + db; }
      if ((d2 & 1) && d > 2) { c2 = c2 + dc; }
      if (2 * w > 3 * h && 2 * w > 3 * d) // wide: split along a only
      {
         run3(o, a2, b, c);
         run3(o + a2, a - a2, b, c);
      }
      else if (3 * h > 4 * d) // do not split c
      {
         run3(o, b2, c, a2);
         run3(o + b2, a, b - b2, c);
         run3(o + (a - da) + (b2 - db), -b2, c, -(a - a2));
      }
      else if (3 * d > 4 * h) // do not split b
      {
         run3(o, c2, a2, b);
         run3(o + c2, a, b, c - c2);
         run3(o + (a - da) + (c2 - dc), -c2, -(a - a2), b);
      }
      else // split all three
      {
         run3(o, b2, c2, a2);
         run3(o + b2, c, a2, b - b2);
         run3(o + (b2 - db) + (c - dc), a, -b2, -(c - c2));
         run3(o + (a - da) + b2 + (c - dc), -c, -(a - a2), b - b2);
         run3(o + (a - da) + (b2 - db), -b2, c2, -(a - a2));
      }
   }
};

std::vector<V3> grid_order(int dim, int nx, int ny, int nz, bool sfc)
{
   std::vector<V3> ord;
   if (!sfc)
   {
      ord.reserve((size_t)nx * ny * nz);
      for (int z = 0; z < nz; z++)
         for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) { ord.push_back({x, y, z}); }
      return ord;
   }
   Gilbert g;
   g.out.reserve((size_t)nx * ny * nz);
   if (dim == 2)
   {
      if (nx >= ny) { g.run2({0, 0, 0}, {nx, 0, 0}, {0, ny, 0}); }
      else { g.run2({0, 0, 0}, {0, ny, 0}, {nx, 0, 0}); }
   }
   else
   {
      if (nx >= ny && nx >= nz) { g.run3({0, 0, 0}, {nx, 0, 0}, {0, ny, 0}, {0, 0, nz}); }
      else if (ny >= nx && ny >= nz) { g.run3({0, 0, 0}, {0, ny, 0}, {nx, 0, 0}, {0, 0, nz}); }
      else { g.run3({0, 0, 0}, {0, 0, nz}, {nx, 0, 0}, {0, ny, 0}); }
   }
   return std::move(g.out);
}

// reference-element tables (fem/geom.cpp:960,1020-1036): vertex coordinates, edges, faces
const int kQuadVert[4][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}};
const int kQuadEdge[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
const int kHexVert[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                            {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
const int kHexEdge[12][2] = {{0, 1}, {1, 2}, {3, 2}, {0, 3}, {4, 5}, {5, 6},
                             {7, 6}, {4, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
const int kHexFace[6][4] = {{3, 2, 1, 0}, {0, 1, 5, 4}, {1, 2, 6, 5},
                            {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};

struct PairHash
{
   size_t operator()(uint64_t k) const
   {
      k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33;
      return (size_t)k;
   }
};

struct FaceKey
{
   int a, b, c; // the three smallest vertex ids, ascending
   bool operator==(const FaceKey &o) const { return a == o.a && b == o.b && c == o.c; }
};
struct FaceHash
{
   size_t operator()(const FaceKey &k) const
   {
      uint64_t h = (uint64_t)k.a * 0x9E3779B97F4A7C15ULL;
      h ^= ((uint64_t)k.b + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2));
      h ^= ((uint64_t)k.c + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2));
      return (size_t)h;
   }
};

} // namespace

// lexicographic -> native element dof map of the order-p H1 tensor element
void h1_dof_map(int dim, int p, int *dof_map)
{
   const int p1 = p + 1;
   const int nv = dim == 2 ? 4 : 8;
   const int(*vc)[3] = dim == 2 ? kQuadVert : kHexVert;
   auto lex = [&](int i, int j, int k) { return i + p1 * (j + p1 * k); };
   int o = 0;
   for (int v = 0; v < nv; v++) { dof_map[lex(p * vc[v][0], p * vc[v][1], p * vc[v][2])] = o++; }
   const int ne = dim == 2 ? 4 : 12;
   for (int e = 0; e < ne; e++)
   {
      const int *ev = dim == 2 ? kQuadEdge[e] : kHexEdge[e];
      const int *a = vc[ev[0]], *b = vc[ev[1]];
      for (int i = 1; i < p; i++)
      {
         dof_map[lex(p * a[0] + i * (b[0] - a[0]), p * a[1] + i * (b[1] - a[1]),
                     p * a[2] + i * (b[2] - a[2]))] = o++;
      }
   }
   if (dim == 3)
   {
      for (int f = 0; f < 6; f++)
      {
         const int *a = vc[kHexFace[f][0]], *b = vc[kHexFace[f][1]], *c = vc[kHexFace[f][3]];
         for (int j = 1; j < p; j++)
            for (int i = 1; i < p; i++)
            {
               dof_map[lex(p * a[0] + i * (b[0] - a[0]) + j * (c[0] - a[0]),
                           p * a[1] + i * (b[1] - a[1]) + j * (c[1] - a[1]),
                           p * a[2] + i * (b[2] - a[2]) + j * (c[2] - a[2]))] = o++;
            }
      }
      for (int k = 1; k < p; k++)
         for (int j = 1; j < p; j++)
            for (int i = 1; i < p; i++) { dof_map[lex(i, j, k)] = o++; }
   }
   else
   {
      for (int j = 1; j < p; j++)
         for (int i = 1; i < p; i++) { dof_map[lex(i, j, 0)] = o++; }
   }
}

struct H1Space
{
   int ne = 0, ndofs = 0;
   std::vector<int> elem_dof;  // native order
   std::vector<int> bdr_dofs;  // sorted
};

// H1 numbering of a conforming quad/hex mesh given element -> vertex lists (reference
// vertex order), for any element ordering.
// `vstride` (optional): vertex-id strides {1, sy, sz} of a structured grid.  With it edges and
// faces are looked up by direct addressing (slot = 3*lowest vertex + direction) instead of
// hashing -- same first-encounter numbering, several times faster on 10^6..10^8 elements.
void h1_number(int dim, int p, int nverts, const std::vector<int> &elem_vert, H1Space &S,
               const int *vstride = nullptr)
{
   const int nv = dim == 2 ? 4 : 8, nedge = dim == 2 ? 4 : 12, nface = dim == 2 ? 0 : 6;
   const int ne = (int)(elem_vert.size() / nv);
   const int pm1 = p - 1;
   const int dof = dim == 2 ? (p + 1) * (p + 1) : (p + 1) * (p + 1) * (p + 1);
   // --- edges: first-encounter numbering
   std::unordered_map<uint64_t, int, PairHash> edge_id;
   std::vector<int> edge_slot; // direct table: 3*min vertex + direction
   if (vstride) { edge_slot.assign((size_t)3 * nverts, -1); }
   else { edge_id.reserve((size_t)ne * 4); }
   // (order 1 in 3D has no edge or face dofs: the edge numbering is not needed at all, and on a
   // structured grid neither is the face numbering -- the boundary dofs are the boundary vertices)
   const bool need_edges = pm1 > 0 || dim == 2;
   const bool need_faces = dim == 3 && (pm1 > 0 || !vstride);
   std::vector<int> el_edge(need_edges ? (size_t)ne * nedge : 0);
   int nedges = 0;
   for (int e = 0; need_edges && e < ne; e++)
   {
      const int *v = &elem_vert[(size_t)e * nv];
      for (int j = 0; j < nedge; j++)
      {
         const int *ev = dim == 2 ? kQuadEdge[j] : kHexEdge[j];
         int a = v[ev[0]], b = v[ev[1]];
         if (a > b) { std::swap(a, b); }
         int id;
         if (vstride)
         {
            const int d = b - a;
            int &slot = edge_slot[(size_t)3 * a + (d == vstride[0] ? 0 : d == vstride[1] ? 1 : 2)];
            if (slot < 0) { slot = nedges++; }
            id = slot;
         }
         else
         {
            const uint64_t key = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
            auto it = edge_id.find(key);
            if (it == edge_id.end()) { id = nedges++; edge_id.emplace(key, id); }
            else { id = it->second; }
         }
         el_edge[(size_t)e * nedge + j] = id;
      }
   }
   std::vector<int>().swap(edge_slot);
   // edge -> number of incident elements is not needed; boundary comes from faces (3D) or
   // from edges with a single element (2D)
   std::vector<int> edge_count;
   if (dim == 2)
   {
      edge_count.assign(nedges, 0);
      for (size_t i = 0; i < el_edge.size(); i++) { edge_count[el_edge[i]]++; }
   }
   // --- faces: first-encounter numbering; the first element fixes the face's frame
   int nfaces = 0;
   std::vector<int> el_face, face_count;
   std::vector<std::array<int, 4>> face_base; // vertices of the owning element's local face
   if (need_faces)
   {
      std::unordered_map<FaceKey, int, FaceHash> face_id;
      std::vector<int> face_slot; // direct table: 3*min vertex + normal direction
      if (vstride) { face_slot.assign((size_t)3 * nverts, -1); }
      else { face_id.reserve((size_t)ne * 3 + 16); }
      el_face.resize((size_t)ne * 6);
      for (int e = 0; e < ne; e++)
      {
         const int *v = &elem_vert[(size_t)e * nv];
         for (int j = 0; j < 6; j++)
         {
            std::array<int, 4> fv = {v[kHexFace[j][0]], v[kHexFace[j][1]], v[kHexFace[j][2]],
                                     v[kHexFace[j][3]]};
            std::array<int, 4> s = fv;
            std::sort(s.begin(), s.end());
            int id;
            bool fresh;
            if (vstride)
            {
               // the two smallest offsets from the lowest vertex tell the face's normal
               const int d1 = s[1] - s[0];
               const int nrm = d1 == vstride[1] ? 0 : (s[2] - s[0] == vstride[2] ? 1 : 2);
               int &slot = face_slot[(size_t)3 * s[0] + nrm];
               fresh = slot < 0;
               if (fresh) { slot = nfaces; }
               id = slot;
            }
            else
            {
               const FaceKey key{s[0], s[1], s[2]};
               auto it = face_id.find(key);
               fresh = it == face_id.end();
               if (fresh) { id = nfaces; face_id.emplace(key, id); }
               else { id = it->second; }
            }
            if (fresh)
            {
               nfaces++;
               face_base.push_back(fv);
               face_count.push_back(1);
            }
            else { face_count[id]++; }
            el_face[(size_t)e * 6 + j] = id;
         }
      }
   }
   // --- dof layout
   const int nvdofs = nverts;
   const int nedofs = nedges * pm1;
   const int nfdofs = nfaces * pm1 * pm1;
   const int nb = dim == 2 ? pm1 * pm1 : pm1 * pm1 * pm1;
   S.ne = ne;
   S.ndofs = nvdofs + nedofs + nfdofs + ne * nb;
   S.elem_dof.resize((size_t)ne * dof);
   parallel_for(ne, [&](int64_t e_begin, int64_t e_end)
   {
   for (int e = (int)e_begin; e < (int)e_end; e++)
   {
      const int *v = &elem_vert[(size_t)e * nv];
      int *out = &S.elem_dof[(size_t)e * dof];
      int o = 0;
      for (int i = 0; i < nv; i++) { out[o++] = v[i]; }
      for (int j = 0; pm1 > 0 && j < nedge; j++)
      {
         const int *ev = dim == 2 ? kQuadEdge[j] : kHexEdge[j];
         const bool fwd = v[ev[0]] < v[ev[1]];
         const int base = nvdofs + el_edge[(size_t)e * nedge + j] * pm1;
         for (int i = 0; i < pm1; i++) { out[o++] = base + (fwd ? i : pm1 - 1 - i); }
      }
      for (int j = 0; pm1 > 0 && j < nface; j++)
      {
         const int f = el_face[(size_t)e * 6 + j];
         const std::array<int, 4> &bv = face_base[f];
         // owning frame: base vertex k sits at lattice corner (ca[k], cb[k]) * p
         static const int ca[4] = {0, 1, 1, 0}, cb[4] = {0, 0, 1, 1};
         int pos[4]; // position of this element's local face vertex k in the base list
         for (int k = 0; k < 4; k++)
         {
            const int gv = v[kHexFace[j][k]];
            pos[k] = (bv[0] == gv) ? 0 : (bv[1] == gv) ? 1 : (bv[2] == gv) ? 2 : 3;
         }
         const int a0 = ca[pos[0]], b0 = cb[pos[0]];
         const int ai = ca[pos[1]] - a0, bi = cb[pos[1]] - b0; // direction of local i
         const int aj = ca[pos[3]] - a0, bj = cb[pos[3]] - b0; // direction of local j
         const int base = nvdofs + nedofs + f * pm1 * pm1;
         for (int jj = 1; jj < p; jj++)
            for (int ii = 1; ii < p; ii++)
            {
               const int a = p * a0 + ii * ai + jj * aj;
               const int b = p * b0 + ii * bi + jj * bj;
               out[o++] = base + (a - 1) + (b - 1) * pm1;
            }
      }
      const int base = nvdofs + nedofs + nfdofs + e * nb;
      for (int i = 0; i < nb; i++) { out[o++] = base + i; }
   }
   });
   // --- boundary dofs (all boundary attributes essential, as ex1.cpp:165-183 / bench_ceed)
   std::vector<char> mark(S.ndofs, 0);
   if (dim == 3 && !need_faces)
   {
      // structured grid, order 1: the boundary vertices, from their grid coordinates
      const int sy = vstride[1], sz = vstride[2], ny1 = sz / sy, nz1 = nverts / sz;
      parallel_for(nverts, [&](int64_t b, int64_t e_)
      {
         for (int64_t id = b; id < e_; id++)
         {
            const int x = (int)(id % sy), y = (int)((id / sy) % ny1), z = (int)(id / sz);
            mark[id] = x == 0 || x == sy - 1 || y == 0 || y == ny1 - 1 || z == 0 || z == nz1 - 1;
         }
      });
   }
   for (int e = 0; (dim == 2 || need_faces) && e < ne; e++)
   {
      const int *v = &elem_vert[(size_t)e * nv];
      const int *ed = &S.elem_dof[(size_t)e * dof];
      if (dim == 3)
      {
         for (int j = 0; j < 6; j++)
         {
            const int f = el_face[(size_t)e * 6 + j];
            if (face_count[f] != 1) { continue; }
            for (int k = 0; k < 4; k++) { mark[v[kHexFace[j][k]]] = 1; }
            // the face's four edges: local edges whose two vertices both lie on the face
            for (int l = 0; l < 12; l++)
            {
               int hits = 0;
               for (int k = 0; k < 4; k++)
               {
                  hits += (kHexEdge[l][0] == kHexFace[j][k]) + (kHexEdge[l][1] == kHexFace[j][k]);
               }
               if (hits == 2)
               {
                  for (int i = 0; i < pm1; i++) { mark[ed[8 + l * pm1 + i]] = 1; }
               }
            }
            for (int i = 0; i < pm1 * pm1; i++) { mark[ed[8 + 12 * pm1 + j * pm1 * pm1 + i]] = 1; }
         }
      }
      else
      {
         for (int j = 0; j < 4; j++)
         {
            if (edge_count[el_edge[(size_t)e * 4 + j]] != 1) { continue; }
            mark[v[kQuadEdge[j][0]]] = 1;
            mark[v[kQuadEdge[j][1]]] = 1;
            for (int i = 0; i < pm1; i++) { mark[ed[4 + j * pm1 + i]] = 1; }
         }
      }
   }
   S.bdr_dofs.clear();
   for (int i = 0; i < S.ndofs; i++)
   {
      if (mark[i]) { S.bdr_dofs.push_back(i); }
   }
}

} // namespace b2f

static int cartesian_h1_space_impl(int dim, int nx, int ny, int nz, int p, int sfc_ordering,
                                        int *ne, int *ndofs, int *elem_dof_h, int *dof_map_h,
                                        int *nbdr, int *bdr_dofs_h, double *elem_vertices_h)
{
   using namespace b2f;
   if ((dim != 2 && dim != 3) || nx < 1 || ny < 1 || (dim == 3 && nz < 1) || p < 1 || !ne || !ndofs)
   {
      set_error("b2f_cartesian_h1_space_h: bad arguments");
      return B2F_ERR_ARG;
   }
   if (dim == 2) { nz = 1; }
   const int nv = dim == 2 ? 4 : 8;
   const int64_t nel = (int64_t)nx * ny * nz;
   const int p1 = p + 1;
   const int64_t dof = dim == 2 ? p1 * p1 : p1 * p1 * p1;
   const int64_t nd_total = dim == 2 ? ((int64_t)nx * p + 1) * ((int64_t)ny * p + 1)
                            : ((int64_t)nx * p + 1) * ((int64_t)ny * p + 1) * ((int64_t)nz * p + 1);
   if (nel * dof > INT32_MAX || nd_total > INT32_MAX)
   {
      set_error("b2f_cartesian_h1_space_h: %lld E-dofs exceed the int32 index range of the "
                "restriction tables", (long long)(nel * dof));
      return B2F_ERR_ARG;
   }
   *ne = (int)nel;
   *ndofs = (int)nd_total;
   if (dof_map_h) { h1_dof_map(dim, p, dof_map_h); }
   if (!elem_dof_h && !bdr_dofs_h && !elem_vertices_h && !nbdr) { return B2F_OK; }

   // The caller asks in steps (sizes; element table, vertices and number of boundary dofs;
   // boundary dofs): the boundary list of the last numbering is kept so that the last step is a copy.
   struct BdrCache { bool valid = false; int key[6]; std::vector<int> bdr; };
   static thread_local BdrCache cache;
   const int key[6] = {dim, nx, ny, nz, p, sfc_ordering};
   if (!elem_dof_h && !elem_vertices_h && cache.valid && std::equal(key, key + 6, cache.key))
   {
      if (nbdr) { *nbdr = (int)cache.bdr.size(); }
      if (bdr_dofs_h) { std::copy(cache.bdr.begin(), cache.bdr.end(), bdr_dofs_h); }
      return B2F_OK;
   }
   const std::vector<V3> ord = grid_order(dim, nx, ny, nz, sfc_ordering != 0);
   const int nverts = dim == 2 ? (nx + 1) * (ny + 1) : (nx + 1) * (ny + 1) * (nz + 1);
   auto vtx = [&](int x, int y, int z) { return x + (y + z * (ny + 1)) * (nx + 1); };
   std::vector<int> ev((size_t)nel * nv);
   parallel_for(nel, [&](int64_t e_begin, int64_t e_end)
   {
      for (int64_t e = e_begin; e < e_end; e++)
      {
         const V3 c = ord[e];
         const int(*vc)[3] = dim == 2 ? kQuadVert : kHexVert;
         for (int k = 0; k < nv; k++) { ev[e * nv + k] = vtx(c.x + vc[k][0], c.y + vc[k][1], c.z + vc[k][2]); }
      }
   });
   if (elem_vertices_h)
   {
      // nodes E-vector of the order-1 mesh: (2^dim lexicographic, dim, NE); coordinates
      // computed as the reference does, ((real_t) x / nx) * sx with sx = 1
      parallel_for(nel, [&](int64_t e_begin, int64_t e_end)
      {
         for (int64_t e = e_begin; e < e_end; e++)
         {
            const V3 c = ord[e];
            for (int k = 0; k < nv; k++)
            {
               const int ix = c.x + (k & 1), iy = c.y + ((k >> 1) & 1), iz = c.z + ((k >> 2) & 1);
               double *o = elem_vertices_h + e * nv * dim;
               o[k + nv * 0] = ((double)ix / nx) * 1.0;
               o[k + nv * 1] = ((double)iy / ny) * 1.0;
               if (dim == 3) { o[k + nv * 2] = ((double)iz / nz) * 1.0; }
            }
         }
      });
   }
   if (!elem_dof_h && !bdr_dofs_h && !nbdr) { return B2F_OK; }
   if (!elem_dof_h && cache.valid && std::equal(key, key + 6, cache.key))
   {
      if (nbdr) { *nbdr = (int)cache.bdr.size(); }
      if (bdr_dofs_h) { std::copy(cache.bdr.begin(), cache.bdr.end(), bdr_dofs_h); }
      return B2F_OK;
   }
   H1Space S;
   const int vstride[3] = {1, nx + 1, (nx + 1) * (ny + 1)};
   h1_number(dim, p, nverts, ev, S, vstride);
   if (S.ndofs != *ndofs)
   {
      set_error("b2f_cartesian_h1_space_h: internal dof count mismatch (%d vs %d)", S.ndofs, *ndofs);
      return B2F_ERR_ARG;
   }
   if (elem_dof_h) { std::copy(S.elem_dof.begin(), S.elem_dof.end(), elem_dof_h); }
   if (nbdr) { *nbdr = (int)S.bdr_dofs.size(); }
   if (bdr_dofs_h) { std::copy(S.bdr_dofs.begin(), S.bdr_dofs.end(), bdr_dofs_h); }
   std::copy(key, key + 6, cache.key);
   cache.bdr.swap(S.bdr_dofs);
   cache.valid = true;
   return B2F_OK;
}

// nothing may throw across the C ABI (allocation failures of the multi-GB tables)
extern "C" int b2f_cartesian_h1_space_h(int dim, int nx, int ny, int nz, int p, int sfc_ordering,
                                        int *ne, int *ndofs, int *elem_dof_h, int *dof_map_h,
                                        int *nbdr, int *bdr_dofs_h, double *elem_vertices_h)
{
   try
   {
      return cartesian_h1_space_impl(dim, nx, ny, nz, p, sfc_ordering, ne, ndofs, elem_dof_h, dof_map_h, nbdr,
                                     bdr_dofs_h, elem_vertices_h);
   }
   catch (const std::exception &ex)
   {
      b2f::set_error("b2f_cartesian_h1_space_h: %s", ex.what());
      return B2F_ERR_ARG;
   }
}
