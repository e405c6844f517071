/* TEST INFRASTRUCTURE ONLY — what the reference's cell-assembly block
 * (nonlinear_elasticity.cc:626-1036: Assembler_Base with PerTaskData_ASM / ScratchData_ASM /
 * assemble_neumann_contribution_one_cell, Assembler<dim,double>::
 * assemble_system_tangent_residual_one_cell) and PointHistory (nonlinear_elasticity.h:64-117) need
 * around them to compile and run OUTSIDE deal.II. The two line ranges are cut out of the
 * reference at build time into oracle/_ref/ as .inc files (never committed, see oracle/Makefile) because
 * the structs live inside a translation unit whose other 1,000 lines need the real library.
 *
 * FEValues / FEFaceValues here do not compute anything: they hand out the shape-function tables,
 * JxW and normals the driver was given (so the finite-element tables remain a restatement of
 * deal.II - they come from tests/ref_formulas.py - while every line of arithmetic on them is the
 * reference's own). `Solid` is a plain struct with the members the block reads through
 * `data.solid->`. */
#ifndef ASSEMBLY_SHIM_H
#define ASSEMBLY_SHIM_H
#include <deal.II/base/dealii_min.h>

#include <memory>
#include <string>
#include <stdexcept>
#include <utility>

#define AssertThrow(cond, exc)         \
  do                                   \
    {                                  \
      if (!(cond))                     \
        throw std::runtime_error(#cond); \
    }                                  \
  while (0)

namespace dealii
{
  struct ExcPureFunctionCalled
  {};
  inline const char *ExcMessage(const char *m) { return m; }
  namespace types
  {
    using global_dof_index = unsigned int;
  }
  template <typename N>
  class FullMatrix
  {
  public:
    unsigned       m_, n_;
    std::vector<N> a;
    FullMatrix(unsigned m, unsigned n)
      : m_(m)
      , n_(n)
      , a(size_t(m) * n, N(0))
    {}
    FullMatrix &operator=(const N s)
    {
      for (auto &x : a)
        x = s;
      return *this;
    }
    N &      operator()(unsigned i, unsigned j) { return a[size_t(i) * n_ + j]; }
    const N &operator()(unsigned i, unsigned j) const { return a[size_t(i) * n_ + j]; }
  };
  template <typename N>
  class BlockVector : public Vector<N>
  {
  public:
    using Vector<N>::Vector;
    BlockVector() = default;
    // BlockVector(block_sizes): the reference only ever has one block
    explicit BlockVector(const std::vector<types::global_dof_index> &block_sizes)
      : Vector<N>(block_sizes.at(0))
    {}
    Vector<N> &      block(unsigned) { return *this; }
    const Vector<N> &block(unsigned) const { return *this; }
  };
  // dense stand-in for SparseMatrix: vmult only
  template <typename N>
  class SparseMatrix
  {
  public:
    unsigned       n = 0;
    std::vector<N> a;
    void           vmult(Vector<N> &dst, const Vector<N> &src) const
    {
      for (unsigned i = 0; i < n; ++i)
        {
          N s = N(0);
          for (unsigned j = 0; j < n; ++j)
            s += a[size_t(i) * n + j] * src[j];
          dst[i] = s;
        }
    }
  };
  template <typename N>
  class BlockSparseMatrix
  {};
  template <typename N>
  class AffineConstraints
  {
  public:
    std::vector<unsigned char> constrained;
    bool is_constrained(unsigned i) const { return i < constrained.size() && constrained[i] != 0; }
    template <class M, class V, class I, class GM, class GV>
    void distribute_local_to_global(const M &, const V &, const I &, GM &, GV &) const
    {
      throw std::runtime_error("not part of the stand-in");
    }
  };
  // system_to_component_index of local dof i: (node, component); empty tables = node-major
  struct ShimNumbering
  {
    std::vector<unsigned> node_of, comp_of;
    static ShimNumbering &get()
    {
      static ShimNumbering n;
      return n;
    }
  };
  template <int dim>
  inline unsigned shim_node(unsigned i)
  {
    const auto &n = ShimNumbering::get();
    return n.node_of.empty() ? i / dim : n.node_of[i];
  }
  template <int dim>
  inline unsigned shim_component(unsigned i)
  {
    const auto &n = ShimNumbering::get();
    return n.comp_of.empty() ? i % dim : n.comp_of[i];
  }
  // optional tail of a driver's input: node_of[dpc] comp_of[dpc]
  inline void read_shim_numbering(std::istream &in, unsigned dpc)
  {
    auto &              n = ShimNumbering::get();
    std::vector<unsigned> a(dpc), b(dpc);
    for (unsigned i = 0; i < dpc; ++i)
      if (!(in >> a[i]))
        return;
    for (unsigned i = 0; i < dpc; ++i)
      if (!(in >> b[i]))
        return;
    n.node_of = a;
    n.comp_of = b;
  }
  template <int dim>
  class FiniteElement
  {
  public:
    unsigned dofs_per_cell = 0;
    // FESystem(FE_Q(p), dim): one base element; dof i = node * dim + component up to degree 2,
    // the table of the driver's input (entity-major numbering) beyond
    std::pair<std::pair<unsigned, unsigned>, unsigned> system_to_base_index(unsigned i) const
    {
      return {{0u, shim_component<dim>(i)}, shim_node<dim>(i)};
    }
    std::pair<unsigned, unsigned> system_to_component_index(unsigned i) const
    {
      return {shim_component<dim>(i), shim_node<dim>(i)};
    }
  };
  template <int dim>
  class FESystem : public FiniteElement<dim>
  {};
  template <int dim>
  class QGauss
  {
  public:
    unsigned n = 0;
    unsigned size() const { return n; }
  };
  namespace FEValuesExtractors
  {
    struct Vector
    {
      unsigned first_vector_component = 0;
    };
  } // namespace FEValuesExtractors

  template <int dim>
  class DoFHandler
  {
  public:
    struct Face
    {
      bool     boundary = false;
      unsigned id = 0, number = 0;
      bool     at_boundary() const { return boundary; }
      unsigned boundary_id() const { return id; }
    };
    struct Cell
    {
      std::vector<types::global_dof_index> dofs;
      std::vector<Face>                    faces; // in deal.II face order 0..2*dim-1
      void get_dof_indices(std::vector<types::global_dof_index> &v) const { v = dofs; }
      std::vector<const Face *> face_iterators() const
      {
        std::vector<const Face *> r;
        for (const auto &f : faces)
          r.push_back(&f);
        return r;
      }
    };
    using active_cell_iterator = const Cell *;
  };

  // tables the driver provides for THE cell being assembled (real-space gradients of the scalar
  // shape functions, JxW) and for each of its faces (values, JxW, unit normals)
  template <int dim>
  struct ShimTables
  {
    unsigned            nq = 0, nqf = 0, npc = 0;
    std::vector<double> N, gradN, JxW;                 // [nq][npc], [nq][npc][dim], [nq]
    std::vector<std::vector<double>> Nf, JxWf, normal; // per face: [nqf][npc], [nqf], [nqf][dim]
    static ShimTables &get()
    {
      static ShimTables t;
      return t;
    }
  };

  template <int dim, bool face_values>
  class FEValuesShim
  {
  public:
    const FiniteElement<dim> &fe;
    unsigned                  n_q;
    UpdateFlags               flags;
    unsigned                  dofs_per_cell;
    unsigned                  n_quadrature_points;
    QGauss<dim>               cell_q;
    QGauss<dim - 1>           face_q;
    const typename DoFHandler<dim>::Cell *cell = nullptr;
    unsigned                              face = 0;

    template <class Q>
    FEValuesShim(const FiniteElement<dim> &fe, const Q &q, const UpdateFlags f)
      : fe(fe)
      , n_q(q.size())
      , flags(f)
      , dofs_per_cell(fe.dofs_per_cell)
      , n_quadrature_points(q.size())
    {
      cell_q.n = q.size();
      face_q.n = q.size();
    }
    const FiniteElement<dim> &get_fe() const { return fe; }
    const auto &              get_quadrature() const
    {
      if constexpr (face_values)
        return face_q;
      else
        return cell_q;
    }
    UpdateFlags get_update_flags() const { return flags; }
    void        reinit(const typename DoFHandler<dim>::active_cell_iterator &c) { cell = c; }
    void        reinit(const typename DoFHandler<dim>::active_cell_iterator &c,
                       const typename DoFHandler<dim>::Face *                 f)
    {
      cell = c;
      face = f->number;
    }
    double shape_value(unsigned i, unsigned q) const
    {
      const auto &t = ShimTables<dim>::get();
      return face_values ? t.Nf[face][q * t.npc + shim_node<dim>(i)] : t.N[q * t.npc + shim_node<dim>(i)];
    }
    // gradient of the only non-zero component of shape function i (FEValues::shape_grad)
    Tensor<1, dim> shape_grad(unsigned i, unsigned q) const
    {
      const auto &   t = ShimTables<dim>::get();
      Tensor<1, dim> g;
      for (int d = 0; d < dim; ++d)
        g[d] = t.gradN[(q * t.npc + shim_node<dim>(i)) * dim + d];
      return g;
    }
    // FEValuesBase::get_function_values for a vector-valued element
    template <class V>
    void get_function_values(const V &global, std::vector<Vector<double>> &out) const
    {
      for (unsigned q = 0; q < out.size(); ++q)
        {
          for (int d = 0; d < dim; ++d)
            out[q][d] = 0.0;
          for (unsigned k = 0; k < dofs_per_cell; ++k)
            out[q][shim_component<dim>(k)] += global(cell->dofs[k]) * shape_value(k, q);
        }
    }
    double JxW(unsigned q) const
    {
      const auto &t = ShimTables<dim>::get();
      return face_values ? t.JxWf[face][q] : t.JxW[q];
    }
    Tensor<1, dim> normal_vector(unsigned q) const
    {
      const auto &   t = ShimTables<dim>::get();
      Tensor<1, dim> n;
      for (int d = 0; d < dim; ++d)
        n[d] = t.normal[face][q * dim + d];
      return n;
    }
    struct View
    {
      const FEValuesShim &fv;
      // value / gradient of vector-valued shape function k: only its own component is non-zero
      Tensor<1, dim> value(unsigned k, unsigned q) const
      {
        Tensor<1, dim> r;
        r[shim_component<dim>(k)] = fv.shape_value(k, q);
        return r;
      }
      Tensor<2, dim> gradient(unsigned k, unsigned q) const
      {
        const auto &   t = ShimTables<dim>::get();
        Tensor<2, dim> r;
        for (int d = 0; d < dim; ++d)
          r[shim_component<dim>(k)][d] = t.gradN[(q * t.npc + shim_node<dim>(k)) * dim + d];
        return r;
      }
      template <class V>
      void get_function_values(const V &global, std::vector<Tensor<1, dim>> &out) const
      {
        for (unsigned q = 0; q < out.size(); ++q)
          {
            out[q] = Tensor<1, dim>();
            for (unsigned k = 0; k < fv.dofs_per_cell; ++k)
              out[q][shim_component<dim>(k)] += global(fv.cell->dofs[k]) * fv.shape_value(k, q);
          }
      }
      template <class V>
      void get_function_gradients(const V &global, std::vector<Tensor<2, dim>> &out) const
      {
        for (unsigned q = 0; q < out.size(); ++q)
          {
            out[q] = Tensor<2, dim>();
            for (unsigned k = 0; k < fv.dofs_per_cell; ++k)
              {
                const Tensor<2, dim> g = gradient(k, q);
                for (int d = 0; d < dim; ++d)
                  out[q][shim_component<dim>(k)][d] += global(fv.cell->dofs[k]) * g[shim_component<dim>(k)][d];
              }
          }
      }
    };
    View operator[](const FEValuesExtractors::Vector &) const { return View{*this}; }
  };
  template <int dim>
  using FEValues = FEValuesShim<dim, false>;
  template <int dim>
  using FEFaceValues = FEValuesShim<dim, true>;
} // namespace dealii

namespace Parameters
{
  struct AllParameters
  {
    double mu = 0, nu = 0, rho = 0;
    double beta = 0, gamma = 0, theta = 0, delta_t = 0;
    bool   data_consistent = true;
    int          output_interval = 1;
    std::string  scenario = "FSI3";
    double       flap_location = 0.0;
    unsigned int max_iterations_NR = 10;
    double       tol_f = 1e-9, tol_u = 1e-6;
  };
} // namespace Parameters
#endif
