/* TEST INFRASTRUCTURE ONLY — minimal stand-in for the few deal.II types that three header-only
 * pieces of the reference use, so that those pieces can be compiled IN PLACE from
 * /root/reference and run here (oracle/ref_driver.cc, `make -C oracle ref`):
 *   source/nonlinear_elasticity/include/compressible_neo_hook_material.h   (tau, Jc, Psi)
 *   source/{non,}linear_elasticity/include/postprocessor.h                 (strain unrolling)
 *   include/adapter/time_handler.h                                         (Time)
 * Everything else of the reference needs the real deal.II (Triangulation, DoFHandler, FEValues,
 * AffineConstraints, SolverCG, ...) and stays unbuildable.
 *
 * Semantics restated from the deal.II 9.5 documentation, with FULL (uncompressed) index storage
 * on purpose - independent of the oracle's compressed SymmetricTensor code that it checks:
 *   SymmetricTensor<4> * SymmetricTensor<2> : (A:b)_ij   = sum_kl A_ijkl b_kl
 *   SymmetricTensor<4> * SymmetricTensor<4> : (A:B)_ijkl = sum_mn A_ijmn B_mnkl
 *   outer_product(a, b)_ijkl = a_ij b_kl ;  trace(a) = sum_i a_ii
 *   StandardTensors<dim>: I = delta_ij ; S = (delta_ik delta_jl + delta_il delta_jk)/2 ;
 *                         IxI = I (x) I ; dev_P = S - IxI/dim
 *   Tensor<2,dim>::component_to_unrolled_index((d,e)) = d*dim + e
 */
#ifndef DEALII_MIN_H
#define DEALII_MIN_H
#include <cassert>
#include <cmath>
#include <string>
#include <vector>

#define Assert(cond, exc) assert(cond)

namespace dealii
{
  struct ExcInternalError
  {};

  template <int rank, int dim, typename Number = double>
  class SymmetricTensor;

  template <int dim, typename Number>
  class SymmetricTensor<2, dim, Number>
  {
  public:
    Number v[dim][dim];
    SymmetricTensor()
    {
      for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j)
          v[i][j] = Number(0);
    }
  };
  template <int dim, typename Number>
  class SymmetricTensor<4, dim, Number>
  {
  public:
    Number v[dim][dim][dim][dim];
    SymmetricTensor()
    {
      for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j)
          for (int k = 0; k < dim; ++k)
            for (int l = 0; l < dim; ++l)
              v[i][j][k][l] = Number(0);
    }
  };

  template <int dim, typename N>
  SymmetricTensor<2, dim, N> operator+(const SymmetricTensor<2, dim, N> &a,
                                       const SymmetricTensor<2, dim, N> &b)
  {
    SymmetricTensor<2, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        r.v[i][j] = a.v[i][j] + b.v[i][j];
    return r;
  }
  template <int dim, typename N>
  SymmetricTensor<2, dim, N> operator*(const double s, const SymmetricTensor<2, dim, N> &a)
  {
    SymmetricTensor<2, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        r.v[i][j] = s * a.v[i][j];
    return r;
  }
  template <int dim, typename N>
  N trace(const SymmetricTensor<2, dim, N> &a)
  {
    N t = N(0);
    for (int i = 0; i < dim; ++i)
      t += a.v[i][i];
    return t;
  }
  template <int dim, typename N>
  SymmetricTensor<4, dim, N> operator+(const SymmetricTensor<4, dim, N> &a,
                                       const SymmetricTensor<4, dim, N> &b)
  {
    SymmetricTensor<4, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        for (int k = 0; k < dim; ++k)
          for (int l = 0; l < dim; ++l)
            r.v[i][j][k][l] = a.v[i][j][k][l] + b.v[i][j][k][l];
    return r;
  }
  template <int dim, typename N>
  SymmetricTensor<4, dim, N> operator-(const SymmetricTensor<4, dim, N> &a,
                                       const SymmetricTensor<4, dim, N> &b)
  {
    SymmetricTensor<4, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        for (int k = 0; k < dim; ++k)
          for (int l = 0; l < dim; ++l)
            r.v[i][j][k][l] = a.v[i][j][k][l] - b.v[i][j][k][l];
    return r;
  }
  template <int dim, typename N>
  SymmetricTensor<4, dim, N> operator*(const double s, const SymmetricTensor<4, dim, N> &a)
  {
    SymmetricTensor<4, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        for (int k = 0; k < dim; ++k)
          for (int l = 0; l < dim; ++l)
            r.v[i][j][k][l] = s * a.v[i][j][k][l];
    return r;
  }
  // double contraction 4:2 -> 2
  template <int dim, typename N>
  SymmetricTensor<2, dim, N> operator*(const SymmetricTensor<4, dim, N> &a,
                                       const SymmetricTensor<2, dim, N> &b)
  {
    SymmetricTensor<2, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        for (int k = 0; k < dim; ++k)
          for (int l = 0; l < dim; ++l)
            r.v[i][j] += a.v[i][j][k][l] * b.v[k][l];
    return r;
  }
  // double contraction 4:4 -> 4
  template <int dim, typename N>
  SymmetricTensor<4, dim, N> operator*(const SymmetricTensor<4, dim, N> &a,
                                       const SymmetricTensor<4, dim, N> &b)
  {
    SymmetricTensor<4, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        for (int k = 0; k < dim; ++k)
          for (int l = 0; l < dim; ++l)
            for (int m = 0; m < dim; ++m)
              for (int n = 0; n < dim; ++n)
                r.v[i][j][k][l] += a.v[i][j][m][n] * b.v[m][n][k][l];
    return r;
  }
  template <int dim, typename N>
  SymmetricTensor<4, dim, N> outer_product(const SymmetricTensor<2, dim, N> &a,
                                           const SymmetricTensor<2, dim, N> &b)
  {
    SymmetricTensor<4, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        for (int k = 0; k < dim; ++k)
          for (int l = 0; l < dim; ++l)
            r.v[i][j][k][l] = a.v[i][j] * b.v[k][l];
    return r;
  }

  namespace Physics
  {
    namespace Elasticity
    {
      template <int dim>
      struct StandardTensors
      {
        static const SymmetricTensor<2, dim> I;
        static const SymmetricTensor<4, dim> S, IxI, dev_P;
      };
      namespace internal
      {
        template <int dim>
        SymmetricTensor<2, dim> make_I()
        {
          SymmetricTensor<2, dim> r;
          for (int i = 0; i < dim; ++i)
            r.v[i][i] = 1.0;
          return r;
        }
        template <int dim>
        SymmetricTensor<4, dim> make_S()
        {
          SymmetricTensor<4, dim> r;
          for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j)
              for (int k = 0; k < dim; ++k)
                for (int l = 0; l < dim; ++l)
                  r.v[i][j][k][l] =
                    0.5 * ((i == k && j == l ? 1.0 : 0.0) + (i == l && j == k ? 1.0 : 0.0));
          return r;
        }
        template <int dim>
        SymmetricTensor<4, dim> make_dev_P()
        {
          return make_S<dim>() - (1.0 / dim) * outer_product(make_I<dim>(), make_I<dim>());
        }
      } // namespace internal
      template <int dim>
      const SymmetricTensor<2, dim> StandardTensors<dim>::I = internal::make_I<dim>();
      template <int dim>
      const SymmetricTensor<4, dim> StandardTensors<dim>::S = internal::make_S<dim>();
      template <int dim>
      const SymmetricTensor<4, dim> StandardTensors<dim>::IxI =
        outer_product(internal::make_I<dim>(), internal::make_I<dim>());
      template <int dim>
      const SymmetricTensor<4, dim> StandardTensors<dim>::dev_P = internal::make_dev_P<dim>();
    } // namespace Elasticity
  }   // namespace Physics

  // ---- what postprocessor.h touches --------------------------------------------------------
  template <typename Number>
  class Vector
  {
  public:
    std::vector<Number> data;
    Vector() = default;
    explicit Vector(unsigned n)
      : data(n, Number(0))
    {}
    unsigned      size() const { return unsigned(data.size()); }
    Number &      operator[](unsigned i) { return data[i]; }
    const Number &operator[](unsigned i) const { return data[i]; }
    Number &      operator()(unsigned i) { return data[i]; }
    const Number &operator()(unsigned i) const { return data[i]; }
  };
  template <int rank>
  struct TableIndices;
  template <>
  struct TableIndices<2>
  {
    unsigned i[2];
    TableIndices(unsigned a, unsigned b)
      : i{a, b}
    {}
  };
  template <int rank, int dim, typename Number = double>
  class Tensor;
  template <int dim, typename Number>
  class Tensor<1, dim, Number>
  {
  public:
    Number        v[dim] = {};
    Number &      operator[](unsigned i) { return v[i]; }
    const Number &operator[](unsigned i) const { return v[i]; }
  };
  template <int dim, typename Number>
  class Tensor<2, dim, Number>
  {
  public:
    static unsigned component_to_unrolled_index(const TableIndices<2> &t)
    {
      return t.i[0] * dim + t.i[1]; // row-major unrolling
    }
  };
  namespace DataPostprocessorInputs
  {
    template <int dim>
    struct Vector
    {
      std::vector<dealii::Vector<double>>         solution_values;
      std::vector<std::vector<Tensor<1, dim>>>    solution_gradients;
    };
  } // namespace DataPostprocessorInputs
  namespace DataComponentInterpretation
  {
    enum DataComponentInterpretation
    {
      component_is_scalar,
      component_is_part_of_vector
    };
  }
  enum UpdateFlags
  {
    update_default   = 0,
    update_values    = 1,
    update_gradients = 2
  };
  inline UpdateFlags operator|(UpdateFlags a, UpdateFlags b) { return UpdateFlags(int(a) | int(b)); }
  template <int dim>
  class DataPostprocessor
  {
  public:
    virtual ~DataPostprocessor() = default;
    virtual void evaluate_vector_field(const DataPostprocessorInputs::Vector<dim> &,
                                       std::vector<Vector<double>> &) const = 0;
    virtual std::vector<std::string> get_names() const = 0;
    virtual std::vector<DataComponentInterpretation::DataComponentInterpretation>
                        get_data_component_interpretation() const = 0;
    virtual UpdateFlags get_needed_update_flags() const = 0;
  };
} // namespace dealii
#endif
