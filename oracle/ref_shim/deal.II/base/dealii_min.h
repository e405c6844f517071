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
 *   Tensor<1>*Tensor<1> scalar product; Tensor<1>*Tensor<2> / Tensor<2>*Tensor<1> / Tensor<2>*
 *   Tensor<2> contract the adjacent indices; symmetrize(t) = (t + t^T)/2;
 *   Kinematics::F(Grad_u) = I + Grad_u ; F_iso(F) = det(F)^(-1/dim) F ; b(F) = symmetrize(F F^T)
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
    Vector &      operator=(const Number s)
    {
      for (auto &x : data)
        x = s;
      return *this;
    }
    Number &      operator[](unsigned i) { return data[i]; }
    const Number &operator[](unsigned i) const { return data[i]; }
    Number &      operator()(unsigned i) { return data[i]; }
    const Number &operator()(unsigned i) const { return data[i]; }
    // the BLAS-1 members the time-stepping code uses (Vector::equ / add / sadd-free subset)
    void reinit(unsigned n) { data.assign(n, Number(0)); }
    void equ(const Number a, const Vector &v)
    {
      data.resize(v.data.size());
      for (size_t i = 0; i < data.size(); ++i)
        data[i] = a * v.data[i];
    }
    void add(const Number a, const Vector &v)
    {
      for (size_t i = 0; i < data.size(); ++i)
        data[i] += a * v.data[i];
    }
    void add(const Number a, const Vector &v, const Number b, const Vector &w)
    {
      for (size_t i = 0; i < data.size(); ++i)
        data[i] += a * v.data[i] + b * w.data[i];
    }
    Vector &operator+=(const Vector &v)
    {
      for (size_t i = 0; i < data.size(); ++i)
        data[i] += v.data[i];
      return *this;
    }
    Vector &operator*=(const Number a)
    {
      for (auto &x : data)
        x *= a;
      return *this;
    }
    Number l2_norm() const
    {
      Number s = Number(0);
      for (const auto &x : data)
        s += x * x;
      return std::sqrt(s);
    }
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
    Number        norm() const
    {
      Number s = Number(0);
      for (int i = 0; i < dim; ++i)
        s += v[i] * v[i];
      return std::sqrt(s);
    }
  };
  template <int dim, typename Number>
  class Tensor<2, dim, Number>
  {
  public:
    Tensor<1, dim, Number> r[dim]; // rows
    Tensor() = default;
    Tensor(const SymmetricTensor<2, dim, Number> &s)
    {
      for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j)
          r[i][j] = s.v[i][j];
    }
    Tensor<1, dim, Number> &      operator[](unsigned i) { return r[i]; }
    const Tensor<1, dim, Number> &operator[](unsigned i) const { return r[i]; }
    static unsigned component_to_unrolled_index(const TableIndices<2> &t)
    {
      return t.i[0] * dim + t.i[1]; // row-major unrolling
    }
  };
  // ---- Tensor algebra used by the cell assembly (nonlinear_elasticity.cc:791-1036) ----------
  template <int dim, typename N>
  Tensor<1, dim, N> operator*(const Tensor<1, dim, N> &a, const double s)
  {
    Tensor<1, dim, N> r;
    for (int i = 0; i < dim; ++i)
      r[i] = a[i] * s;
    return r;
  }
  template <int dim, typename N>
  N operator*(const Tensor<1, dim, N> &a, const Tensor<1, dim, N> &b) // scalar product
  {
    N s = N(0);
    for (int i = 0; i < dim; ++i)
      s += a[i] * b[i];
    return s;
  }
  template <int dim, typename N>
  Tensor<2, dim, N> operator*(const Tensor<2, dim, N> &a, const Tensor<2, dim, N> &b)
  {
    Tensor<2, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        for (int k = 0; k < dim; ++k)
          r[i][j] += a[i][k] * b[k][j];
    return r;
  }
  template <int dim, typename N>
  Tensor<1, dim, N> operator*(const Tensor<2, dim, N> &a, const Tensor<1, dim, N> &b)
  {
    Tensor<1, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int k = 0; k < dim; ++k)
        r[i] += a[i][k] * b[k];
    return r;
  }
  template <int dim, typename N>
  Tensor<1, dim, N> operator*(const Tensor<1, dim, N> &a, const Tensor<2, dim, N> &b)
  {
    Tensor<1, dim, N> r; // contraction over the last index of a and the first of b
    for (int j = 0; j < dim; ++j)
      for (int k = 0; k < dim; ++k)
        r[j] += a[k] * b[k][j];
    return r;
  }
  template <int dim, typename N>
  Tensor<2, dim, N> operator*(const double s, const Tensor<2, dim, N> &a)
  {
    Tensor<2, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        r[i][j] = s * a[i][j];
    return r;
  }
  template <typename N>
  N determinant(const Tensor<2, 2, N> &t)
  {
    return t[0][0] * t[1][1] - t[1][0] * t[0][1];
  }
  template <typename N>
  N determinant(const Tensor<2, 3, N> &t)
  {
    return t[0][0] * (t[1][1] * t[2][2] - t[1][2] * t[2][1]) -
           t[0][1] * (t[1][0] * t[2][2] - t[1][2] * t[2][0]) +
           t[0][2] * (t[1][0] * t[2][1] - t[1][1] * t[2][0]);
  }
  template <typename N>
  Tensor<2, 2, N> invert(const Tensor<2, 2, N> &t)
  {
    const N         id = N(1) / determinant(t);
    Tensor<2, 2, N> r;
    r[0][0] = t[1][1] * id;
    r[0][1] = -t[0][1] * id;
    r[1][0] = -t[1][0] * id;
    r[1][1] = t[0][0] * id;
    return r;
  }
  template <typename N>
  Tensor<2, 3, N> invert(const Tensor<2, 3, N> &t)
  {
    const N         id = N(1) / determinant(t);
    Tensor<2, 3, N> r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        {
          // cofactor of t[j][i]
          const int a = (j + 1) % 3, b = (j + 2) % 3, c = (i + 1) % 3, d = (i + 2) % 3;
          r[i][j]     = (t[a][c] * t[b][d] - t[a][d] * t[b][c]) * id;
        }
    return r;
  }
  template <int dim, typename N>
  Tensor<2, dim, N> transpose(const Tensor<2, dim, N> &t)
  {
    Tensor<2, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        r[i][j] = t[j][i];
    return r;
  }
  template <int dim, typename N>
  SymmetricTensor<2, dim, N> symmetrize(const Tensor<2, dim, N> &t)
  {
    SymmetricTensor<2, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        r.v[i][j] = (t[i][j] + t[j][i]) / 2;
    return r;
  }
  // double contractions 2:2 -> scalar and 2:4 -> 2
  template <int dim, typename N>
  N operator*(const SymmetricTensor<2, dim, N> &a, const SymmetricTensor<2, dim, N> &b)
  {
    N s = N(0);
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        s += a.v[i][j] * b.v[i][j];
    return s;
  }
  template <int dim, typename N>
  SymmetricTensor<2, dim, N> operator*(const SymmetricTensor<2, dim, N> &a,
                                       const SymmetricTensor<4, dim, N> &c)
  {
    SymmetricTensor<2, dim, N> r;
    for (int i = 0; i < dim; ++i)
      for (int j = 0; j < dim; ++j)
        for (int k = 0; k < dim; ++k)
          for (int l = 0; l < dim; ++l)
            r.v[k][l] += a.v[i][j] * c.v[i][j][k][l];
    return r;
  }
  namespace Physics
  {
    namespace Elasticity
    {
      namespace Kinematics
      {
        template <int dim, typename N>
        Tensor<2, dim, N> F(const Tensor<2, dim, N> &Grad_u)
        {
          Tensor<2, dim, N> r = Grad_u;
          for (int i = 0; i < dim; ++i)
            r[i][i] += N(1);
          return r;
        }
        template <int dim, typename N>
        Tensor<2, dim, N> F_iso(const Tensor<2, dim, N> &F)
        {
          return std::pow(determinant(F), -1.0 / dim) * F;
        }
        template <int dim, typename N>
        SymmetricTensor<2, dim, N> b(const Tensor<2, dim, N> &F)
        {
          return symmetrize(F * transpose(F));
        }
      } // namespace Kinematics
    }   // namespace Elasticity
  }     // namespace Physics
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
