// Minimal, definition-complete stand-in for the parts of Eigen 3 that the reference's physics
// translation units use (fixed-size dense algebra, AlignedBox, Hyperplane, Affine3d, JacobiSVD).
// TEST INFRASTRUCTURE: lets oracle/build_ref.sh compile the reference's OWN sources (unmodified,
// where they lie under /root/reference) although Eigen is absent from this image.  Written from
// the documented Eigen API, not from Eigen's source.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <optional>
#include <vector>

namespace Eigen {

constexpr int Dynamic = -1;
enum DecompositionOptions { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };

template <typename T, int R, int C>
class Matrix;

template <typename T, int R, int C>
class ArrayWrapper
{
  public:
    std::array<T, static_cast<std::size_t>(R * C)> a{};
    ArrayWrapper operator*(ArrayWrapper const& o) const
    {
        ArrayWrapper r;
        for (int i = 0; i < R * C; ++i)
            r.a[i] = a[i] * o.a[i];
        return r;
    }
    T sum() const
    {
        T s = T(0);
        for (int i = 0; i < R * C; ++i)
            s += a[i];
        return s;
    }
};

// column view used as an l-value: M.col(i) = v
template <typename T, int R, int C>
class ColRef
{
  public:
    ColRef(Matrix<T, R, C>& m, int c) : m_(m), c_(c) {}
    ColRef& operator=(Matrix<T, R, 1> const& v)
    {
        for (int r = 0; r < R; ++r)
            m_(r, c_) = v(r, 0);
        return *this;
    }
    template <int RR = R, typename = std::enable_if_t<RR != 1>>
    ColRef& operator=(Matrix<T, 1, R> const& v) // vector transposition on assignment
    {
        for (int r = 0; r < R; ++r)
            m_(r, c_) = v(0, r);
        return *this;
    }
    ColRef& operator=(ColRef const& o) { return *this = static_cast<Matrix<T, R, 1>>(o); }
    operator Matrix<T, R, 1>() const
    {
        Matrix<T, R, 1> v;
        for (int r = 0; r < R; ++r)
            v(r, 0) = m_(r, c_);
        return v;
    }
    Matrix<T, R, 1> operator-() const { return -static_cast<Matrix<T, R, 1>>(*this); }

  private:
    Matrix<T, R, C>& m_;
    int c_;
};

template <typename T, int R, int C>
class Matrix
{
  public:
    using Scalar = T;
    Matrix() { d_.fill(T(0)); } // Eigen leaves this uninitialised; zero is a valid instance of that
    template <int N = R * C, typename = std::enable_if_t<N == 3>>
    Matrix(T x, T y, T z) : d_{x, y, z}
    {
    }
    template <int N = R * C, typename = std::enable_if_t<N == 2>>
    Matrix(T x, T y) : d_{x, y}
    {
    }
    template <int N = R * C, typename = std::enable_if_t<N == 4>>
    Matrix(T x, T y, T z, T w) : d_{x, y, z, w}
    {
    }

    static Matrix Zero() { return Matrix(); }
    static Matrix Ones()
    {
        Matrix m;
        m.d_.fill(T(1));
        return m;
    }
    static Matrix Identity()
    {
        Matrix m;
        for (int i = 0; i < (R < C ? R : C); ++i)
            m(i, i) = T(1);
        return m;
    }
    void setZero() { d_.fill(T(0)); }

    // column-major storage, like Eigen's default
    T& operator()(int r, int c) { return d_[static_cast<std::size_t>(c * R + r)]; }
    T const& operator()(int r, int c) const { return d_[static_cast<std::size_t>(c * R + r)]; }
    T& operator()(int i) { return d_[static_cast<std::size_t>(i)]; }
    T const& operator()(int i) const { return d_[static_cast<std::size_t>(i)]; }
    T& operator[](int i) { return d_[static_cast<std::size_t>(i)]; }
    T const& operator[](int i) const { return d_[static_cast<std::size_t>(i)]; }
    T& x() { return d_[0]; }
    T& y() { return d_[1]; }
    T& z() { return d_[2]; }
    T const& x() const { return d_[0]; }
    T const& y() const { return d_[1]; }
    T const& z() const { return d_[2]; }
    T* data() { return d_.data(); }
    T const* data() const { return d_.data(); }

    ColRef<T, R, C> col(int c) { return ColRef<T, R, C>(*this, c); }
    Matrix<T, R, 1> col(int c) const
    {
        Matrix<T, R, 1> v;
        for (int r = 0; r < R; ++r)
            v(r, 0) = (*this)(r, c);
        return v;
    }

    Matrix operator+(Matrix const& o) const
    {
        Matrix m;
        for (int i = 0; i < R * C; ++i)
            m.d_[i] = d_[i] + o.d_[i];
        return m;
    }
    Matrix operator-(Matrix const& o) const
    {
        Matrix m;
        for (int i = 0; i < R * C; ++i)
            m.d_[i] = d_[i] - o.d_[i];
        return m;
    }
    Matrix operator-() const
    {
        Matrix m;
        for (int i = 0; i < R * C; ++i)
            m.d_[i] = -d_[i];
        return m;
    }
    Matrix operator*(T s) const
    {
        Matrix m;
        for (int i = 0; i < R * C; ++i)
            m.d_[i] = d_[i] * s;
        return m;
    }
    Matrix operator/(T s) const
    {
        Matrix m;
        for (int i = 0; i < R * C; ++i)
            m.d_[i] = d_[i] / s;
        return m;
    }
    Matrix& operator+=(Matrix const& o)
    {
        for (int i = 0; i < R * C; ++i)
            d_[i] += o.d_[i];
        return *this;
    }
    Matrix& operator-=(Matrix const& o)
    {
        for (int i = 0; i < R * C; ++i)
            d_[i] -= o.d_[i];
        return *this;
    }
    Matrix& operator*=(T s)
    {
        for (int i = 0; i < R * C; ++i)
            d_[i] *= s;
        return *this;
    }
    Matrix& operator/=(T s)
    {
        for (int i = 0; i < R * C; ++i)
            d_[i] /= s;
        return *this;
    }
    template <int K>
    Matrix<T, R, K> operator*(Matrix<T, C, K> const& o) const
    {
        Matrix<T, R, K> m;
        for (int r = 0; r < R; ++r)
            for (int k = 0; k < K; ++k)
            {
                T s = T(0);
                for (int c = 0; c < C; ++c)
                    s += (*this)(r, c) * o(c, k);
                m(r, k) = s;
            }
        return m;
    }
    Matrix<T, C, R> transpose() const
    {
        Matrix<T, C, R> m;
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C; ++c)
                m(c, r) = (*this)(r, c);
        return m;
    }
    T trace() const
    {
        T s = T(0);
        for (int i = 0; i < (R < C ? R : C); ++i)
            s += (*this)(i, i);
        return s;
    }
    T squaredNorm() const
    {
        T s = T(0);
        for (int i = 0; i < R * C; ++i)
            s += d_[i] * d_[i];
        return s;
    }
    T norm() const { return std::sqrt(squaredNorm()); }
    bool isZero(T prec = std::numeric_limits<T>::epsilon() * T(100)) const
    {
        for (int i = 0; i < R * C; ++i)
            if (std::abs(d_[i]) > prec)
                return false;
        return true;
    }
    bool isApprox(Matrix const& o, T prec = std::numeric_limits<T>::epsilon() * T(100)) const
    {
        return (*this - o).squaredNorm() <= prec * prec * std::min(squaredNorm(), o.squaredNorm());
    }
    Matrix normalized() const
    {
        T const n = norm();
        return n > T(0) ? (*this) / n : *this;
    }
    void normalize()
    {
        T const n = norm();
        if (n > T(0))
            *this /= n;
    }
    T dot(Matrix const& o) const
    {
        T s = T(0);
        for (int i = 0; i < R * C; ++i)
            s += d_[i] * o.d_[i];
        return s;
    }
    template <int N = R * C, typename = std::enable_if_t<N == 3>>
    Matrix cross(Matrix const& o) const
    {
        return Matrix(d_[1] * o.d_[2] - d_[2] * o.d_[1], d_[2] * o.d_[0] - d_[0] * o.d_[2],
                      d_[0] * o.d_[1] - d_[1] * o.d_[0]);
    }
    ArrayWrapper<T, R, C> array() const
    {
        ArrayWrapper<T, R, C> a;
        a.a = d_;
        return a;
    }
    Matrix<T, R + 1, 1> homogeneous() const
    {
        static_assert(C == 1, "homogeneous() of a column vector");
        Matrix<T, R + 1, 1> h;
        for (int r = 0; r < R; ++r)
            h(r, 0) = d_[r];
        h(R, 0) = T(1);
        return h;
    }
    template <typename U>
    Matrix<U, R, C> cast() const
    {
        Matrix<U, R, C> m;
        for (int i = 0; i < R * C; ++i)
            m.data()[i] = static_cast<U>(d_[i]);
        return m;
    }
    template <int RR = R, int CC = C, typename = std::enable_if_t<RR == 3 && CC == 3>>
    T determinant() const
    {
        auto const& m = *this;
        return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
               m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
    }
    template <int RR = R, int CC = C, typename = std::enable_if_t<RR == 3 && CC == 3>>
    Matrix inverse() const
    { // cofactor / determinant, the closed form Eigen uses for 3x3
        auto const& m = *this;
        T const id    = T(1) / determinant();
        Matrix r;
        r(0, 0) = (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) * id;
        r(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) * id;
        r(0, 2) = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) * id;
        r(1, 0) = (m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2)) * id;
        r(1, 1) = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) * id;
        r(1, 2) = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) * id;
        r(2, 0) = (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0)) * id;
        r(2, 1) = (m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1)) * id;
        r(2, 2) = (m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0)) * id;
        return r;
    }

  private:
    std::array<T, static_cast<std::size_t>(R * C)> d_;
};

template <typename T, int R, int C>
Matrix<T, R, C> operator*(T s, Matrix<T, R, C> const& m)
{
    return m * s;
}
template <typename T, int R, int C, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value && !std::is_same<S, T>::value>>
Matrix<T, R, C> operator*(S s, Matrix<T, R, C> const& m)
{
    return m * static_cast<T>(s);
}
template <typename T, int R, int C, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value && !std::is_same<S, T>::value>>
Matrix<T, R, C> operator*(Matrix<T, R, C> const& m, S s)
{
    return m * static_cast<T>(s);
}

using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector4d = Matrix<double, 4, 1>;
using Vector3f = Matrix<float, 3, 1>;
using Vector2f = Matrix<float, 2, 1>;
using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;
using RowVector3d = Matrix<double, 1, 3>;

template <typename T, int N>
class AlignedBox
{
  public:
    using V = Matrix<T, N, 1>;
    AlignedBox()
    {
        for (int i = 0; i < N; ++i)
        {
            min_(i) = std::numeric_limits<T>::max();
            max_(i) = std::numeric_limits<T>::lowest();
        }
    }
    AlignedBox(V const& mn, V const& mx) : min_(mn), max_(mx) {}
    V const& min() const { return min_; }
    V const& max() const { return max_; }
    V& min() { return min_; }
    V& max() { return max_; }
    AlignedBox& extend(V const& p)
    {
        for (int i = 0; i < N; ++i)
        {
            min_(i) = std::min(min_(i), p(i));
            max_(i) = std::max(max_(i), p(i));
        }
        return *this;
    }
    V diagonal() const { return max_ - min_; }
    V center() const { return (min_ + max_) / T(2); }
    bool contains(V const& p) const
    {
        for (int i = 0; i < N; ++i)
            if (p(i) < min_(i) || p(i) > max_(i))
                return false;
        return true;
    }

  private:
    V min_, max_;
};
using AlignedBox3d = AlignedBox<double, 3>;

template <typename T, int N>
class Hyperplane
{
  public:
    using V = Matrix<T, N, 1>;
    Hyperplane() = default;
    // plane through point e with (unit) normal n: offset = -n.e
    Hyperplane(V const& n, V const& e) : n_(n), offset_(-n.dot(e)) {}
    Hyperplane(V const& n, T offset) : n_(n), offset_(offset) {}
    T signedDistance(V const& p) const { return n_.dot(p) + offset_; }
    V const& normal() const { return n_; }
    T offset() const { return offset_; }

  private:
    V n_;
    T offset_ = T(0);
};

class AngleAxisd
{
  public:
    AngleAxisd(double angle, Vector3d const& axis) : angle_(angle), axis_(axis) {}
    Matrix3d toRotationMatrix() const
    {
        double const c = std::cos(angle_), s = std::sin(angle_), t = 1.0 - c;
        double const x = axis_.x(), y = axis_.y(), z = axis_.z();
        Matrix3d m;
        m(0, 0) = t * x * x + c;
        m(0, 1) = t * x * y - s * z;
        m(0, 2) = t * x * z + s * y;
        m(1, 0) = t * x * y + s * z;
        m(1, 1) = t * y * y + c;
        m(1, 2) = t * y * z - s * x;
        m(2, 0) = t * x * z - s * y;
        m(2, 1) = t * y * z + s * x;
        m(2, 2) = t * z * z + c;
        return m;
    }

  private:
    double angle_;
    Vector3d axis_;
};

class Translation3d
{
  public:
    Translation3d(double x, double y, double z) : t_(x, y, z) {}
    explicit Translation3d(Vector3d const& t) : t_(t) {}
    Vector3d const& vector() const { return t_; }

  private:
    Vector3d t_;
};

class Affine3d
{
  public:
    Affine3d() : lin_(Matrix3d::Identity()) {}
    Affine3d(Translation3d const& t) : lin_(Matrix3d::Identity()), t_(t.vector()) {}
    static Affine3d Identity() { return Affine3d(); }
    Affine3d& rotate(AngleAxisd const& r)
    {
        lin_ = lin_ * r.toRotationMatrix();
        return *this;
    }
    Affine3d& scale(Vector3d const& s)
    {
        Matrix3d d;
        d(0, 0) = s.x();
        d(1, 1) = s.y();
        d(2, 2) = s.z();
        lin_    = lin_ * d;
        return *this;
    }
    Affine3d& translate(Vector3d const& t)
    {
        t_ += lin_ * t;
        return *this;
    }
    Matrix3d& linear() { return lin_; }
    Matrix3d const& linear() const { return lin_; }
    Vector3d& translation() { return t_; }
    Vector3d const& translation() const { return t_; }
    // affine * p.homogeneous() -> 3-vector (the AffineCompact-style product Eigen returns)
    Vector3d operator*(Vector4d const& h) const { return lin_ * Vector3d(h(0), h(1), h(2)) + t_ * h(3); }
    Vector3d operator*(Vector3d const& p) const { return lin_ * p + t_; }

  private:
    Matrix3d lin_;
    Vector3d t_;
};

// Two-sided Jacobi SVD of a square matrix: sigma sorted descending, non-negative, full U and V.
template <typename M>
class JacobiSVD
{
  public:
    using T = typename M::Scalar;
    using Vec = Matrix<T, 3, 1>;
    JacobiSVD(M const& A, unsigned int /*options*/ = 0) { compute(A); }
    Vec const& singularValues() const { return s_; }
    M const& matrixU() const { return U_; }
    M const& matrixV() const { return V_; }

  private:
    static void cols_times(M& m, int p, int q, T const g[4])
    {
        for (int k = 0; k < 3; ++k)
        {
            T const a = m(k, p), b = m(k, q);
            m(k, p)   = a * g[0] + b * g[2];
            m(k, q)   = a * g[1] + b * g[3];
        }
    }
    static void rows_timesT(M& m, int p, int q, T const g[4])
    {
        for (int k = 0; k < 3; ++k)
        {
            T const a = m(p, k), b = m(q, k);
            m(p, k)   = g[0] * a + g[2] * b;
            m(q, k)   = g[1] * a + g[3] * b;
        }
    }
    void compute(M const& A)
    {
        T scale = T(0);
        for (int i = 0; i < 9; ++i)
            scale = std::max(scale, std::abs(A.data()[i]));
        if (scale == T(0))
            scale = T(1);
        M W = A / scale;
        U_  = M::Identity();
        V_  = M::Identity();
        T const precision = T(2) * std::numeric_limits<T>::epsilon();
        T const tiny      = std::numeric_limits<T>::min();
        for (int sweep = 0; sweep < 64; ++sweep)
        {
            bool finished = true;
            for (int q = 1; q < 3; ++q)
                for (int p = 0; p < q; ++p)
                {
                    T const thr = std::max(tiny, precision * std::max(std::abs(W(p, p)), std::abs(W(q, q))));
                    if (std::abs(W(p, q)) <= thr && std::abs(W(q, p)) <= thr)
                        continue;
                    finished   = false;
                    T const a = W(p, p), b = W(p, q), c = W(q, p), d = W(q, q);
                    T c1 = T(1), s1 = T(0);
                    T const h = std::hypot(a + d, b - c);
                    if (h > tiny)
                    {
                        c1 = (a + d) / h;
                        s1 = (b - c) / h;
                    }
                    T const x = c1 * a - s1 * c, y = c1 * b - s1 * d, z = s1 * b + c1 * d;
                    T cj = T(1), sj = T(0);
                    if (std::abs(y) > tiny)
                    {
                        T const tau = (z - x) / (T(2) * y);
                        T const t   = tau >= 0 ? T(1) / (tau + std::sqrt(T(1) + tau * tau))
                                               : T(1) / (tau - std::sqrt(T(1) + tau * tau));
                        cj = T(1) / std::sqrt(T(1) + t * t);
                        sj = t * cj;
                    }
                    T const Rm[4] = {c1, s1, -s1, c1};
                    T const J[4]  = {cj, sj, -sj, cj};
                    T const L[4]  = {Rm[0] * J[0] + Rm[1] * J[2], Rm[0] * J[1] + Rm[1] * J[3],
                                     Rm[2] * J[0] + Rm[3] * J[2], Rm[2] * J[1] + Rm[3] * J[3]};
                    rows_timesT(W, p, q, L);
                    cols_times(W, p, q, J);
                    cols_times(U_, p, q, L);
                    cols_times(V_, p, q, J);
                }
            if (finished)
                break;
        }
        for (int i = 0; i < 3; ++i)
        {
            T a = W(i, i);
            if (a < T(0))
            {
                a = -a;
                for (int k = 0; k < 3; ++k)
                    U_(k, i) = -U_(k, i);
            }
            s_(i) = a * scale;
        }
        for (int i = 0; i < 3; ++i)
        {
            int best = i;
            for (int j = i + 1; j < 3; ++j)
                if (s_(j) > s_(best))
                    best = j;
            if (best != i)
            {
                std::swap(s_(i), s_(best));
                for (int k = 0; k < 3; ++k)
                {
                    std::swap(U_(k, i), U_(k, best));
                    std::swap(V_(k, i), V_(k, best));
                }
            }
        }
    }
    M U_, V_;
    Vec s_;
};

} // namespace Eigen
