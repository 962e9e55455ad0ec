/*
 * m3d_eigen_shim.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal stand-in for the slice of Eigen that the reference's hot-path headers
 * (include/misc3d/common/ransac.h, include/misc3d/utils.h) use, so that those headers can be
 * compiled UNMODIFIED from /root/reference in a container that has no Eigen (oracle/Makefile,
 * target _ref).  It is not Eigen and copies nothing from it: a dense column-major matrix with the
 * handful of members those headers call.  Where Eigen's evaluation order matters for bit patterns
 * the order documented in SURVEY.md Appendix D is used (3-vectors linear, 4-vectors in SSE2 pairs,
 * the 3.4 determinant), i.e. the same orders oracle/m3d_oracle.cpp restates -- what the compiled
 * reference adds is the reference's own control flow and formulas, not an independent Eigen.
 */
#pragma once
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <type_traits>
#include <vector>

namespace Eigen {

constexpr int Dynamic = -1;
enum DecompositionOptions { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };

template <typename T, int R, int C>
class Matrix;

namespace shim {
template <typename T, int R, int C, bool Dyn = (R == Dynamic || C == Dynamic)>
struct Storage;
template <typename T, int R, int C>
struct Storage<T, R, C, false> {
    T d[R * C];
    Storage() {
        for (int i = 0; i < R * C; ++i) d[i] = T(0);
    }
    int rows() const { return R; }
    int cols() const { return C; }
    void resize(long r, long c) { assert(r == R && c == C); }
    T *data() { return d; }
    const T *data() const { return d; }
};
template <typename T, int R, int C>
struct Storage<T, R, C, true> {
    std::vector<T> d;
    long r_ = (R == Dynamic ? 0 : R), c_ = (C == Dynamic ? 0 : C);
    int rows() const { return (int)r_; }
    int cols() const { return (int)c_; }
    void resize(long r, long c) {
        r_ = r;
        c_ = c;
        d.assign((size_t)(r * c), T(0));
    }
    T *data() { return d.data(); }
    const T *data() const { return d.data(); }
};

/* dot product in the order Eigen's SSE2 (2-double packet) reduction produces for the sizes the
 * hot path uses: size 3 -> linear, size 4 -> (0,2)+(1,3); other sizes linear */
template <typename T>
inline T dot_n(const T *a, const T *b, long n) {
    if (n == 4) return (a[0] * b[0] + a[2] * b[2]) + (a[1] * b[1] + a[3] * b[3]);
    if (n == 0) return T(0);
    T s = a[0] * b[0];
    for (long i = 1; i < n; ++i) s = s + a[i] * b[i];
    return s;
}
}  // namespace shim

template <typename T>
struct ArrayX; /* coefficient-wise view used by SphereEstimator::GeneralFit */

template <typename M>
struct TransposeView {
    const M &m;
    /* row-vector * column-vector -> scalar (the only product the hot path forms) */
    template <int R2, int C2>
    typename M::Scalar operator*(const Matrix<typename M::Scalar, R2, C2> &v) const {
        assert(m.size() == v.size());
        return shim::dot_n(m.data(), v.data(), (long)m.size());
    }
};

/* writable view of one column / one row */
template <typename M, bool IsCol>
struct LineRef {
    M &m;
    long k;
    using T = typename M::Scalar;
    long size() const { return IsCol ? m.rows() : m.cols(); }
    using Ref = typename std::conditional<std::is_const<M>::value, const T &, T &>::type;
    Ref at(long i) const { return IsCol ? m(i, k) : m(k, i); }
    template <int R2, int C2>
    LineRef &operator=(const Matrix<T, R2, C2> &v) {
        assert((long)v.size() == size());
        for (long i = 0; i < size(); ++i) at(i) = v.data()[i];
        return *this;
    }
    template <typename M2, bool C2>
    LineRef &operator=(const LineRef<M2, C2> &o) {
        for (long i = 0; i < size(); ++i) at(i) = o.at(i);
        return *this;
    }
    LineRef &operator=(const LineRef &o) {
        for (long i = 0; i < size(); ++i) at(i) = o.at(i);
        return *this;
    }
    Matrix<T, Dynamic, 1> transpose() const; /* as a plain dynamic vector (orientation-free) */
    ArrayX<T> array() const;
    operator Matrix<T, Dynamic, 1>() const { return transpose(); }
};

template <typename T, int R, int C>
class Matrix : public shim::Storage<T, R, C> {
    using S = shim::Storage<T, R, C>;

public:
    using Scalar = T;
    static constexpr bool kVector = (C == 1);
    Matrix() {}
    /* dynamic vector of a given size / fixed-size coefficient constructors */
    template <typename I, typename = typename std::enable_if<std::is_integral<I>::value && (R == Dynamic) && (C == 1)>::type>
    explicit Matrix(I n) {
        S::resize((long)n, 1);
    }
    Matrix(T a, T b) {
        static_assert(R * C == 2, "");
        this->d[0] = a, this->d[1] = b;
    }
    Matrix(T a, T b, T c) {
        static_assert(R * C == 3, "");
        this->d[0] = a, this->d[1] = b, this->d[2] = c;
    }
    Matrix(T a, T b, T c, T e) {
        static_assert(R * C == 4, "");
        this->d[0] = a, this->d[1] = b, this->d[2] = c, this->d[3] = e;
    }
    template <int R2, int C2, typename = typename std::enable_if<(R2 != R || C2 != C)>::type>
    Matrix(const Matrix<T, R2, C2> &o) {
        *this = o;
    }
    template <int R2, int C2>
    typename std::enable_if<(R2 != R || C2 != C), Matrix &>::type operator=(const Matrix<T, R2, C2> &o) {
        S::resize(o.rows(), o.cols());
        for (size_t i = 0; i < o.size(); ++i) data()[i] = o.data()[i];
        return *this;
    }

    using S::cols;
    using S::data;
    using S::rows;
    size_t size() const { return (size_t)rows() * (size_t)cols(); }
    void resize(long r, long c) { S::resize(r, c); }
    void resize(long n) { S::resize(n, 1); }
    Matrix &setZero() {
        for (size_t i = 0; i < size(); ++i) data()[i] = T(0);
        return *this;
    }
    Matrix &setZero(long n) {
        S::resize(n, 1);
        return setZero();
    }
    Matrix &setZero(long r, long c) {
        S::resize(r, c);
        return setZero();
    }
    Matrix &setOnes(long r, long c) {
        S::resize(r, c);
        for (size_t i = 0; i < size(); ++i) data()[i] = T(1);
        return *this;
    }

    T &operator()(long i) { return data()[i]; }
    const T &operator()(long i) const { return data()[i]; }
    T &operator[](long i) { return data()[i]; }
    const T &operator[](long i) const { return data()[i]; }
    T &operator()(long i, long j) { return data()[i + j * (long)rows()]; }
    const T &operator()(long i, long j) const { return data()[i + j * (long)rows()]; }

    LineRef<Matrix, true> col(long j) { return {*this, j}; }
    LineRef<const Matrix, true> col(long j) const { return {*this, j}; }
    LineRef<Matrix, false> row(long i) { return {*this, i}; }
    LineRef<const Matrix, false> row(long i) const { return {*this, i}; }

    template <int N>
    Matrix<T, N, 1> head() const {
        Matrix<T, N, 1> r;
        for (int i = 0; i < N; ++i) r(i) = data()[i];
        return r;
    }
    TransposeView<Matrix> transpose() const { return {*this}; }

    /* ---- arithmetic (element order = storage order) */
    Matrix operator+(const Matrix &o) const {
        Matrix r = *this;
        for (size_t i = 0; i < size(); ++i) r.data()[i] = data()[i] + o.data()[i];
        return r;
    }
    Matrix operator-(const Matrix &o) const {
        Matrix r = *this;
        for (size_t i = 0; i < size(); ++i) r.data()[i] = data()[i] - o.data()[i];
        return r;
    }
    Matrix operator-() const {
        Matrix r = *this;
        for (size_t i = 0; i < size(); ++i) r.data()[i] = -data()[i];
        return r;
    }
    Matrix operator*(T s) const {
        Matrix r = *this;
        for (size_t i = 0; i < size(); ++i) r.data()[i] = data()[i] * s;
        return r;
    }
    Matrix operator/(T s) const {
        Matrix r = *this;
        for (size_t i = 0; i < size(); ++i) r.data()[i] = data()[i] / s;
        return r;
    }
    friend Matrix operator*(T s, const Matrix &m) {
        Matrix r = m;
        for (size_t i = 0; i < m.size(); ++i) r.data()[i] = s * m.data()[i];
        return r;
    }
    Matrix &operator+=(const Matrix &o) {
        for (size_t i = 0; i < size(); ++i) data()[i] += o.data()[i];
        return *this;
    }
    Matrix &operator-=(const Matrix &o) {
        for (size_t i = 0; i < size(); ++i) data()[i] -= o.data()[i];
        return *this;
    }
    Matrix &operator*=(T s) {
        for (size_t i = 0; i < size(); ++i) data()[i] *= s;
        return *this;
    }
    Matrix &operator/=(T s) {
        for (size_t i = 0; i < size(); ++i) data()[i] /= s;
        return *this;
    }

    T dot(const Matrix &o) const { return shim::dot_n(data(), o.data(), (long)size()); }
    T squaredNorm() const { return shim::dot_n(data(), data(), (long)size()); }
    T norm() const { return std::sqrt(squaredNorm()); }
    void normalize() {
        const T z = squaredNorm();
        if (z > T(0)) *this /= std::sqrt(z);
    }
    /* each component mul, mul, sub */
    Matrix cross(const Matrix &b) const {
        static_assert(R * C == 3, "cross() is for 3-vectors");
        const Matrix &a = *this;
        return Matrix(a(1) * b(2) - a(2) * b(1), a(2) * b(0) - a(0) * b(2), a(0) * b(1) - a(1) * b(0));
    }

    /* 4x4 determinant in the operation order of Eigen 3.4 */
    T determinant() const {
        static_assert(R == 4 && C == 4, "only the 4x4 determinant is needed");
        const Matrix &m = *this;
        auto d2 = [&](int i, int j) { return m(i, 0) * m(j, 1) - m(j, 0) * m(i, 1); };
        auto d3 = [&](int i0, T a, int i1, T b, int i2, T c) { return m(i0, 2) * a + (-m(i1, 2) * b + m(i2, 2) * c); };
        const T d01 = d2(0, 1), d02 = d2(0, 2), d03 = d2(0, 3), d12 = d2(1, 2), d13 = d2(1, 3), d23 = d2(2, 3);
        const T d3_0 = d3(1, d23, 2, d13, 3, d12);
        const T d3_1 = d3(0, d23, 2, d03, 3, d02);
        const T d3_2 = d3(0, d13, 1, d03, 3, d01);
        const T d3_3 = d3(0, d12, 1, d02, 2, d01);
        return (-m(0, 3) * d3_0 + m(1, 3) * d3_1) + (-m(2, 3) * d3_2 + m(3, 3) * d3_3);
    }

    /* least squares: `A.bdcSvd(flags).solve(b)` -> minimum-norm LS solution.  Solved here by
     * Householder QR in long double (full column rank assumed); agrees with any stable SVD solve
     * to rounding -- SphereEstimator::GeneralFit is compared with a tolerance, never bitwise. */
    struct LsqSolver {
        const Matrix &A;
        Matrix<T, Dynamic, 1> solve(const Matrix<T, Dynamic, 1> &b) const {
            const long m = A.rows(), n = A.cols();
            std::vector<long double> a((size_t)(m * n)), y((size_t)m);
            for (long j = 0; j < n; ++j)
                for (long i = 0; i < m; ++i) a[(size_t)(i + j * m)] = A(i, j);
            for (long i = 0; i < m; ++i) y[(size_t)i] = b(i);
            for (long k = 0; k < n; ++k) {
                long double nrm = 0;
                for (long i = k; i < m; ++i) nrm += a[(size_t)(i + k * m)] * a[(size_t)(i + k * m)];
                nrm = std::sqrt(nrm);
                if (nrm == 0) continue;
                const long double alpha = a[(size_t)(k + k * m)] > 0 ? -nrm : nrm;
                std::vector<long double> v((size_t)(m - k));
                for (long i = k; i < m; ++i) v[(size_t)(i - k)] = a[(size_t)(i + k * m)];
                v[0] -= alpha;
                long double vv = 0;
                for (auto e : v) vv += e * e;
                if (vv == 0) continue;
                for (long j = k; j < n; ++j) {
                    long double s = 0;
                    for (long i = k; i < m; ++i) s += v[(size_t)(i - k)] * a[(size_t)(i + j * m)];
                    s = 2 * s / vv;
                    for (long i = k; i < m; ++i) a[(size_t)(i + j * m)] -= s * v[(size_t)(i - k)];
                }
                long double s = 0;
                for (long i = k; i < m; ++i) s += v[(size_t)(i - k)] * y[(size_t)i];
                s = 2 * s / vv;
                for (long i = k; i < m; ++i) y[(size_t)i] -= s * v[(size_t)(i - k)];
            }
            Matrix<T, Dynamic, 1> x((int)n);
            for (long k = n - 1; k >= 0; --k) {
                long double s = y[(size_t)k];
                for (long j = k + 1; j < n; ++j) s -= a[(size_t)(k + j * m)] * (long double)x(j);
                x(k) = (T)(s / a[(size_t)(k + k * m)]);
            }
            return x;
        }
    };
    LsqSolver bdcSvd(unsigned = 0) const { return {*this}; }
    LsqSolver jacobiSvd(unsigned = 0) const { return {*this}; }
};

template <typename T>
struct ArrayX {
    std::vector<T> v;
    ArrayX pow(int e) const {
        ArrayX r = *this;
        for (auto &x : r.v) x = std::pow(x, e);
        return r;
    }
    ArrayX operator+(const ArrayX &o) const {
        ArrayX r = *this;
        for (size_t i = 0; i < v.size(); ++i) r.v[i] = v[i] + o.v[i];
        return r;
    }
    Matrix<T, Dynamic, 1> matrix() const {
        Matrix<T, Dynamic, 1> r((long)v.size());
        for (size_t i = 0; i < v.size(); ++i) r(i) = v[i];
        return r;
    }
};
template <typename M, bool IsCol>
Matrix<typename LineRef<M, IsCol>::T, Dynamic, 1> LineRef<M, IsCol>::transpose() const {
    Matrix<T, Dynamic, 1> r(size());
    for (long i = 0; i < size(); ++i) r(i) = at(i);
    return r;
}
template <typename M, bool IsCol>
ArrayX<typename LineRef<M, IsCol>::T> LineRef<M, IsCol>::array() const {
    ArrayX<T> r;
    r.v.resize((size_t)size());
    for (long i = 0; i < size(); ++i) r.v[(size_t)i] = at(i);
    return r;
}

/* Eigen::Map<const MatrixXd>(ptr, rows, cols): read-only column-major view (src/knn.cpp) */
template <typename M>
class Map {
public:
    using T = typename std::remove_const<M>::type::Scalar;
    Map(const T *p, long r, long c) : p_(p), r_(r), c_(c) {}
    long rows() const { return r_; }
    long cols() const { return c_; }
    struct ColView {
        const T *p;
        const T *data() const { return p; }
    };
    ColView col(long j) const { return {p_ + j * r_}; }

private:
    const T *p_;
    long r_, c_;
};

using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector4d = Matrix<double, 4, 1>;
using Vector6d = Matrix<double, 6, 1>; /* Open3D adds this alias to namespace Eigen */
using VectorXd = Matrix<double, Dynamic, 1>;
using Vector2i = Matrix<int, 2, 1>;
using Vector3i = Matrix<int, 3, 1>;
using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;
using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using Matrix3Xd = Matrix<double, 3, Dynamic>;

}  // namespace Eigen
