// linalg.hpp -- the few dense linear-algebra pieces the host mirror needs.
//
// The reference uses Eigen (VectorXd, MatrixXd, LLT, SelfAdjointEigenSolver); Eigen is not installed in this
// image, so the host mirror ships a minimal stand-in with the same spellings for the operations it uses.
// A reference-side build defines SEPAIHRD_HOST_USE_EIGEN and gets the real types instead.
#pragma once

#ifdef SEPAIHRD_HOST_USE_EIGEN
#include <Eigen/Dense>
namespace epidemic {
using VectorXd = Eigen::VectorXd;
using MatrixXd = Eigen::MatrixXd;
}  // namespace epidemic
#else

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <stdexcept>
#include <vector>

namespace epidemic {

class VectorXd {
public:
    VectorXd() = default;
    explicit VectorXd(std::ptrdiff_t n) : v_(static_cast<size_t>(n)) {}
    VectorXd(std::initializer_list<double> il) : v_(il) {}
    static VectorXd Zero(std::ptrdiff_t n) { VectorXd r(n); std::fill(r.v_.begin(), r.v_.end(), 0.0); return r; }
    static VectorXd Constant(std::ptrdiff_t n, double c) { VectorXd r(n); std::fill(r.v_.begin(), r.v_.end(), c); return r; }
    static VectorXd FromPointer(const double* p, std::ptrdiff_t n) { VectorXd r(n); std::copy(p, p + n, r.v_.begin()); return r; }
    std::ptrdiff_t size() const { return static_cast<std::ptrdiff_t>(v_.size()); }
    void resize(std::ptrdiff_t n) { v_.resize(static_cast<size_t>(n)); }
    double& operator()(std::ptrdiff_t i) { return v_[static_cast<size_t>(i)]; }
    double operator()(std::ptrdiff_t i) const { return v_[static_cast<size_t>(i)]; }
    double& operator[](std::ptrdiff_t i) { return v_[static_cast<size_t>(i)]; }
    double operator[](std::ptrdiff_t i) const { return v_[static_cast<size_t>(i)]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
    double sum() const { double s = 0; for (double x : v_) s += x; return s; }
    void setZero() { std::fill(v_.begin(), v_.end(), 0.0); }
    bool operator==(const VectorXd& o) const { return v_ == o.v_; }
    VectorXd& operator+=(const VectorXd& o) { for (size_t i = 0; i < v_.size(); ++i) v_[i] += o.v_[i]; return *this; }
    VectorXd& operator-=(const VectorXd& o) { for (size_t i = 0; i < v_.size(); ++i) v_[i] -= o.v_[i]; return *this; }
    VectorXd& operator*=(double c) { for (double& x : v_) x *= c; return *this; }
    VectorXd& operator/=(double c) { for (double& x : v_) x /= c; return *this; }
    friend VectorXd operator+(VectorXd a, const VectorXd& b) { a += b; return a; }
    friend VectorXd operator-(VectorXd a, const VectorXd& b) { a -= b; return a; }
    friend VectorXd operator*(double c, VectorXd a) { a *= c; return a; }
    friend VectorXd operator*(VectorXd a, double c) { a *= c; return a; }
    const std::vector<double>& std() const { return v_; }
private:
    std::vector<double> v_;
};

// column-major like Eigen's default: data()[j * rows + i] == (*this)(i, j)
class MatrixXd {
public:
    MatrixXd() = default;
    MatrixXd(std::ptrdiff_t r, std::ptrdiff_t c) : r_(r), c_(c), v_(static_cast<size_t>(r * c), 0.0) {}
    static MatrixXd Zero(std::ptrdiff_t r, std::ptrdiff_t c) { return MatrixXd(r, c); }
    static MatrixXd Identity(std::ptrdiff_t r, std::ptrdiff_t c) { MatrixXd m(r, c); for (std::ptrdiff_t i = 0; i < std::min(r, c); ++i) m(i, i) = 1.0; return m; }
    std::ptrdiff_t rows() const { return r_; }
    std::ptrdiff_t cols() const { return c_; }
    std::ptrdiff_t size() const { return r_ * c_; }
    void resize(std::ptrdiff_t r, std::ptrdiff_t c) { r_ = r; c_ = c; v_.assign(static_cast<size_t>(r * c), 0.0); }
    double& operator()(std::ptrdiff_t i, std::ptrdiff_t j) { return v_[static_cast<size_t>(j * r_ + i)]; }
    double operator()(std::ptrdiff_t i, std::ptrdiff_t j) const { return v_[static_cast<size_t>(j * r_ + i)]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
    VectorXd row(std::ptrdiff_t i) const { VectorXd r(c_); for (std::ptrdiff_t j = 0; j < c_; ++j) r(j) = (*this)(i, j); return r; }
    MatrixXd transpose() const { MatrixXd t(c_, r_); for (std::ptrdiff_t i = 0; i < r_; ++i) for (std::ptrdiff_t j = 0; j < c_; ++j) t(j, i) = (*this)(i, j); return t; }
    double trace() const { double s = 0; for (std::ptrdiff_t i = 0; i < std::min(r_, c_); ++i) s += (*this)(i, i); return s; }
    MatrixXd& operator+=(const MatrixXd& o) { for (size_t i = 0; i < v_.size(); ++i) v_[i] += o.v_[i]; return *this; }
    MatrixXd& operator*=(double c) { for (double& x : v_) x *= c; return *this; }
    friend MatrixXd operator+(MatrixXd a, const MatrixXd& b) { a += b; return a; }
    friend MatrixXd operator*(double c, MatrixXd a) { a *= c; return a; }
    friend MatrixXd operator*(MatrixXd a, double c) { a *= c; return a; }
    friend VectorXd operator*(const MatrixXd& m, const VectorXd& x) {
        VectorXd y = VectorXd::Zero(m.r_);
        for (std::ptrdiff_t j = 0; j < m.c_; ++j) for (std::ptrdiff_t i = 0; i < m.r_; ++i) y(i) += m(i, j) * x(j);
        return y;
    }
    friend MatrixXd operator*(const MatrixXd& a, const MatrixXd& b) {
        MatrixXd c(a.r_, b.c_);
        for (std::ptrdiff_t j = 0; j < b.c_; ++j) for (std::ptrdiff_t k = 0; k < a.c_; ++k) for (std::ptrdiff_t i = 0; i < a.r_; ++i) c(i, j) += a(i, k) * b(k, j);
        return c;
    }
private:
    std::ptrdiff_t r_ = 0, c_ = 0;
    std::vector<double> v_;
};

}  // namespace epidemic
#endif  // SEPAIHRD_HOST_USE_EIGEN

namespace epidemic {
namespace linalg {

// Lower Cholesky factor (Eigen::LLT::matrixL). Returns false when the matrix is not positive definite.
inline bool cholesky_lower(const MatrixXd& a, MatrixXd& L) {
    const std::ptrdiff_t n = a.rows();
    L = MatrixXd::Zero(n, n);
    for (std::ptrdiff_t j = 0; j < n; ++j) {
        double d = a(j, j);
        for (std::ptrdiff_t k = 0; k < j; ++k) d -= L(j, k) * L(j, k);
        if (!(d > 0.0)) return false;
        L(j, j) = std::sqrt(d);
        for (std::ptrdiff_t i = j + 1; i < n; ++i) {
            double s = a(i, j);
            for (std::ptrdiff_t k = 0; k < j; ++k) s -= L(i, k) * L(j, k);
            L(i, j) = s / L(j, j);
        }
    }
    return true;
}

// Symmetric eigendecomposition by cyclic Jacobi rotations (stands in for Eigen::SelfAdjointEigenSolver).
inline void symmetric_eigen(MatrixXd a, VectorXd& evals, MatrixXd& evecs) {
    const std::ptrdiff_t n = a.rows();
    evecs = MatrixXd::Identity(n, n);
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0;
        for (std::ptrdiff_t p = 0; p < n; ++p) for (std::ptrdiff_t q = p + 1; q < n; ++q) off += a(p, q) * a(p, q);
        if (off < 1e-300) break;
        for (std::ptrdiff_t p = 0; p < n; ++p)
            for (std::ptrdiff_t q = p + 1; q < n; ++q) {
                if (std::fabs(a(p, q)) < 1e-300) continue;
                const double theta = (a(q, q) - a(p, p)) / (2.0 * a(p, q));
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (std::ptrdiff_t k = 0; k < n; ++k) {
                    const double akp = a(k, p), akq = a(k, q);
                    a(k, p) = c * akp - s * akq; a(k, q) = s * akp + c * akq;
                }
                for (std::ptrdiff_t k = 0; k < n; ++k) {
                    const double apk = a(p, k), aqk = a(q, k);
                    a(p, k) = c * apk - s * aqk; a(q, k) = s * apk + c * aqk;
                }
                for (std::ptrdiff_t k = 0; k < n; ++k) {
                    const double vkp = evecs(k, p), vkq = evecs(k, q);
                    evecs(k, p) = c * vkp - s * vkq; evecs(k, q) = s * vkp + c * vkq;
                }
            }
    }
    evals = VectorXd(n);
    for (std::ptrdiff_t i = 0; i < n; ++i) evals(i) = a(i, i);
}

}  // namespace linalg
}  // namespace epidemic
