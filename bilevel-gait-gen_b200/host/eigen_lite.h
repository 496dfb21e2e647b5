// bilevel-gait-gen_b200 -- stand-in for the handful of Eigen types on the reference's MPC call surface, used ONLY when
// <Eigen/Core> is not installed (this build image has no Eigen).  With Eigen present mpc_b200.h uses Eigen::VectorXd,
// Eigen::MatrixXd, Eigen::Vector3d / Vector2d directly, which is what the reference's callers pass
// (mpc/include/mpc.h:29-30, mpc/include/spline/end_effector_splines.h vector_3t).
#pragma once
#include <cassert>
#include <cstddef>
#include <initializer_list>
#include <vector>

namespace bgg_lite {

class VectorXd {
public:
    VectorXd() {}
    explicit VectorXd(int n) : v_(static_cast<size_t>(n), 0.0) {}
    VectorXd(std::initializer_list<double> l) : v_(l) {}
    static VectorXd Zero(int n) { return VectorXd(n); }
    static VectorXd Constant(int n, double c) { VectorXd x(n); for (auto& e : x.v_) e = c; return x; }
    int size() const { return static_cast<int>(v_.size()); }
    void resize(int n) { v_.assign(static_cast<size_t>(n), 0.0); }
    void setZero() { for (auto& e : v_) e = 0.0; }
    double& operator()(int i) { return v_[static_cast<size_t>(i)]; }
    double operator()(int i) const { return v_[static_cast<size_t>(i)]; }
    double& operator[](int i) { return v_[static_cast<size_t>(i)]; }
    double operator[](int i) const { return v_[static_cast<size_t>(i)]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
    double dot(const VectorXd& o) const { double s = 0; for (int i = 0; i < size(); ++i) s += v_[i] * o.v_[i]; return s; }

private:
    std::vector<double> v_;
};

template <int N>
class VectorNd {
public:
    VectorNd() { for (double& e : v_) e = 0.0; }
    VectorNd(double a, double b) { static_assert(N == 2, ""); v_[0] = a; v_[1] = b; }
    VectorNd(double a, double b, double c) { static_assert(N == 3, ""); v_[0] = a; v_[1] = b; v_[2] = c; }
    static VectorNd Zero() { return VectorNd(); }
    int size() const { return N; }
    double& operator()(int i) { return v_[i]; }
    double operator()(int i) const { return v_[i]; }
    double* data() { return v_; }
    const double* data() const { return v_; }

private:
    double v_[N];
};

class MatrixXd {   // column-major, as Eigen
public:
    MatrixXd() {}
    MatrixXd(int r, int c) : r_(r), c_(c), v_(static_cast<size_t>(r) * c, 0.0) {}
    static MatrixXd Zero(int r, int c) { return MatrixXd(r, c); }
    static MatrixXd Identity(int r, int c) { MatrixXd m(r, c); for (int i = 0; i < (r < c ? r : c); ++i) m(i, i) = 1.0; return m; }
    int rows() const { return r_; }
    int cols() const { return c_; }
    void resize(int r, int c) { r_ = r; c_ = c; v_.assign(static_cast<size_t>(r) * c, 0.0); }
    void setZero() { for (auto& e : v_) e = 0.0; }
    double& operator()(int i, int j) { return v_[static_cast<size_t>(j) * r_ + i]; }
    double operator()(int i, int j) const { return v_[static_cast<size_t>(j) * r_ + i]; }
    const double* data() const { return v_.data(); }

private:
    int r_ = 0, c_ = 0;
    std::vector<double> v_;
};

}  // namespace bgg_lite
