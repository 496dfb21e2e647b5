// TEST INFRASTRUCTURE ONLY.  Stand-in for OsqpEigen::Solver: mpc/include/qp/osqp_interface.h and gait_optimizer.h hold one as a
// member.  The MPC hot path never calls it (the OSQP branch is dead code).  GaitOptimizer::OptimizeContactTimes does
// (gait_optimizer.cpp:185-364): OSQP itself is absent from this image, so the stand-in RECORDS the problem the reference hands over
// (P, q, A, l, u -- the parity tests compare them with the restatement's) and returns the solution the test driver injected
// (OsqpEigen::recorded().x), which lets the reference's own post-processing of the step run; without one, solveProblem throws.
#pragma once
#include <Eigen/SparseCore>
#include <memory>
#include <stdexcept>

struct OSQPSettings { double rho, alpha, sigma, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf; int max_iter, polish, polishing, verbose, warm_starting, warm_start, scaling, linsys_solver, scaled_termination, check_termination, adaptive_rho; double time_limit; };
struct OSQPInfo { int status_val; int iter; double obj_val, prim_res, dual_res; char status[32]; };
struct OSQPSolverStub { OSQPInfo* info; OSQPSettings* settings; };

namespace OsqpEigen {
enum class Status { Solved = 1, SolvedInaccurate = 2, PrimalInfeasibleInaccurate = 3, PrimalInfeasible = -3, DualInfeasibleInaccurate = 4, DualInfeasible = -4,
                    MaxIterReached = -2, TimeLimitReached = -6, NonCvx = -7, Sigint = -5, Unsolved = -10 };
enum class ErrorExitFlag { NoError = 0, DataValidationError, SettingsValidationError, LinsysSolverLoadError, LinsysSolverInitError, NonCvxError, MemAllocError, WorkspaceNotInitError };

[[noreturn]] inline void unavailable() { throw std::runtime_error("ref_shim/OsqpEigen: compile-only stand-in"); }

struct Recorded {
    Eigen::SparseMatrix<double> A, P;
    Eigen::VectorXd l, u, q;
    Eigen::VectorXd x;        // injected solution
    bool have_x = false;
    int solves = 0;
};
inline Recorded& recorded() { static Recorded r; return r; }

class Settings {
public:
    OSQPSettings s_{};
    OSQPSettings* getSettings() { return &s_; }
    void setVerbosity(bool) {}
    void setPolish(bool) {}
    void setPrimalInfeasibilityTolerance(double) {}
    void setPrimalInfeasibilityTollerance(double) {}
    void setDualInfeasibilityTolerance(double) {}
    void setDualInfeasibilityTollerance(double) {}
    void setAbsoluteTolerance(double) {}
    void setRelativeTolerance(double) {}
    void setScaledTerimination(bool) {}
    void setMaxIteration(int) {}
    void setRho(double) {}
    void setAlpha(double) {}
    void setSigma(double) {}
    void setWarmStart(bool) {}
    void setScaling(int) {}
    void setLinearSystemSolver(int) {}
    void setTimeLimit(double) {}
    void setAdaptiveRho(bool) {}
    void setCheckTermination(int) {}
};
class Data {
public:
    void setNumberOfVariables(int) {}
    void setNumberOfConstraints(int) {}
    template <typename M> bool setHessianMatrix(const M& P) { recorded().P = P; return true; }
    template <typename M> bool setLinearConstraintsMatrix(const M& A) { recorded().A = A; return true; }
    template <typename V> bool setGradient(V& q) { recorded().q = q; return true; }
    template <typename V> bool setLowerBound(V& l) { recorded().l = l; return true; }
    template <typename V> bool setUpperBound(V& u) { recorded().u = u; return true; }
    template <typename V> bool setBounds(V& l, V& u) { recorded().l = l; recorded().u = u; return true; }
    void clearHessianMatrix() {}
    void clearLinearConstraintsMatrix() {}
    bool isSet() const { return false; }
};
class Solver {
public:
    Solver() : settings_(new Settings), data_(new Data) {}
    const std::unique_ptr<Settings>& settings() const { return settings_; }
    const std::unique_ptr<Data>& data() const { return data_; }
    bool initSolver() { return true; }
    bool isInitialized() const { return false; }
    void clearSolver() {}
    bool clearSolverVariables() { return true; }
    ErrorExitFlag solveProblem() {
        if (!recorded().have_x) throw std::runtime_error("ref_shim/OsqpEigen: no solution was injected (OSQP is not in this image)");
        recorded().solves++;
        return ErrorExitFlag::NoError;
    }
    Status getStatus() const { return recorded().have_x ? Status::Solved : Status::Unsolved; }
    Eigen::VectorXd getSolution() { return recorded().x; }
    Eigen::VectorXd getDualSolution() { return Eigen::VectorXd::Zero(recorded().l.size()); }
    double getObjValue() const { return 0; }
    template <typename V> bool setWarmStart(const V&, const V&) { return true; }
    template <typename V> bool setPrimalVariable(const V&) { return true; }
    template <typename V> bool setDualVariable(const V&) { return true; }
    template <typename V> bool updateGradient(const V&) { return true; }
    template <typename V> bool updateBounds(const V&, const V&) { return true; }
    template <typename M> bool updateHessianMatrix(const M&) { return true; }
    template <typename M> bool updateLinearConstraintsMatrix(const M&) { return true; }
    template <typename V> bool computeAdjointDerivative(V&, V&, V&) { unavailable(); }
    template <typename M> bool adjointDerivativeGetMat(M&, M&) { unavailable(); }
    template <typename V> bool adjointDerivativeGetVec(V&, V&, V&) { unavailable(); }
    const std::unique_ptr<OSQPSolverStub>& solver() const { return raw_; }
private:
    std::unique_ptr<Settings> settings_;
    std::unique_ptr<Data> data_;
    std::unique_ptr<OSQPSolverStub> raw_;
};
}  // namespace OsqpEigen
