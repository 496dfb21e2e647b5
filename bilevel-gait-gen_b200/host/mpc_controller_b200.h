// bilevel-gait-gen_b200 -- controller::MPCController's MPC thread (controllers/include/mpc_controller.h:25-113,
// controllers/mpc_controller.cpp:286-399 MPCUpdate, :518-573 GaitOpt) for a BATCH of robots on the CUDA path: the three-mode
// schedule of the reference's while-loop body, one bgg_controller_tick_batch call per pass.
//
//     run_num % gait_opt_freq == 0 and a derivative is ready   ->  GaitOptimizer::LineSearch          (:323-336)
//     (run_num + 1) % gait_opt_freq == 0                        ->  GetRealTimeUpdate, then GaitOpt    (:337-340)
//     otherwise                                                 ->  GetRealTimeUpdate                  (:341-345)
//
// deriv_ready is kept per robot (GaitOpt returns false when the last solve was not Solved, mpc.cpp:1047-1057).  What the
// GetTargetsFromTraj (:414-511: two inverse-kinematics solves and a finite difference per robot) is the batch call
// bgg_targets_from_traj_batch.  What the reference's class does around this -- the std::thread and its five mutexes, the
// whole-body QPControl, visualisation -- is not on the path and stays with the caller.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/bgg.h"

namespace controller {

class MPCController {
public:
    enum Mode { kSolve = BGG_TICK_SOLVE, kSolveAndGaitOpt = BGG_TICK_SOLVE_GAIT_OPT, kLineSearch = BGG_TICK_LINE_SEARCH };
    static constexpr int LS_SIZE = 10;   // gait_optimizer.h:164

    // `mpc`: a handle with a batch (bgg_batch_reset) and costs set; not owned
    MPCController(bgg_handle* mpc, int batch, int num_nodes, int gait_opt_freq, int ls_size = LS_SIZE);

    Mode NextMode() const;   // the branch the coming pass takes (:323-345)

    // One pass of the while-loop body for every robot: state [batch][13], time [batch], ee_locations [batch][4][3].
    // Returns the mode that ran; results are in the accessors below.
    Mode MPCUpdate(const double* state, const double* time, const double* ee_locations);

    // bookkeeping of a pass without the device work (host-logic tests): the pass NextMode() names is taken as run, `deriv_ready`
    // is what its GaitOpt would have reported
    Mode AdvanceWithoutDevice(const int32_t* deriv_ready);

    // q_des_ before the first targets call: the robot's initial configuration (mpc_controller.cpp:50-52), [batch][BGG_NQ]
    void SetInitialConfig(const double* q);
    // MPCController::GetTargetsFromTraj (:414-511) for every robot at `time` [batch], from the trajectories of the last pass.
    // Needs bgg_set_kinematics on the handle.  Returns BGG_OK or a negative code; per-robot outcome in target_status()
    // (0; 1 "IK did not converge."; 2 "bad interp."; 3 time beyond the trajectory) -- q_des_ of a robot that failed is kept.
    int GetTargetsFromTraj(const double* time);
    const std::vector<double>& q_des() const { return q_des_; }                    // [batch][BGG_NQ]
    const std::vector<double>& v_des() const { return v_des_; }                    // [batch][BGG_NV]
    const std::vector<double>& force_des() const { return force_des_; }            // [batch][4][3]
    const std::vector<int32_t>& target_status() const { return target_status_; }

    int run_num() const { return run_num_; }
    const std::vector<int32_t>& status() const { return status_; }
    const std::vector<int32_t>& iters() const { return iters_; }
    const std::vector<double>& alpha() const { return alpha_; }
    const std::vector<double>& cost() const { return cost_; }
    const std::vector<double>& cost_reduction() const { return cost_red_; }       // prev_cost - cost (:340, 372-374)
    const std::vector<int32_t>& deriv_ready() const { return deriv_ready_; }
    const std::vector<double>& dHdtheta() const { return dHdtheta_; }              // [batch][4][BGG_MAX_CONTACTS], GaitOpt passes
    const std::vector<int32_t>& ls_best() const { return ls_best_; }
    const std::vector<double>& ls_costs() const { return ls_costs_; }              // [batch][ls_size], line-search passes
    const std::vector<int32_t>& ls_quality() const { return ls_quality_; }
    const std::vector<double>& trajectory() const { return z_; }                   // decision vectors [batch][z_stride()]
    int z_stride() const { return z_stride_; }
    int ls_size() const { return ls_size_; }

private:
    bgg_handle* mpc_;
    int batch_, gait_opt_freq_, ls_size_, z_stride_, run_num_ = 0;
    std::vector<int32_t> status_, iters_, deriv_ready_, ls_best_, ls_quality_;
    std::vector<double> alpha_, cost_, prev_cost_, cost_red_, dHdtheta_, ls_costs_, z_;
    std::vector<double> q_des_, v_des_, force_des_;
    std::vector<int32_t> target_status_;
};

}  // namespace controller

// C entry points over the class (for bindings; bilevel-gait-gen_b200/mpc_controller.py)
extern "C" {
void* bggc_create(bgg_handle* mpc, int batch, int num_nodes, int gait_opt_freq, int ls_size);
void bggc_destroy(void* c);
int bggc_next_mode(void* c);
int bggc_run_num(void* c);
// returns the mode that ran, or a negative BGG_E* code (bgg_last_error has the message)
int bggc_mpc_update(void* c, const double* state, const double* time, const double* ee_locations);
// copies of the last pass's results; any pointer may be NULL
int bggc_advance_without_device(void* c, const int32_t* deriv_ready);
void bggc_set_initial_config(void* c, const double* q);
// q_des / v_des / force_des / status: copies of the outcome, any may be NULL
int bggc_targets_from_traj(void* c, const double* time, double* q_des, double* v_des, double* force_des, int32_t* status);
void bggc_results(void* c, int32_t* status, int32_t* iters, double* alpha, double* cost, double* cost_red, int32_t* deriv_ready,
                  double* dHdtheta, int32_t* ls_best, double* ls_costs, int32_t* ls_quality);
}
