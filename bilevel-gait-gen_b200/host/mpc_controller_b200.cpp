// bilevel-gait-gen_b200 -- see mpc_controller_b200.h
#include "mpc_controller_b200.h"

#include <algorithm>
#include <stdexcept>
#include <string>

namespace controller {

MPCController::MPCController(bgg_handle* mpc, int batch, int num_nodes, int gait_opt_freq, int ls_size)
    : mpc_(mpc), batch_(batch), gait_opt_freq_(gait_opt_freq), ls_size_(ls_size), z_stride_(12 * (num_nodes + 1) + 160),
      status_(batch, BGG_UNSOLVED), iters_(batch, 0), deriv_ready_(batch, 0), ls_best_(batch, -1),
      ls_quality_(static_cast<size_t>(batch) * ls_size, 0), alpha_(batch, 0.0), cost_(batch, 0.0), prev_cost_(batch, 1e10),   // :293
      cost_red_(batch, 0.0), dHdtheta_(static_cast<size_t>(batch) * BGG_NUM_EE * BGG_MAX_CONTACTS, 0.0),
      ls_costs_(static_cast<size_t>(batch) * ls_size, 0.0), z_(static_cast<size_t>(batch) * z_stride_, 0.0),
      q_des_(static_cast<size_t>(batch) * BGG_NQ, 0.0), v_des_(static_cast<size_t>(batch) * BGG_NV, 0.0),
      force_des_(static_cast<size_t>(batch) * BGG_NUM_EE * 3, 0.0), target_status_(batch, 0) {
    if (!mpc || batch <= 0 || gait_opt_freq <= 0 || ls_size <= 0) throw std::runtime_error("MPCController: bad arguments");
    for (int b = 0; b < batch; ++b) q_des_[static_cast<size_t>(b) * BGG_NQ + 6] = 1.0;   // pinocchio::neutral: identity quaternion
}

void MPCController::SetInitialConfig(const double* q) { std::copy(q, q + q_des_.size(), q_des_.begin()); }

int MPCController::GetTargetsFromTraj(const double* time) {
    // the device call overwrites q_des only where the robot's two IK solves converged
    return bgg_targets_from_traj_batch(mpc_, time, q_des_.data(), v_des_.data(), force_des_.data(), target_status_.data());
}

MPCController::Mode MPCController::NextMode() const {
    const int r = run_num_, f = gait_opt_freq_;
    const bool any_ready = std::any_of(deriv_ready_.begin(), deriv_ready_.end(), [](int32_t v) { return v != 0; });
    if (r % f == 0 && r > 0 && any_ready) return kLineSearch;       // :323
    if ((r + 1) % f == 0 && r > 0) return kSolveAndGaitOpt;         // :337
    return kSolve;
}

MPCController::Mode MPCController::MPCUpdate(const double* state, const double* time, const double* ee_locations) {
    const Mode mode = NextMode();
    const int rc = bgg_controller_tick_batch(mpc_, mode, ls_size_, state, time, ee_locations, status_.data(), iters_.data(), alpha_.data(),
                                             cost_.data(), z_.data(), z_stride_, deriv_ready_.data(), dHdtheta_.data(), ls_best_.data(),
                                             ls_costs_.data(), ls_quality_.data());
    if (rc != BGG_OK) throw std::runtime_error(std::string("bgg_controller_tick_batch: ") + bgg_last_error());
    if (mode != kLineSearch)
        for (int b = 0; b < batch_; ++b) {
            cost_red_[b] = prev_cost_[b] - cost_[b];
            prev_cost_[b] = cost_[b];
        }
    run_num_++;
    return mode;
}

MPCController::Mode MPCController::AdvanceWithoutDevice(const int32_t* deriv_ready) {
    const Mode mode = NextMode();
    for (int b = 0; b < batch_; ++b) deriv_ready_[b] = (mode == kSolveAndGaitOpt) ? deriv_ready[b] : 0;   // :335, :339, :344
    run_num_++;
    return mode;
}

}  // namespace controller

using controller::MPCController;
extern "C" {
void* bggc_create(bgg_handle* mpc, int batch, int num_nodes, int gait_opt_freq, int ls_size) {
    try {
        return new MPCController(mpc, batch, num_nodes, gait_opt_freq, ls_size);
    } catch (const std::exception&) {
        return nullptr;
    }
}
void bggc_destroy(void* c) { delete static_cast<MPCController*>(c); }
int bggc_next_mode(void* c) { return static_cast<MPCController*>(c)->NextMode(); }
int bggc_run_num(void* c) { return static_cast<MPCController*>(c)->run_num(); }
int bggc_mpc_update(void* c, const double* state, const double* time, const double* ee_locations) {
    try {
        return static_cast<MPCController*>(c)->MPCUpdate(state, time, ee_locations);
    } catch (const std::exception&) {
        return BGG_ESTATE;
    }
}
int bggc_advance_without_device(void* c, const int32_t* deriv_ready) { return static_cast<MPCController*>(c)->AdvanceWithoutDevice(deriv_ready); }
void bggc_set_initial_config(void* c, const double* q) { static_cast<MPCController*>(c)->SetInitialConfig(q); }
int bggc_targets_from_traj(void* cv, const double* time, double* q_des, double* v_des, double* force_des, int32_t* status) {
    MPCController* c = static_cast<MPCController*>(cv);
    const int rc = c->GetTargetsFromTraj(time);
    if (rc != BGG_OK) return rc;
    if (q_des) std::copy(c->q_des().begin(), c->q_des().end(), q_des);
    if (v_des) std::copy(c->v_des().begin(), c->v_des().end(), v_des);
    if (force_des) std::copy(c->force_des().begin(), c->force_des().end(), force_des);
    if (status) std::copy(c->target_status().begin(), c->target_status().end(), status);
    return BGG_OK;
}
void bggc_results(void* cv, int32_t* status, int32_t* iters, double* alpha, double* cost, double* cost_red, int32_t* deriv_ready,
                  double* dHdtheta, int32_t* ls_best, double* ls_costs, int32_t* ls_quality) {
    const MPCController& c = *static_cast<MPCController*>(cv);
    auto put = [](const auto& v, auto* out) { if (out) std::copy(v.begin(), v.end(), out); };
    put(c.status(), status); put(c.iters(), iters); put(c.alpha(), alpha); put(c.cost(), cost); put(c.cost_reduction(), cost_red);
    put(c.deriv_ready(), deriv_ready); put(c.dHdtheta(), dHdtheta); put(c.ls_best(), ls_best); put(c.ls_costs(), ls_costs);
    put(c.ls_quality(), ls_quality);
}
}
