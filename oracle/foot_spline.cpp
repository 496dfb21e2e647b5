// TEST INFRASTRUCTURE ONLY -- CPU oracle (see foot_spline.hpp for scope and citations).
#include "foot_spline.hpp"

#include <cassert>
#include <cmath>

namespace oracle {

namespace {
// Hermite basis weights, end_effector_splines.cpp:1179-1197 (same expression shape, pow() included, so the
// values agree with the reference to the last bit on the same libm).
double hx0(double tau, double dT) { return 1 - (1 / std::pow(dT, 2)) * 3 * std::pow(tau, 2) + (1 / std::pow(dT, 3)) * 2 * std::pow(tau, 3); }
double hx1(double tau, double dT) { return (1 / std::pow(dT, 2)) * 3 * std::pow(tau, 2) - (1 / std::pow(dT, 3)) * 2 * std::pow(tau, 3); }
double hd0(double tau, double dT) { return tau - (1 / dT) * 2 * std::pow(tau, 2) + (1 / std::pow(dT, 2)) * std::pow(tau, 3); }
double hd1(double tau, double dT) { return -(1 / dT) * std::pow(tau, 2) + (1 / std::pow(dT, 2)) * std::pow(tau, 3); }

// Partials of the basis weights wrt a contact time, given d(tau)/d(theta) and d(DeltaT)/d(theta),
// end_effector_splines.cpp:1199-1244.
double hx0_d(double tau, double dT, double dtau, double ddT) {
    return (6 * std::pow(dT, -3) * std::pow(tau, 2) - 6 * std::pow(dT, -4) * std::pow(tau, 3)) * ddT +
           (-6 * std::pow(dT, -2) * tau + 6 * std::pow(dT, -3) * std::pow(tau, 2)) * dtau;
}
double hx1_d(double tau, double dT, double dtau, double ddT) {
    return (-6 * std::pow(dT, -3) * std::pow(tau, 2) + 6 * std::pow(dT, -4) * std::pow(tau, 3)) * ddT +
           (6 * std::pow(dT, -2) * tau - 6 * std::pow(dT, -3) * std::pow(tau, 2)) * dtau;
}
double hd0_d(double tau, double dT, double dtau, double ddT) {
    return (2 * std::pow(dT, -2) * std::pow(tau, 2) - 2 * std::pow(dT, -3) * std::pow(tau, 3)) * ddT +
           (1 - std::pow(dT, -1) * 4 * tau + std::pow(dT, -2) * 3 * std::pow(tau, 2)) * dtau;
}
double hd1_d(double tau, double dT, double dtau, double ddT) {
    return (std::pow(dT, -2) * std::pow(tau, 2) - 2 * std::pow(dT, -3) * std::pow(tau, 3)) * ddT +
           (-std::pow(dT, -1) * 2 * tau + std::pow(dT, -2) * 3 * std::pow(tau, 2)) * dtau;
}
}  // namespace

// end_effector_splines.cpp:34-153
FootSpline::FootSpline(int num_contacts, const std::vector<double>& times, bool start_in_contact, int num_force_polys)
    : num_force_polys_(num_force_polys) {
    if (num_force_polys_ < 2) {
        throw std::runtime_error("The number of force polynomials between constant sections must be at least 2.");
    }
    int pattern_len = num_force_polys + 1;
    spline_stride_ = num_force_polys;
    if (num_force_polys % 2) pattern_len++;

    std::vector<NodeType> fpat, ppat, zpat;
    std::vector<TimeType> tpat;
    auto push = [&](NodeType f, NodeType p, NodeType z, TimeType t) {
        fpat.push_back(f); ppat.push_back(p); zpat.push_back(z); tpat.push_back(t);
    };
    for (int i = 0; i < pattern_len; i++) {
        if (!start_in_contact) {
            if (i == 0) push(NoDeriv, NoDeriv, NoDeriv, LiftOff);
            else if (i == 1) push(Empty, Empty, FullDeriv, Inter);
            else if (i == 2) push(NoDeriv, NoDeriv, NoDeriv, TouchDown);
            else push(FullDeriv, Empty, Empty, Inter);
        } else {
            if (i == 0) push(NoDeriv, NoDeriv, NoDeriv, TouchDown);
            else if (i < num_force_polys) push(FullDeriv, Empty, Empty, Inter);
            else if (i == pattern_len - 1) push(Empty, Empty, FullDeriv, Inter);
            else push(NoDeriv, NoDeriv, NoDeriv, LiftOff);
        }
    }
    const int P = pattern_len;
    for (int coord = 0; coord < 3; coord++) {
        int i = 0, j = 0, k = 1;
        while (i < num_contacts) {
            const NodeType ft = fpat[j % P];
            forces_[coord].push_back(Knot{ft, {0.0, 0.0}});
            positions_[coord].push_back(Knot{coord < 2 ? ppat[j % P] : zpat[j % P], {0.0, 0.0}});
            if (ft == FullDeriv) {
                if (coord == 0) {
                    times_.push_back(KnotTime{times.at(i - 1) + k * (times.at(i) - times.at(i - 1)) / (num_force_polys), tpat[j % P]});
                    k++;
                }
            } else if (ft == Empty) {
                if (coord == 0) {
                    times_.push_back(KnotTime{times.at(i - 1) + (times.at(i) - times.at(i - 1)) / 2, tpat[j % P]});
                }
            } else {
                if (coord == 0) times_.push_back(KnotTime{times.at(i), tpat[j % P]});
                i++;
                k = 1;
            }
            j++;
        }
    }
    assert(forces_[1].size() == positions_[1].size());
    assert(forces_[1].size() == times_.size());
}

// end_effector_splines.cpp:1062-1084
int FootSpline::GetLowerNodeIdx(SplineType type, int coord, double time) const {
    const double t_first = times_.front().t, t_last = times_.back().t;
    if (time < t_first && time - t_first >= -1e-4) time = t_first;
    else if (time < t_first) throw std::runtime_error("Time requested is too small.");
    if (time > t_last && time - t_last <= 1e4) time = t_last;
    else if (time > t_last) throw std::runtime_error("Time requested is too large.");
    const std::vector<Knot>& s = Knots(type, coord);
    for (int i = static_cast<int>(times_.size()) - 1; i >= 0; i--) {
        if (time >= times_[i].t && s[i].type != Empty) return i;
    }
    throw std::runtime_error("Invalid time.");
}

// end_effector_splines.cpp:1086-1112
int FootSpline::GetUpperNodeIdx(SplineType type, int coord, double time) const {
    const double t_first = times_.front().t, t_last = times_.back().t;
    if (time < t_first && time - t_first >= -1e-4) time = t_first;
    else if (time < t_first) throw std::runtime_error("Time requested is too small.");
    if (time > t_last && time - t_last <= 1e4) time = t_last;
    else if (time > t_last) throw std::runtime_error("Time requested is too large.");
    const std::vector<Knot>& s = Knots(type, coord);
    for (int i = 0; i < static_cast<int>(times_.size()); i++) {
        if (time < times_[i].t && s[i].type != Empty) return i;
    }
    if (time == t_last) return static_cast<int>(times_.size()) - 1;
    throw std::runtime_error("Invalid time.");
}

static double KnotVar(const Knot& k, int i) {    // SplineNode::GetVars, spline_node.cpp:19-25
    if (k.type == Empty) throw std::runtime_error("Can't get the vars in this node. This node is set to empty.");
    return k.v[i];
}

// end_effector_splines.cpp:169-199
double FootSpline::ValueAt(SplineType type, int coord, double time) const {
    const std::vector<Knot>& s = Knots(type, coord);
    const int lo = GetLowerNodeIdx(type, coord, time);
    const int up = GetUpperNodeIdx(type, coord, time);
    if (up == lo) return KnotVar(s[lo], 0);
    const double dT = times_[up].t - times_[lo].t;
    const double tau = time - times_[lo].t;
    double x0 = KnotVar(s[lo], 0), x1 = KnotVar(s[up], 0), x0d = KnotVar(s[lo], 1), x1d = KnotVar(s[up], 1);
    if (type == Force) {
        x0d *= kForceMult;
        x1d *= kForceMult;
    }
    const double a2 = -(1 / std::pow(dT, 2)) * 3 * (x0 - x1) - (1 / dT) * (2 * x0d + x1d);
    const double a3 = (1 / std::pow(dT, 3)) * 2 * (x0 - x1) + (1 / std::pow(dT, 2)) * (x0d + x1d);
    return x0 + x0d * tau + a2 * std::pow(tau, 2) + a3 * std::pow(tau, 3);
}

// end_effector_splines.cpp:201-282
std::vector<double> FootSpline::GetPolyVarsLin(SplineType type, int coord, double time) const {
    const std::vector<Knot>& s = Knots(type, coord);
    const int lo = GetLowerNodeIdx(type, coord, time);
    const int up = GetUpperNodeIdx(type, coord, time);
    if (lo == up) return {1.0};
    const double tau = time - times_[lo].t;
    const double dT = times_[up].t - times_[lo].t;
    if (type == Force) {
        const NodeType tl = s[lo].type, tu = s[up].type;
        if (tl == NoDeriv && tu == NoDeriv) throw std::runtime_error("There is no mutable variables at the provided time.");
        if (tl == NoDeriv && tu == FullDeriv) return {hx1(tau, dT), hd1(tau, dT) * kForceMult};
        if (tl == FullDeriv && tu == NoDeriv) return {hx0(tau, dT), hd0(tau, dT) * kForceMult};
        return {hx0(tau, dT), hd0(tau, dT) * kForceMult, hx1(tau, dT), hd1(tau, dT) * kForceMult};
    }
    if (coord != 2) {
        if (forces_[coord].at(lo).type == NoDeriv && forces_[coord].at(lo + 2).type == NoDeriv) {
            return {hx0(tau, dT), hx1(tau, dT)};
        }
        return {1.0};
    }
    const NodeType tl = positions_[2][lo].type, tu = positions_[2][up].type;
    if (tl == NoDeriv && tu == FullDeriv) return {hx0(tau, dT), hx1(tau, dT), hd1(tau, dT)};
    if (tl == FullDeriv && tu == NoDeriv) return {hx0(tau, dT), hd0(tau, dT), hx1(tau, dT)};
    return {1.0};
}

// end_effector_splines.cpp:284-354
std::pair<int, int> FootSpline::GetVarsIdx(SplineType type, int coord, double time) const {
    const std::vector<Knot>& s = Knots(type, coord);
    const int lo = GetLowerNodeIdx(type, coord, time);
    const int up = GetUpperNodeIdx(type, coord, time);
    const std::vector<int> mut = GetMutableNodes(type, coord);
    int idx = 0;
    if (type == Force) {
        for (int i = 0; i < static_cast<int>(mut.size()); i++) {
            if (mut[i] < lo) idx = 2 * (i + 1);
        }
        const NodeType tl = s[lo].type, tu = s[up].type;
        if (tl == NoDeriv && tu == NoDeriv) throw std::runtime_error("There is no mutable variables at the provided time.");
        if ((tl == NoDeriv && tu == FullDeriv) || (tl == FullDeriv && tu == NoDeriv)) return {idx, 2};
        if (lo == up) return {idx, 1};
        return {idx, 4};
    }
    idx--;
    for (int m : mut) {
        if (m <= lo) idx++;
    }
    if (lo == up) return {idx, 1};
    if (coord != 2) {
        if (forces_[coord].at(lo).type == NoDeriv && forces_[coord].at(lo + 2).type == NoDeriv) return {idx, 2};
        return {idx, 1};
    }
    for (int m : mut) {
        if (positions_[2][m].type == FullDeriv && m < lo) idx++;
    }
    const NodeType tl = positions_[2][lo].type, tu = positions_[2][up].type;
    if ((tl == NoDeriv && tu == FullDeriv) || (tl == FullDeriv && tu == NoDeriv)) return {idx, 3};
    if (forces_[2].at(lo).type == NoDeriv && forces_[2].at(lo + 2).type == NoDeriv) return {idx, 2};
    return {idx, 1};
}

// end_effector_splines.cpp:356-364
bool FootSpline::IsForceMutable(double time) const {
    const int lo = GetLowerNodeIdx(Force, 0, time);
    const int up = GetUpperNodeIdx(Force, 0, time);
    return !(forces_[0][lo].type == NoDeriv && forces_[0][up].type == NoDeriv);
}

// end_effector_splines.cpp:366-449
void FootSpline::AddPoly(double additional_time) {
    const int n = GetNumNodes();
    auto mk = [](NodeType t) { return Knot{t, {0.0, 0.0}}; };
    if (forces_[0][n - 1].type == NoDeriv && forces_[0][n - 2].type == FullDeriv) {
        // the spline ends with a lift-off: append a swing (mid-swing knot + touch-down)
        for (int i = 0; i < 2; i++) {
            for (int coord = 0; coord < 3; coord++) {
                if (i == 0) {
                    forces_[coord].push_back(mk(Empty));
                    if (coord == 0) times_.push_back(KnotTime{times_.back().t + additional_time / 2, Inter});
                    positions_[coord].push_back(mk(coord == 2 ? FullDeriv : Empty));
                } else {
                    forces_[coord].push_back(mk(NoDeriv));
                    positions_[coord].push_back(mk(NoDeriv));
                    if (coord == 0) times_.push_back(KnotTime{times_.back().t + additional_time / 2, TouchDown});
                }
            }
        }
    } else {
        // the spline ends with a touch-down: append a stance (interior force knots + lift-off)
        for (int coord = 0; coord < 3; coord++) {
            for (int i = 0; i < num_force_polys_ - 1; i++) {
                forces_[coord].push_back(mk(FullDeriv));
                positions_[coord].push_back(mk(Empty));
                if (coord == 0) times_.push_back(KnotTime{times_.back().t + additional_time / num_force_polys_, Inter});
            }
            forces_[coord].push_back(mk(NoDeriv));
            positions_[coord].push_back(mk(NoDeriv));
            if (coord == 0) times_.push_back(KnotTime{times_.back().t + additional_time / num_force_polys_, LiftOff});
        }
    }
}

// end_effector_splines.cpp:451-465
void FootSpline::RemovePoly(double start_time) {
    const int lo = GetLowerNodeIdx(Position, 0, start_time);
    if (lo != 0) {
        times_.erase(times_.begin(), times_.begin() + lo);
        for (int coord = 0; coord < 3; coord++) {
            forces_[coord].erase(forces_[coord].begin(), forces_[coord].begin() + lo);
            positions_[coord].erase(positions_[coord].begin(), positions_[coord].begin() + lo);
        }
    }
    if (GetLowerNodeIdx(Position, 0, start_time) != 0) throw std::runtime_error("Poly remove did not work.");
}

// number of FullDeriv force knots walked back from `lower_node` (the loop at .cpp:602-609 / 686-693)
int FootSpline::ForceChainBack(int coord, int lower_node) const {
    int j = 0, idx = lower_node;
    while (forces_[coord].at(idx).type == FullDeriv) {
        j++;
        idx--;
    }
    return j;
}

// end_effector_splines.cpp:513-648
double FootSpline::ComputePartialWrtTime(SplineType type, int coord, double time, int time_idx) const {
    const std::vector<Knot>& s = Knots(type, coord);
    const int up = GetUpperNodeIdx(type, coord, time);
    const int lo = GetLowerNodeIdx(type, coord, time);
    const double dT = times_[up].t - times_[lo].t;
    const double tau = time - times_[lo].t;
    const int node = ConvertContactNodeToSplineNode(time_idx);
    const bool direct = (node == lo || node == up);
    const bool wrt_lower = (node == lo);
    const double P = static_cast<double>(num_force_polys_);

    const double x0 = KnotVar(s[lo], 0);
    const double x1 = KnotVar(s[up], 0);
    double x0d = 0, x1d = 0;
    if (s[lo].type == FullDeriv) x0d = (type == Force) ? s[lo].v[1] * kForceMult : s[lo].v[1];
    if (s[up].type == FullDeriv) x1d = (type == Force) ? s[up].v[1] * kForceMult : s[up].v[1];

    auto da2 = [&](double ddT) { return 6 * std::pow(dT, -3) * (x0 - x1) * ddT + (2 * x0d + x1d) * std::pow(dT, -2) * ddT; };
    auto da3 = [&](double ddT) { return -6 * std::pow(dT, -4) * (x0 - x1) * ddT - 2 * std::pow(dT, -3) * (x0d + x1d) * ddT; };
    const double a2 = -std::pow(dT, -2) * (3 * (x0 - x1) + dT * (2 * x0d + x1d));
    const double a3 = std::pow(dT, -3) * (2 * (x0 - x1) + dT * (x0d + x1d));

    if (direct && wrt_lower) {
        double ddT = -1.0;
        if (type == Force) ddT = -1.0 / P;
        return da2(ddT) * std::pow(tau, 2) + da3(ddT) * std::pow(tau, 3) - x0d - a2 * 2 * tau - a3 * 3 * std::pow(tau, 2);
    }
    if (direct && !wrt_lower) {
        double ddT = 1.0, dtau = 0.0;
        if (type == Force) {
            ddT = 1.0 / P;
            dtau = -static_cast<double>(num_force_polys_ - 1) / P;
        }
        return da2(ddT) * std::pow(tau, 2) + da3(ddT) * std::pow(tau, 3) + (x0d + a2 * 2 * tau + a3 * 3 * std::pow(tau, 2)) * dtau;
    }
    if (node > up && node <= GetUpperNodeIdx(Position, 0, time)) {
        const double ddT = 1.0 / P;
        const int j = ForceChainBack(coord, lo);
        const double dtau = -static_cast<double>(j) / P;
        return da2(ddT) * std::pow(tau, 2) + da3(ddT) * std::pow(tau, 3) + (x0d + a2 * 2 * tau + a3 * 3 * std::pow(tau, 2)) * dtau;
    }
    if (node < lo && node >= GetLowerNodeIdx(Position, 0, time)) {
        const double ddT = -1.0 / P;
        const int j = ForceChainBack(coord, lo);
        const double dtau = -(static_cast<double>(-j) / P + 1.0);
        return da2(ddT) * std::pow(tau, 2) + da3(ddT) * std::pow(tau, 3) + (x0d + a2 * 2 * tau + a3 * 3 * std::pow(tau, 2)) * dtau;
    }
    return 0;
}

// end_effector_splines.cpp:655-803
std::vector<double> FootSpline::ComputeCoefPartialWrtTime(SplineType type, int coord, double time, int time_idx,
                                                           double dtwdth) const {
    const std::vector<Knot>& s = Knots(type, coord);
    const int up = GetUpperNodeIdx(type, coord, time);
    const int lo = GetLowerNodeIdx(type, coord, time);
    double dT = times_[up].t - times_[lo].t;
    if (dT == 0) dT = times_[up].t - times_[GetLowerNodeIdx(type, coord, time - 1e-4)].t;
    const double tau = time - times_[lo].t;
    const std::pair<int, int> vi = GetVarsIdx(type, coord, time);
    std::vector<double> out(vi.second, 0.0);
    const int node = ConvertContactNodeToSplineNode(time_idx);
    const bool direct = (node == lo || node == up);
    bool wrt_lower = (node == lo);
    double dtau = dtwdth;
    const double P = static_cast<double>(num_force_polys_);

    if (type == Force) {
        const int j = ForceChainBack(coord, lo);
        double ddT = 1.0 / P;
        if (wrt_lower) {
            ddT = -1.0 / P;
            dtau += static_cast<double>(j) / P - 1.0;
        } else {
            dtau += -static_cast<double>(j) / P;
        }
        auto fill = [&]() {   // shared by the two indirect branches, .cpp:723-739 / 746-759
            if (s[lo].type == FullDeriv) {
                out.at(0) = hx0_d(tau, dT, dtau, ddT);
                out.at(1) = kForceMult * hd0_d(tau, dT, dtau, ddT);
                if (s[up].type == FullDeriv) {
                    out.at(2) = hx1_d(tau, dT, dtau, ddT);
                    out.at(3) = kForceMult * hd1_d(tau, dT, dtau, ddT);
                }
            } else if (s[up].type == FullDeriv) {
                out.at(0) = hx1_d(tau, dT, dtau, ddT);
                out.at(1) = kForceMult * hd1_d(tau, dT, dtau, ddT);
            }
        };
        if (direct) {
            if (wrt_lower) {
                out.at(0) = hx1_d(tau, dT, dtau, ddT);
                out.at(1) = hd1_d(tau, dT, dtau, ddT) * kForceMult;
            } else {
                out.at(0) = hx0_d(tau, dT, dtau, ddT);
                out.at(1) = hd0_d(tau, dT, dtau, ddT) * kForceMult;
            }
        } else if (node > up && node <= GetUpperNodeIdx(Position, 0, time)) {
            wrt_lower = false;
            ddT = 1.0 / P;
            dtau = dtwdth - static_cast<double>(j) / P;
            fill();
        } else if (node < lo && node >= GetLowerNodeIdx(Position, 0, time)) {
            wrt_lower = true;
            ddT = -1.0 / P;
            dtau = dtwdth + static_cast<double>(j) / P - 1.0;
            fill();
        }
        return out;
    }
    double ddT = 1.0;
    if (wrt_lower) {
        dtau += -1.0;
        ddT = -1.0;
    }
    if (direct) {
        if (lo == up) {
            out.at(0) = 0;
        } else if (positions_[2].at(lo + 1).type == FullDeriv) {
            out.at(0) = hx0_d(tau, dT, dtau, ddT);
            out.at(1) = hx1_d(tau, dT, dtau, ddT);
        } else {
            out.at(0) = 0;
        }
    }
    return out;
}

// end_effector_splines.cpp:805-813
bool FootSpline::IsInContact(double time) const {
    const int lo = GetLowerNodeIdx(Position, 0, time);
    const int up = GetUpperNodeIdx(Position, 0, time);
    return times_[lo].type == TouchDown && times_[up].type == LiftOff;
}

static void KnotSet(Knot& k, double v0, double v1) {   // SplineNode::SetVars, spline_node.cpp:31-41
    if (k.type == Empty) throw std::runtime_error("Can't set the vars in this node. This node is set to empty.");
    k.v[0] = v0;
    if (k.type != NoDeriv) k.v[1] = v1;
}

// end_effector_splines.cpp:815-858
void FootSpline::SetVars(SplineType type, int coord, int node_idx, double v0, double v1) {
    std::vector<Knot>& s = Sel(type, coord);
    const int n = static_cast<int>(s.size());
    if (s.at(node_idx).type == Empty) throw std::runtime_error("Can't set this node's variables. This node is empty.");
    if (type == Force && s[node_idx].type == NoDeriv) {
        throw std::runtime_error("Force spline cannot be changed at that node. Always set to 0.");
    }
    if (type == Position) {
        const bool not_fd = (coord != 2) || positions_[coord][node_idx].type != FullDeriv;
        if (node_idx < n - 1 && not_fd && forces_[coord].at(node_idx + 1).type == FullDeriv) {
            KnotSet(s[node_idx], v0, v1);
            KnotSet(s.at(node_idx + spline_stride_), v0, v1);
        } else if (node_idx > 0 && not_fd && forces_[coord].at(node_idx - 1).type == FullDeriv) {
            KnotSet(s[node_idx], v0, v1);
            if (node_idx >= spline_stride_) KnotSet(s[node_idx - spline_stride_], v0, v1);
        } else {
            KnotSet(s[node_idx], v0, v1);
        }
    } else {
        KnotSet(s[node_idx], v0, v1);
    }
}

// end_effector_splines.cpp:860-892
void FootSpline::SetContactTimes(std::vector<KnotTime>& ct) {
    for (auto& c : ct) {
        if (c.t < 0 && std::abs(c.t) < 1e-3) c.t = 0;
        else if (c.t < 0) throw std::runtime_error("Invalid time: negative");
    }
    int ci = 0;
    for (int i = 0; i < GetNumNodes(); i++) {
        if (times_[i].type == LiftOff || times_[i].type == TouchDown) {
            times_[i].t = ct.at(ci).t;
            ci++;
        } else if (forces_[0][i].type == Empty) {
            times_[i].t = times_[i - 1].t + (ct.at(ci).t - ct.at(ci - 1).t) / 2;
        } else {
            double span = 0.2 + ct.at(ci - 1).t;
            if (ci < static_cast<int>(ct.size())) span = ct[ci].t - ct[ci - 1].t;
            times_[i].t = times_[i - 1].t + span / num_force_polys_;
        }
    }
}

NodeType FootSpline::GetNodeType(SplineType type, int coord, int node_idx) const { return Knots(type, coord).at(node_idx).type; }

// end_effector_splines.cpp:905-940
std::vector<int> FootSpline::GetMutableNodes(SplineType type, int coord) const {
    std::vector<int> out;
    const int n = GetNumNodes();
    if (type == Force) {
        for (int i = 0; i < n; i++) {
            if (forces_[coord][i].type == FullDeriv) out.push_back(i);
        }
        return out;
    }
    for (int i = 0; i < n; i++) {
        if (positions_[coord][i].type != Empty) {
            out.push_back(i);
            if (i + spline_stride_ < n && positions_[coord][i + spline_stride_].type == NoDeriv) i += spline_stride_;
        }
    }
    return out;
}

std::vector<double> FootSpline::GetTimes() const {
    std::vector<double> out;
    for (const auto& k : times_) out.push_back(k.t);
    return out;
}

// end_effector_splines.cpp:950-979
std::vector<double> FootSpline::GetSplineAsQPVec(SplineType type, int coord) const {
    const std::vector<Knot>& s = Knots(type, coord);
    std::vector<double> out;
    for (int m : GetMutableNodes(type, coord)) {
        out.push_back(KnotVar(s[m], 0));
        if (s[m].type != NoDeriv) out.push_back(s[m].v[1]);
    }
    return out;
}

// end_effector_splines.cpp:990-997
int FootSpline::GetTotalPolyVars(SplineType type, int coord) const {
    const int m = static_cast<int>(GetMutableNodes(type, coord).size());
    return type == Force ? 2 * m : m;
}

int FootSpline::GetNumContacts() const {
    int c = 0;
    for (const auto& k : times_) c += (k.type == LiftOff || k.type == TouchDown);
    return c;
}

std::vector<double> FootSpline::GetContactTimeValues() const {
    std::vector<double> out;
    for (const auto& k : times_) {
        if (k.type == LiftOff || k.type == TouchDown) out.push_back(k.t);
    }
    return out;
}

std::vector<KnotTime> FootSpline::GetContactTimes() const {
    std::vector<KnotTime> out;
    for (const auto& k : times_) {
        if (k.type == LiftOff || k.type == TouchDown) out.push_back(k);
    }
    return out;
}

// end_effector_splines.cpp:1033-1040
double FootSpline::GetNextTouchDownTime(double time) const {
    const int up = GetUpperNodeIdx(Position, 0, time);
    if (times_[up].type == TouchDown) return times_[up].t;
    return times_[GetUpperNodeIdx(Position, 0, times_[up].t + 0.001)].t;
}

// end_effector_splines.cpp:1042-1060
void FootSpline::SetToTouchdown(double time) {
    const int up = GetUpperNodeIdx(Position, 0, time);
    if (times_[up].type != TouchDown) throw std::runtime_error("Attempting to change a lift off to a touchdown node.");
    if (std::abs(times_[up].t - time) > 1e-1) {
        throw std::runtime_error("Attempting to change a touchdown node too far away from the current time.");
    }
    const int up2 = GetUpperNodeIdx(Position, 0, times_[up].t + 0.001);
    const double time2 = times_[up2].t;
    times_[up].t = time;
    for (int i = 1; i < num_force_polys_; i++) times_.at(up + i).t = i * (time2 - time) / num_force_polys_ + time;
}

// end_effector_splines.cpp:1114-1128
int FootSpline::ConvertContactNodeToSplineNode(int contact_idx) const {
    int contacts = 0;
    for (int i = 0; i < static_cast<int>(times_.size()); i++) {
        if (contacts == contact_idx && times_[i].type != Inter) return i;
        if (times_[i].type == LiftOff || times_[i].type == TouchDown) contacts++;
    }
    throw std::runtime_error("not a valid contact index.");
}

// end_effector_splines.cpp:1155-1163
double FootSpline::GetSwingTime(double time) const {
    const int lo = GetLowerNodeIdx(Position, 0, time);
    if (times_[lo].type != LiftOff) return -1;
    const int up = GetUpperNodeIdx(Position, 0, time);
    return times_[up].t - times_[lo].t;
}

// end_effector_splines.cpp:1165-1173
double FootSpline::GetFirstTDTime() const {
    for (const auto& k : times_) {
        if (k.type == TouchDown) return k.t;
    }
    return 1e30;
}

}  // namespace oracle
