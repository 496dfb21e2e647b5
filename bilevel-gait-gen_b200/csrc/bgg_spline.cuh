// bilevel-gait-gen_b200 -- contact-spline evaluation and maintenance on the device (and host).
//
// Re-design of mpc::EndEffectorSplines (mpc/spline/end_effector_splines.cpp) for one-CTA-per-instance execution:
// a foot's splines are a fixed-capacity POD (bgg::FootSpline) living in shared memory; every query is a short scan
// over <= kMaxKnots knots with no allocation, no exceptions and no pow().  Behaviour (knot lookup and clamping,
// Hermite weights, FORCE_MULT scaling of stored force derivatives, shared touch-down/lift-off position variable,
// variable indexing) follows the reference line for line where cited; parity is tested against the CPU oracle.
#pragma once
#include "bgg_types.cuh"

namespace bgg {

enum SplineKind : int { kForce = 0, kPosXY = 1, kPosZ = 2 };

BGG_HD const uint8_t* knot_types(const FootSpline& s, int kind) {
    return kind == kForce ? s.ftype : (kind == kPosXY ? s.ptype : s.ztype);
}

// end_effector_splines.cpp:1062-1084 -- last non-empty knot with time <= t (t clamped into the spline's span;
// the reference throws below start-1e-4, we clamp and let the caller's status word record it).
BGG_HD int lower_idx(const FootSpline& s, int kind, double t) {
    const uint8_t* ty = knot_types(s, kind);
    if (t < s.t[0]) t = s.t[0];
    if (t > s.t[s.n - 1]) t = s.t[s.n - 1];
    for (int i = s.n - 1; i >= 0; --i)
        if (t >= s.t[i] && ty[i] != kEmpty) return i;
    return 0;
}

// end_effector_splines.cpp:1086-1112 -- first non-empty knot with time > t, or the last knot at the very end.
BGG_HD int upper_idx(const FootSpline& s, int kind, double t) {
    const uint8_t* ty = knot_types(s, kind);
    if (t < s.t[0]) t = s.t[0];
    if (t > s.t[s.n - 1]) t = s.t[s.n - 1];
    for (int i = 0; i < s.n; ++i)
        if (t < s.t[i] && ty[i] != kEmpty) return i;
    return s.n - 1;
}

struct Hermite {   // basis weights of end_effector_splines.cpp:1179-1197 at (tau, dT)
    double x0, x1, d0, d1;
};
BGG_HD Hermite hermite(double tau, double dT) {
    const double t2 = tau * tau, t3 = t2 * tau;
    const double i1 = 1.0 / dT, i2 = 1.0 / (dT * dT), i3 = 1.0 / (dT * dT * dT);
    Hermite h;
    h.x0 = 1.0 - i2 * 3.0 * t2 + i3 * 2.0 * t3;
    h.x1 = i2 * 3.0 * t2 - i3 * 2.0 * t3;
    h.d0 = tau - i1 * 2.0 * t2 + i2 * t3;
    h.d1 = -i1 * t2 + i2 * t3;
    return h;
}

// ValueAt, end_effector_splines.cpp:169-199.  coord 0..2; kind kForce or position (xy / z picked from coord).
BGG_HD double value_at(const FootSpline& s, bool force, int coord, double t) {
    const int kind = force ? kForce : (coord == 2 ? kPosZ : kPosXY);
    const int lo = lower_idx(s, kind, t), up = upper_idx(s, kind, t);
    const double(*v)[2] = force ? s.f[coord] : s.p[coord];
    if (lo == up) return v[lo][0];
    const double dT = s.t[up] - s.t[lo];
    const double tau = t - s.t[lo];   // raw time, as the reference (:179); only the knot lookup clamps
    const double x0 = v[lo][0], x1 = v[up][0];
    double x0d = v[lo][1], x1d = v[up][1];
    if (force) {
        x0d *= kForceMult;
        x1d *= kForceMult;
    }
    const double a2 = -(1.0 / (dT * dT)) * 3.0 * (x0 - x1) - (1.0 / dT) * (2.0 * x0d + x1d);
    const double a3 = (1.0 / (dT * dT * dT)) * 2.0 * (x0 - x1) + (1.0 / (dT * dT)) * (x0d + x1d);
    return x0 + x0d * tau + a2 * (tau * tau) + a3 * (tau * tau * tau);
}

// IsForceMutable, end_effector_splines.cpp:356-364
BGG_HD bool is_force_mutable(const FootSpline& s, double t) {
    const int lo = lower_idx(s, kForce, t), up = upper_idx(s, kForce, t);
    return !(s.ftype[lo] == kNoDeriv && s.ftype[up] == kNoDeriv);
}

// Force-spline linearisation at t: weights of the active segment's free coefficients (GetPolyVarsLin :216-243) and
// where they sit inside this foot/coord's variable block (GetVarsIdx :293-313).  Returns the count (0 if immutable).
BGG_HD int force_lin(const FootSpline& s, double t, double w[4], int* off) {
    const int lo = lower_idx(s, kForce, t), up = upper_idx(s, kForce, t);
    int nfd = 0;   // FullDeriv knots strictly before `lo`
    for (int i = 0; i < lo; ++i) nfd += (s.ftype[i] == kFullDeriv);
    *off = 2 * nfd;
    const uint8_t tl = s.ftype[lo], tu = s.ftype[up];
    if (tl == kNoDeriv && tu == kNoDeriv) return 0;
    if (lo == up) {
        w[0] = 1.0;
        return 1;
    }
    const Hermite h = hermite(t - s.t[lo], s.t[up] - s.t[lo]);
    if (tl == kNoDeriv && tu == kFullDeriv) {
        w[0] = h.x1;
        w[1] = h.d1 * kForceMult;
        return 2;
    }
    if (tl == kFullDeriv && tu == kNoDeriv) {
        w[0] = h.x0;
        w[1] = h.d0 * kForceMult;
        return 2;
    }
    w[0] = h.x0;
    w[1] = h.d0 * kForceMult;
    w[2] = h.x1;
    w[3] = h.d1 * kForceMult;
    return 4;
}

// Is knot i a "mutable" xy-position knot (GetMutableNodes, end_effector_splines.cpp:915-923): every non-empty knot,
// except that a stance's lift-off shares its variable with the touch-down `stride` knots earlier.
BGG_HD int pos_mutable_count_upto(const FootSpline& s, int upto_inclusive) {
    int cnt = 0;
    for (int i = 0; i < s.n; ++i) {
        if (s.ptype[i] != kEmpty) {
            if (i <= upto_inclusive) cnt++;
            if (i + kNumForcePolys < s.n && s.ptype[i + kNumForcePolys] == kNoDeriv) i += kNumForcePolys;
        }
    }
    return cnt;
}

// xy-position linearisation at t (GetPolyVarsLin :245-257, GetVarsIdx :315-332).  Returns count (1 or 2).
BGG_HD int pos_lin(const FootSpline& s, double t, double w[2], int* off) {
    const int lo = lower_idx(s, kPosXY, t), up = upper_idx(s, kPosXY, t);
    *off = pos_mutable_count_upto(s, lo) - 1;
    if (lo == up) {
        w[0] = 1.0;
        return 1;
    }
    const bool swing = (s.ftype[lo] == kNoDeriv) && (lo + 2 < s.n) && (s.ftype[lo + 2] == kNoDeriv);
    if (!swing) {
        w[0] = 1.0;
        return 1;
    }
    const Hermite h = hermite(t - s.t[lo], s.t[up] - s.t[lo]);
    w[0] = h.x0;
    w[1] = h.x1;
    return 2;
}

BGG_HD int num_force_vars(const FootSpline& s) {   // GetTotalPolyVars(Force, coord), :990-997
    int c = 0;
    for (int i = 0; i < s.n; ++i) c += (s.ftype[i] == kFullDeriv);
    return 2 * c;
}
BGG_HD int num_pos_vars(const FootSpline& s) { return pos_mutable_count_upto(s, s.n - 1); }

// SetVars for a z / xy position knot with the touch-down <-> lift-off pairing of end_effector_splines.cpp:828-853.
BGG_HD void set_pos_knot(FootSpline& s, int coord, int i, double v0, double v1) {
    const uint8_t* ty = (coord == 2) ? s.ztype : s.ptype;
    auto put = [&](int k) {
        s.p[coord][k][0] = v0;
        if (ty[k] != kNoDeriv) s.p[coord][k][1] = v1;
    };
    const bool not_fd = (coord != 2) || ty[i] != kFullDeriv;
    if (i < s.n - 1 && not_fd && s.ftype[i + 1] == kFullDeriv) {
        put(i);
        if (i + kNumForcePolys < s.n) put(i + kNumForcePolys);
    } else if (i > 0 && not_fd && s.ftype[i - 1] == kFullDeriv) {
        put(i);
        if (i >= kNumForcePolys) put(i - kNumForcePolys);
    } else {
        put(i);
    }
}

// Trajectory::SetSwingPosZ, trajectory.cpp:303-317
BGG_HD void set_swing_pos_z(FootSpline& s, double swing_height, double foot_offset) {
    for (int i = 0; i < s.n; ++i) {
        if (s.ztype[i] != kEmpty) {
            if (s.ztype[i] == kFullDeriv) set_pos_knot(s, 2, i, swing_height, 0.0);
            else set_pos_knot(s, 2, i, foot_offset, 0.0);
            if (i + kNumForcePolys < s.n && s.ztype[i + kNumForcePolys] == kNoDeriv) i += kNumForcePolys;
        }
    }
}

// Trajectory::UpdateForceSpline / UpdatePositionSpline (trajectory.cpp:83-111): write a solved variable block back.
BGG_HD void set_force_vars(FootSpline& s, int coord, const double* vars) {
    int k = 0;
    for (int i = 0; i < s.n; ++i)
        if (s.ftype[i] == kFullDeriv) {
            s.f[coord][i][0] = vars[k];
            s.f[coord][i][1] = vars[k + 1];
            k += 2;
        }
}
BGG_HD void set_pos_vars(FootSpline& s, int coord, const double* vars) {
    int k = 0;
    for (int i = 0; i < s.n; ++i) {
        if (s.ptype[i] != kEmpty) {
            set_pos_knot(s, coord, i, vars[k], 0.0);
            k++;
            if (i + kNumForcePolys < s.n && s.ptype[i + kNumForcePolys] == kNoDeriv) i += kNumForcePolys;
        }
    }
}
// GetSplineAsQPVec (:950-979) for the force / xy-position blocks
BGG_HD int get_force_vars(const FootSpline& s, int coord, double* out) {
    int k = 0;
    for (int i = 0; i < s.n; ++i)
        if (s.ftype[i] == kFullDeriv) {
            out[k++] = s.f[coord][i][0];
            out[k++] = s.f[coord][i][1];
        }
    return k;
}
BGG_HD int get_pos_vars(const FootSpline& s, int coord, double* out) {
    int k = 0;
    for (int i = 0; i < s.n; ++i) {
        if (s.ptype[i] != kEmpty) {
            out[k++] = s.p[coord][i][0];
            if (i + kNumForcePolys < s.n && s.ptype[i + kNumForcePolys] == kNoDeriv) i += kNumForcePolys;
        }
    }
    return k;
}

BGG_HD void push_knot(FootSpline& s, uint8_t tt, uint8_t ft, uint8_t pt, uint8_t zt, double time) {
    const int i = s.n;
    if (i >= kMaxKnots) return;
    s.ttype[i] = tt;
    s.ftype[i] = ft;
    s.ptype[i] = pt;
    s.ztype[i] = zt;
    s.t[i] = time;
    for (int c = 0; c < 3; ++c) {
        s.f[c][i][0] = s.f[c][i][1] = 0.0;
        s.p[c][i][0] = s.p[c][i][1] = 0.0;
    }
    s.n = i + 1;
}

// AddPoly, end_effector_splines.cpp:366-449
BGG_HD void add_poly(FootSpline& s, double extra) {
    const int n = s.n;
    if (s.ftype[n - 1] == kNoDeriv && s.ftype[n - 2] == kFullDeriv) {   // ends on a lift-off: append a swing
        push_knot(s, kInter, kEmpty, kEmpty, kFullDeriv, s.t[s.n - 1] + extra / 2);
        push_knot(s, kTouchDown, kNoDeriv, kNoDeriv, kNoDeriv, s.t[s.n - 1] + extra / 2);
    } else {                                                            // ends on a touch-down: append a stance
        for (int i = 0; i < kNumForcePolys - 1; ++i)
            push_knot(s, kInter, kFullDeriv, kEmpty, kEmpty, s.t[s.n - 1] + extra / kNumForcePolys);
        push_knot(s, kLiftOff, kNoDeriv, kNoDeriv, kNoDeriv, s.t[s.n - 1] + extra / kNumForcePolys);
    }
}

// Trajectory::AddPolys for one foot, trajectory.cpp:225-238
BGG_HD void add_polys_until(FootSpline& s, double final_time) {
    while (s.t[s.n - 1] < final_time && s.n + kNumForcePolys <= kMaxKnots) {
        double last = 0, prev = 0;
        int seen = 0;
        for (int i = s.n - 1; i >= 0 && seen < 2; --i)
            if (s.ttype[i] != kInter) {
                if (seen == 0) last = s.t[i];
                else prev = s.t[i];
                seen++;
            }
        const double d = last - prev;
        add_poly(s, d > 0.2 ? d : 0.2);
    }
}

// RemovePoly, end_effector_splines.cpp:451-465
BGG_HD void remove_poly(FootSpline& s, double start_time) {
    const int lo = lower_idx(s, kPosXY, start_time);
    if (lo == 0) return;
    const int n = s.n - lo;
    for (int i = 0; i < n; ++i) {
        s.ttype[i] = s.ttype[i + lo];
        s.ftype[i] = s.ftype[i + lo];
        s.ptype[i] = s.ptype[i + lo];
        s.ztype[i] = s.ztype[i + lo];
        s.t[i] = s.t[i + lo];
        for (int c = 0; c < 3; ++c) {
            s.f[c][i][0] = s.f[c][i + lo][0];
            s.f[c][i][1] = s.f[c][i + lo][1];
            s.p[c][i][0] = s.p[c][i + lo][0];
            s.p[c][i][1] = s.p[c][i + lo][1];
        }
    }
    s.n = n;
}

// GetNextTouchDownTime, :1033-1040
BGG_HD double next_touchdown_time(const FootSpline& s, double t) {
    const int up = upper_idx(s, kPosXY, t);
    if (s.ttype[up] == kTouchDown) return s.t[up];
    return s.t[upper_idx(s, kPosXY, s.t[up] + 0.001)];
}
// GetSwingTime, :1155-1163
BGG_HD double swing_time(const FootSpline& s, double t) {
    const int lo = lower_idx(s, kPosXY, t);
    if (s.ttype[lo] != kLiftOff) return -1.0;
    return s.t[upper_idx(s, kPosXY, t)] - s.t[lo];
}
// IsInContact, :805-813
BGG_HD bool is_in_contact(const FootSpline& s, double t) {
    return s.ttype[lower_idx(s, kPosXY, t)] == kTouchDown && s.ttype[upper_idx(s, kPosXY, t)] == kLiftOff;
}

// SetContactTimes, :860-892 (ct holds one time per LiftOff/TouchDown knot, in order)
BGG_HD void set_contact_times(FootSpline& s, const double* ct, int nct) {
    int ci = 0;
    for (int i = 0; i < s.n; ++i) {
        if (s.ttype[i] != kInter) {
            double v = ct[ci];
            if (v < 0 && v > -1e-3) v = 0;
            s.t[i] = v;
            ci++;
        } else if (s.ftype[i] == kEmpty) {
            s.t[i] = s.t[i - 1] + (ct[ci] - ct[ci - 1]) / 2;
        } else {
            double span = 0.2 + ct[ci - 1];
            if (ci < nct) span = ct[ci] - ct[ci - 1];
            s.t[i] = s.t[i - 1] + span / kNumForcePolys;
        }
    }
}

// Default construction of one foot (EndEffectorSplines ctor :34-153 with num_force_polys = 3) from its contact times.
BGG_HD void init_foot(FootSpline& s, const double* times, int num_contacts, bool start_in_contact) {
    s.n = 0;
    // pattern of length 5 (3 force polys): see end_effector_splines.cpp:54-100
    int i = 0, j = 0, k = 1;
    while (i < num_contacts) {
        const int q = j % 5;
        uint8_t ft, pt, zt, tt;
        if (!start_in_contact) {
            if (q == 0) { ft = kNoDeriv; pt = kNoDeriv; zt = kNoDeriv; tt = kLiftOff; }
            else if (q == 1) { ft = kEmpty; pt = kEmpty; zt = kFullDeriv; tt = kInter; }
            else if (q == 2) { ft = kNoDeriv; pt = kNoDeriv; zt = kNoDeriv; tt = kTouchDown; }
            else { ft = kFullDeriv; pt = kEmpty; zt = kEmpty; tt = kInter; }
        } else {
            if (q == 0) { ft = kNoDeriv; pt = kNoDeriv; zt = kNoDeriv; tt = kTouchDown; }
            else if (q < 3) { ft = kFullDeriv; pt = kEmpty; zt = kEmpty; tt = kInter; }
            else if (q == 4) { ft = kEmpty; pt = kEmpty; zt = kFullDeriv; tt = kInter; }
            else { ft = kNoDeriv; pt = kNoDeriv; zt = kNoDeriv; tt = kLiftOff; }
        }
        double time;
        if (ft == kFullDeriv) {
            time = times[i - 1] + k * (times[i] - times[i - 1]) / kNumForcePolys;
            k++;
        } else if (ft == kEmpty) {
            time = times[i - 1] + (times[i] - times[i - 1]) / 2;
        } else {
            time = times[i];
            i++;
            k = 1;
        }
        push_knot(s, tt, ft, pt, zt, time);
        j++;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Partials with respect to the contact times (the gait optimiser's parameters).  A foot's contact times are its
// LiftOff / TouchDown knots in order; interior force knots move proportionally with the two contact times of their
// stance, the mid-swing knot with the two of its swing (SetContactTimes, :860-892).
BGG_HD int num_contacts(const FootSpline& s) {   // GetNumContacts, :999-1008
    int c = 0;
    for (int i = 0; i < s.n; ++i) c += (s.ttype[i] != kInter);
    return c;
}
BGG_HD int contact_to_knot(const FootSpline& s, int contact_idx) {   // ConvertContactNodeToSplineNode, :1114-1128
    int c = 0;
    for (int i = 0; i < s.n; ++i) {
        if (c == contact_idx && s.ttype[i] != kInter) return i;
        if (s.ttype[i] != kInter) c++;
    }
    return -1;
}
BGG_HD int force_chain_back(const FootSpline& s, int lo) {   // FullDeriv knots walked back from `lo` (:602-609, 686-693)
    int j = 0;
    while (lo - j >= 0 && s.ftype[lo - j] == kFullDeriv) j++;
    return j;
}
struct HermiteD {   // d(basis weight)/d(theta) given d(tau)/d(theta) and d(DeltaT)/d(theta), :1199-1244
    double x0, x1, d0, d1;
};
BGG_HD HermiteD hermite_d(double tau, double dT, double dtau, double ddT) {
    const double i1 = 1.0 / dT, i2 = i1 * i1, i3 = i2 * i1, i4 = i2 * i2, t2 = tau * tau, t3 = t2 * tau;
    HermiteD h;
    h.x0 = (6 * i3 * t2 - 6 * i4 * t3) * ddT + (-6 * i2 * tau + 6 * i3 * t2) * dtau;
    h.x1 = (-6 * i3 * t2 + 6 * i4 * t3) * ddT + (6 * i2 * tau - 6 * i3 * t2) * dtau;
    h.d0 = (2 * i2 * t2 - 2 * i3 * t3) * ddT + (1 - i1 * 4 * tau + i2 * 3 * t2) * dtau;
    h.d1 = (i2 * t2 - 2 * i3 * t3) * ddT + (-i1 * 2 * tau + i2 * 3 * t2) * dtau;
    return h;
}

// ComputePartialWrtTime, :513-648: d(value at `time`)/d(contact time `time_idx`).  Position: coord 0 / 1 only.
BGG_HD double partial_wrt_time(const FootSpline& s, bool force, int coord, double time, int time_idx) {
    const int kind = force ? kForce : kPosXY;
    const int up = upper_idx(s, kind, time), lo = lower_idx(s, kind, time);
    const double dT = s.t[up] - s.t[lo], tau = time - s.t[lo];
    const int node = contact_to_knot(s, time_idx);
    const uint8_t* ty = knot_types(s, kind);
    const double(*v)[2] = force ? s.f[coord] : s.p[coord];
    const bool direct = (node == lo || node == up), wrt_lower = (node == lo);
    const double Pn = static_cast<double>(kNumForcePolys);
    const double x0 = v[lo][0], x1 = v[up][0];
    double x0d = 0, x1d = 0;
    if (ty[lo] == kFullDeriv) x0d = force ? v[lo][1] * kForceMult : v[lo][1];
    if (ty[up] == kFullDeriv) x1d = force ? v[up][1] * kForceMult : v[up][1];
    const double i1 = 1.0 / dT, i2 = i1 * i1, i3 = i2 * i1, i4 = i2 * i2;
    const double a2 = -i2 * (3 * (x0 - x1) + dT * (2 * x0d + x1d));
    const double a3 = i3 * (2 * (x0 - x1) + dT * (x0d + x1d));
    const double t2 = tau * tau, t3 = t2 * tau;
    const double vel = x0d + a2 * 2 * tau + a3 * 3 * t2;
    double ddT, dtau;
    if (direct && wrt_lower) {
        ddT = force ? -1.0 / Pn : -1.0;
        dtau = -1.0;
    } else if (direct) {
        ddT = force ? 1.0 / Pn : 1.0;
        dtau = force ? -static_cast<double>(kNumForcePolys - 1) / Pn : 0.0;
    } else if (node > up && node <= upper_idx(s, kPosXY, time)) {
        ddT = 1.0 / Pn;
        dtau = -static_cast<double>(force_chain_back(s, lo)) / Pn;
    } else if (node < lo && node >= lower_idx(s, kPosXY, time)) {
        ddT = -1.0 / Pn;
        dtau = -(static_cast<double>(-force_chain_back(s, lo)) / Pn + 1.0);
    } else {
        return 0.0;
    }
    const double da2 = 6 * i3 * (x0 - x1) * ddT + (2 * x0d + x1d) * i2 * ddT;
    const double da3 = -6 * i4 * (x0 - x1) * ddT - 2 * i3 * (x0d + x1d) * ddT;
    return da2 * t2 + da3 * t3 + vel * dtau;
}

// ComputeCoefPartialWrtTime(Force, ...), :655-763: partials of the force_lin weights; same for x, y, z.  dtwdth is the
// sensitivity of the query time itself (the constraint samples move with their stance).  Returns the weight count.
BGG_HD int force_coef_partial(const FootSpline& s, double time, int time_idx, double dtwdth, double out[4]) {
    const int up = upper_idx(s, kForce, time), lo = lower_idx(s, kForce, time);
    double dT = s.t[up] - s.t[lo];
    if (dT == 0) dT = s.t[up] - s.t[lower_idx(s, kForce, time - 1e-4)];
    const double tau = time - s.t[lo];
    const uint8_t tl = s.ftype[lo], tu = s.ftype[up];
    int cnt;
    if ((tl == kNoDeriv && tu == kFullDeriv) || (tl == kFullDeriv && tu == kNoDeriv)) cnt = 2;
    else if (lo == up) cnt = 1;
    else cnt = 4;
    for (int i = 0; i < 4; ++i) out[i] = 0.0;
    if (tl == kNoDeriv && tu == kNoDeriv) return 0;
    const int node = contact_to_knot(s, time_idx);
    const bool direct = (node == lo || node == up);
    bool wrt_lower = (node == lo);
    const double Pn = static_cast<double>(kNumForcePolys);
    const int j = force_chain_back(s, lo);
    double ddT = 1.0 / Pn, dtau = dtwdth;
    if (wrt_lower) {
        ddT = -1.0 / Pn;
        dtau += static_cast<double>(j) / Pn - 1.0;
    } else {
        dtau += -static_cast<double>(j) / Pn;
    }
    bool fill = false;
    if (direct) {
        const HermiteD h = hermite_d(tau, dT, dtau, ddT);
        if (wrt_lower) {
            out[0] = h.x1;
            out[1] = h.d1 * kForceMult;
        } else {
            out[0] = h.x0;
            out[1] = h.d0 * kForceMult;
        }
    } else if (node > up && node <= upper_idx(s, kPosXY, time)) {
        ddT = 1.0 / Pn;
        dtau = dtwdth - static_cast<double>(j) / Pn;
        fill = true;
    } else if (node < lo && node >= lower_idx(s, kPosXY, time)) {
        ddT = -1.0 / Pn;
        dtau = dtwdth + static_cast<double>(j) / Pn - 1.0;
        fill = true;
    }
    if (fill) {
        const HermiteD h = hermite_d(tau, dT, dtau, ddT);
        if (tl == kFullDeriv) {
            out[0] = h.x0;
            out[1] = kForceMult * h.d0;
            if (tu == kFullDeriv) {
                out[2] = h.x1;
                out[3] = kForceMult * h.d1;
            }
        } else if (tu == kFullDeriv) {
            out[0] = h.x1;
            out[1] = kForceMult * h.d1;
        }
    }
    return cnt;
}

// ComputeCoefPartialWrtTime(Position, x / y, ...), :764-803: partials of the pos_lin weights.  Returns the count.
BGG_HD int pos_coef_partial(const FootSpline& s, double time, int time_idx, double out[2]) {
    const int up = upper_idx(s, kPosXY, time), lo = lower_idx(s, kPosXY, time);
    double dT = s.t[up] - s.t[lo];
    if (dT == 0) dT = s.t[up] - s.t[lower_idx(s, kPosXY, time - 1e-4)];
    const double tau = time - s.t[lo];
    int cnt = 1;
    if (lo != up && s.ftype[lo] == kNoDeriv && lo + 2 < s.n && s.ftype[lo + 2] == kNoDeriv) cnt = 2;
    out[0] = out[1] = 0.0;
    const int node = contact_to_knot(s, time_idx);
    const bool direct = (node == lo || node == up), wrt_lower = (node == lo);
    double ddT = 1.0, dtau = 0.0;
    if (wrt_lower) {
        dtau = -1.0;
        ddT = -1.0;
    }
    if (direct && lo != up && lo + 1 < s.n && s.ztype[lo + 1] == kFullDeriv) {
        const HermiteD h = hermite_d(tau, dT, dtau, ddT);
        out[0] = h.x0;
        if (cnt > 1) out[1] = h.x1;
    }
    return cnt;
}

}  // namespace bgg
