// TEST INFRASTRUCTURE ONLY.  C entry points over the reference's own mpc::EndEffectorSplines, compiled from
// /root/reference (see ../Makefile, target _ref/libref_splines.so).  Mirrors the orc_spline_* functions of
// ../oracle_capi.cpp one for one so tests can run both through the same driver.
#include <limits>
#include <string>

#include "spline/end_effector_splines.h"

using mpc::EndEffectorSplines;
static thread_local std::string g_err;
static const double kNaN = std::numeric_limits<double>::quiet_NaN();
#define T(h) (static_cast<EndEffectorSplines*>(h))
#define ST(t) (static_cast<EndEffectorSplines::SplineType>(t))
#define TRY try {
#define CATCH(ret) } catch (const std::exception& e) { g_err = e.what(); return ret; }

extern "C" {
const char* orc_last_error() { return g_err.c_str(); }
void orc_clear_error() { g_err.clear(); }
void* orc_spline_create(int num_contacts, const double* times, int start_in_contact, int num_force_polys) {
    TRY std::vector<double> t(times, times + num_contacts);
    return new EndEffectorSplines(num_contacts, t, start_in_contact != 0, num_force_polys);
    CATCH(nullptr)
}
void orc_spline_destroy(void* h) { delete T(h); }
void* orc_spline_clone(void* h) { return new EndEffectorSplines(*T(h)); }
double orc_spline_value(void* h, int type, int coord, double t) { TRY return T(h)->ValueAt(ST(type), coord, t); CATCH(kNaN) }
int orc_spline_lin(void* h, int type, int coord, double t, double* out) {
    TRY const mpc::vector_t v = T(h)->GetPolyVarsLin(ST(type), coord, t);
    for (int i = 0; i < v.size(); i++) out[i] = v(i);
    return v.size();
    CATCH(-1)
}
int orc_spline_vars_idx(void* h, int type, int coord, double t, int* idx, int* cnt) {
    TRY const auto p = T(h)->GetVarsIdx(ST(type), coord, t);
    *idx = p.first; *cnt = p.second; return 0;
    CATCH(-1)
}
int orc_spline_is_force_mutable(void* h, double t) { TRY return T(h)->IsForceMutable(t) ? 1 : 0; CATCH(-1) }
int orc_spline_is_in_contact(void* h, double t) { TRY return T(h)->IsInContact(t) ? 1 : 0; CATCH(-1) }
int orc_spline_add_poly(void* h, double dt) { TRY T(h)->AddPoly(dt); return 0; CATCH(-1) }
int orc_spline_remove_poly(void* h, double t) { TRY T(h)->RemovePoly(t); return 0; CATCH(-1) }
double orc_spline_partial(void* h, int type, int coord, double t, int time_idx) {
    TRY return T(h)->ComputePartialWrtTime(ST(type), coord, t, time_idx); CATCH(kNaN)
}
int orc_spline_coef_partial(void* h, int type, int coord, double t, int time_idx, double dtwdth, double* out) {
    TRY const mpc::vector_t v = T(h)->ComputeCoefPartialWrtTime(ST(type), coord, t, time_idx, dtwdth);
    for (int i = 0; i < v.size(); i++) out[i] = v(i);
    return v.size();
    CATCH(-1)
}
int orc_spline_set_vars(void* h, int type, int coord, int node, double v0, double v1) {
    TRY mpc::vector_2t v; v(0) = v0; v(1) = v1;
    T(h)->SetVars(ST(type), coord, node, v); return 0;
    CATCH(-1)
}
int orc_spline_set_contact_times(void* h, const double* t, int n) {
    TRY mpc::time_v ct = T(h)->GetContactTimes();
    if (static_cast<int>(ct.size()) != n) throw std::runtime_error("contact time count mismatch");
    for (int i = 0; i < n; i++) ct[i].SetTime(t[i]);
    T(h)->SetContactTimes(ct); return 0;
    CATCH(-1)
}
int orc_spline_num_nodes(void* h) { return T(h)->GetNumNodes(); }
int orc_spline_num_contacts(void* h) { return T(h)->GetNumContacts(); }
int orc_spline_node_type(void* h, int type, int coord, int node) { TRY return T(h)->GetNodeType(ST(type), coord, node); CATCH(-1) }
int orc_spline_mutable_nodes(void* h, int type, int coord, int* out) {
    TRY const auto v = T(h)->GetMutableNodes(ST(type), coord);
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return static_cast<int>(v.size());
    CATCH(-1)
}
int orc_spline_times(void* h, double* out, int* types) {
    const auto v = T(h)->GetTimes();
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    (void)types;   // the time types are private in the reference; callers pass NULL for the _ref library
    return static_cast<int>(v.size());
}
int orc_spline_as_qp_vec(void* h, int type, int coord, double* out) {
    TRY const mpc::vector_t v = T(h)->GetSplineAsQPVec(ST(type), coord);
    for (int i = 0; i < v.size(); i++) out[i] = v(i);
    return v.size();
    CATCH(-1)
}
int orc_spline_total_poly_vars(void* h, int type, int coord) { return T(h)->GetTotalPolyVars(ST(type), coord); }
double orc_spline_end_time(void* h) { return T(h)->GetEndTime(); }
double orc_spline_start_time(void* h) { return T(h)->GetStartTime(); }
double orc_spline_next_td(void* h, double t) { TRY return T(h)->GetNextTouchDownTime(t); CATCH(kNaN) }
double orc_spline_swing_time(void* h, double t) { TRY return T(h)->GetSwingTime(t); CATCH(kNaN) }
int orc_spline_set_to_touchdown(void* h, double t) { TRY T(h)->SetToTouchdown(t); return 0; CATCH(-1) }
}  // extern "C"
