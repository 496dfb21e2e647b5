// TEST INFRASTRUCTURE ONLY -- CPU oracle.  Nothing under oracle/ may be imported, linked or executed by the
// product path (bilevel-gait-gen_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
// --impl reference legs use it, and only as the checker / reported baseline.
//
// Restatement (no Eigen) of the reference's per-foot contact splines:
//   mpc/include/spline/end_effector_splines.h:10-157, mpc/spline/end_effector_splines.cpp (whole file),
//   mpc/include/spline/spline_node.h:14-38, mpc/spline/spline_node.cpp:9-37.
// Pinned by tests/test_oracle_splines.py against the reference's own known answers
// (test/splines_tests.cpp:66-105), its reconstruction / add-remove / finite-difference checks (:109-444) and,
// when oracle/_ref has been built, against the reference sources themselves compiled here.
#pragma once
#include <array>
#include <stdexcept>
#include <utility>
#include <vector>

namespace oracle {

enum NodeType { NoDeriv = 0, FullDeriv = 1, Empty = 2 };      // spline_node.h:14-18
enum TimeType { LiftOff = 0, TouchDown = 1, Inter = 2 };       // end_effector_splines.h:11-15
enum SplineType { Force = 0, Position = 1 };                   // end_effector_splines.h:38-41

struct Knot {            // one SplineNode: its type and (value, stored derivative)
    NodeType type;
    double v[2];
};

struct KnotTime {        // one SplineTimes entry
    double t;
    TimeType type;
};

constexpr double kForceMult = 100.0;                           // end_effector_splines.h:152

class FootSpline {
public:
    FootSpline(int num_contacts, const std::vector<double>& times, bool start_in_contact, int num_force_polys);

    double ValueAt(SplineType type, int coord, double time) const;                       // .cpp:169-199
    std::vector<double> GetPolyVarsLin(SplineType type, int coord, double time) const;   // .cpp:201-282
    std::pair<int, int> GetVarsIdx(SplineType type, int coord, double time) const;       // .cpp:284-354
    bool IsForceMutable(double time) const;                                              // .cpp:356-364
    void AddPoly(double additional_time);                                                // .cpp:366-449
    void RemovePoly(double start_time);                                                  // .cpp:451-465
    double ComputePartialWrtTime(SplineType type, int coord, double time, int time_idx) const;   // .cpp:513-648
    std::vector<double> ComputeCoefPartialWrtTime(SplineType type, int coord, double time, int time_idx,
                                                  double dtwdth = 0.0) const;            // .cpp:650-803
    bool IsInContact(double time) const;                                                 // .cpp:805-813
    void SetVars(SplineType type, int coord, int node_idx, double v0, double v1);        // .cpp:815-858
    void SetContactTimes(std::vector<KnotTime>& contact_times);                          // .cpp:860-892
    NodeType GetNodeType(SplineType type, int coord, int node_idx) const;                // .cpp:894-897
    int GetNumNodes() const { return static_cast<int>(times_.size()); }                  // .cpp:899-903
    std::vector<int> GetMutableNodes(SplineType type, int coord) const;                  // .cpp:905-940
    std::vector<double> GetTimes() const;                                                // .cpp:942-948
    std::vector<double> GetSplineAsQPVec(SplineType type, int coord) const;              // .cpp:950-979
    double GetEndTime() const { return times_.back().t; }                                // .cpp:982-984
    double GetStartTime() const { return times_.front().t; }                             // .cpp:986-988
    int GetTotalPolyVars(SplineType type, int coord) const;                              // .cpp:990-997
    int GetNumContacts() const;                                                          // .cpp:999-1008
    std::vector<double> GetContactTimeValues() const;                                    // .cpp:1010-1020
    std::vector<KnotTime> GetContactTimes() const;                                       // .cpp:1022-1031
    double GetNextTouchDownTime(double time) const;                                      // .cpp:1033-1040
    void SetToTouchdown(double time);                                                    // .cpp:1042-1060
    double GetSwingTime(double time) const;                                              // .cpp:1155-1163
    double GetFirstTDTime() const;                                                       // .cpp:1165-1173

    int GetLowerNodeIdx(SplineType type, int coord, double time) const;                  // .cpp:1062-1084
    int GetUpperNodeIdx(SplineType type, int coord, double time) const;                  // .cpp:1086-1112
    int ConvertContactNodeToSplineNode(int contact_idx) const;                           // .cpp:1114-1128

    // raw access for packing into the product's device layout in tests
    const std::vector<Knot>& Knots(SplineType type, int coord) const {
        return type == Force ? forces_[coord] : positions_[coord];
    }
    const std::vector<KnotTime>& KnotTimes() const { return times_; }
    int NumForcePolys() const { return num_force_polys_; }

private:
    std::vector<Knot>& Sel(SplineType type, int coord) { return type == Force ? forces_[coord] : positions_[coord]; }
    int ForceChainBack(int coord, int lower_node) const;

    std::array<std::vector<Knot>, 3> forces_;
    std::array<std::vector<Knot>, 3> positions_;
    std::vector<KnotTime> times_;
    int num_force_polys_;
    int spline_stride_;
};

}  // namespace oracle
