// bilevel-gait-gen_b200 -- robot constants from a URDF without pinocchio.
//
// The reference obtains these from pinocchio when MPCSingleRigidBody is constructed:
//   total mass                                   mpc/models/model.cpp:27 (pinocchio::computeTotalMass)
//   composite rotational inertia about the CoM at the nominal configuration, in the floating-base frame
//                                                mpc/models/single_rigid_body_model.cpp:33-37
//   root -> hip joint translations with the hard-coded +0.025 x / +-0.1 y shifts   :258-308 (GetCOMToHip)
// Here: a minimal URDF reader (links with <inertial>, joints with parent / child / origin / axis), fixed joints merged
// into their parent, revolute joints rotated by the nominal joint angle, children visited in alphabetical order of the
// joint name (pinocchio's ordering: FL, FR, RL, RR for the A1).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <cstdlib>

#include "mpc_b200.h"

namespace mpc {
namespace {

struct M3 {
    double a[3][3];
};
M3 Ident() { return {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
M3 Mul(const M3& x, const M3& y) {
    M3 r{};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) r.a[i][j] += x.a[i][k] * y.a[k][j];
    return r;
}
M3 Tr(const M3& x) {
    M3 r{};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.a[i][j] = x.a[j][i];
    return r;
}
void MulV(const M3& x, const double v[3], double o[3]) {
    for (int i = 0; i < 3; ++i) o[i] = x.a[i][0] * v[0] + x.a[i][1] * v[1] + x.a[i][2] * v[2];
}
M3 Rpy(double r, double p, double y) {
    const double cr = std::cos(r), sr = std::sin(r), cp = std::cos(p), sp = std::sin(p), cy = std::cos(y), sy = std::sin(y);
    const M3 Rx = {{{1, 0, 0}, {0, cr, -sr}, {0, sr, cr}}}, Ry = {{{cp, 0, sp}, {0, 1, 0}, {-sp, 0, cp}}}, Rz = {{{cy, -sy, 0}, {sy, cy, 0}, {0, 0, 1}}};
    return Mul(Rz, Mul(Ry, Rx));
}
M3 AxisAngle(const double ax[3], double q) {
    const double n = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    const double a[3] = {ax[0] / n, ax[1] / n, ax[2] / n};
    const M3 K = {{{0, -a[2], a[1]}, {a[2], 0, -a[0]}, {-a[1], a[0], 0}}};
    const M3 K2 = Mul(K, K);
    M3 r = Ident();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.a[i][j] += std::sin(q) * K.a[i][j] + (1 - std::cos(q)) * K2.a[i][j];
    return r;
}

struct Tag {
    std::string name;
    std::map<std::string, std::string> attr;
    bool closing = false, self_closing = false;
};
// next tag at or after pos; returns false at the end.  Comments and the XML prolog are skipped.
bool NextTag(const std::string& s, size_t& pos, Tag& t) {
    while (true) {
        const size_t lt = s.find('<', pos);
        if (lt == std::string::npos) return false;
        if (s.compare(lt, 4, "<!--") == 0) {
            pos = s.find("-->", lt);
            if (pos == std::string::npos) return false;
            pos += 3;
            continue;
        }
        const size_t gt = s.find('>', lt);
        if (gt == std::string::npos) return false;
        pos = gt + 1;
        if (s[lt + 1] == '?' || s[lt + 1] == '!') continue;
        std::string body = s.substr(lt + 1, gt - lt - 1);
        t = Tag();
        if (!body.empty() && body[0] == '/') {
            t.closing = true;
            body = body.substr(1);
        }
        if (!body.empty() && body.back() == '/') {
            t.self_closing = true;
            body.pop_back();
        }
        size_t i = 0;
        while (i < body.size() && !std::isspace(static_cast<unsigned char>(body[i]))) ++i;
        t.name = body.substr(0, i);
        while (i < body.size()) {
            while (i < body.size() && std::isspace(static_cast<unsigned char>(body[i]))) ++i;
            const size_t eq = body.find('=', i);
            if (eq == std::string::npos) break;
            std::string key = body.substr(i, eq - i);
            while (!key.empty() && std::isspace(static_cast<unsigned char>(key.back()))) key.pop_back();
            size_t q0 = body.find_first_of("\"'", eq);
            if (q0 == std::string::npos) break;
            const size_t q1 = body.find(body[q0], q0 + 1);
            if (q1 == std::string::npos) break;
            t.attr[key] = body.substr(q0 + 1, q1 - q0 - 1);
            i = q1 + 1;
        }
        return true;
    }
}
void Vec3(const std::string& s, double v[3]) {   // C stdio / strtod only: no iostreams in this translation unit
    v[0] = v[1] = v[2] = 0;
    const char* p = s.c_str();
    for (int i = 0; i < 3; ++i) {
        char* end = nullptr;
        const double x = std::strtod(p, &end);
        if (end == p) break;
        v[i] = x;
        p = end;
    }
}
double Num(const std::string& s) {
    char* end = nullptr;
    const double x = std::strtod(s.c_str(), &end);
    if (end == s.c_str()) throw std::runtime_error("URDF: not a number: '" + s + "'");
    return x;
}

struct Link {
    bool has_inertia = false;
    double m = 0, c[3] = {0, 0, 0};
    M3 I{};
};
struct Joint {
    std::string name, type, parent, child;
    double xyz[3] = {0, 0, 0};
    M3 R = Ident();
    double axis[3] = {1, 0, 0};
};
struct Body {
    double m, c[3];
    M3 I;
};

}  // namespace

namespace {
void ParseURDF(const std::string& urdf_path, std::map<std::string, Link>& links, std::vector<Joint>& joints) {
    std::FILE* f = std::fopen(urdf_path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open URDF " + urdf_path);
    std::string s;
    char buf[65536];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, got);
    std::fclose(f);
    size_t pos = 0;
    Tag t;
    std::vector<std::string> stack;
    std::string cur_link;
    Joint cur_joint;
    bool in_joint = false, in_inertial = false, joint_has_parent = false;
    double org_xyz[3] = {0, 0, 0};
    M3 org_R = Ident();
    Link* L = nullptr;
    while (NextTag(s, pos, t)) {
        if (t.closing) {
            if (t.name == "link" && stack.size() == 2) cur_link.clear();
            if (t.name == "inertial") in_inertial = false;
            if (t.name == "joint" && in_joint && stack.size() == 2) {
                if (joint_has_parent) joints.push_back(cur_joint);   // <transmission><joint/> stubs have no parent
                in_joint = false;
            }
            if (!stack.empty()) stack.pop_back();
            continue;
        }
        const size_t depth = stack.size();
        if (t.name == "link" && depth == 1) {
            cur_link = t.attr["name"];
            links[cur_link];
            L = &links[cur_link];
        } else if (t.name == "joint" && depth == 1) {
            cur_joint = Joint();
            cur_joint.name = t.attr["name"];
            cur_joint.type = t.attr["type"];
            in_joint = true;
            joint_has_parent = false;
        } else if (t.name == "inertial" && !cur_link.empty()) {
            in_inertial = true;
            L->has_inertia = true;
            L->I = M3{};
            org_xyz[0] = org_xyz[1] = org_xyz[2] = 0;
            org_R = Ident();
        } else if (t.name == "origin") {
            double xyz[3], rpy[3];
            Vec3(t.attr.count("xyz") ? t.attr["xyz"] : "0 0 0", xyz);
            Vec3(t.attr.count("rpy") ? t.attr["rpy"] : "0 0 0", rpy);
            if (in_inertial) {
                std::memcpy(L->c, xyz, sizeof xyz);
                org_R = Rpy(rpy[0], rpy[1], rpy[2]);
            } else if (in_joint && depth == 2) {
                std::memcpy(cur_joint.xyz, xyz, sizeof xyz);
                cur_joint.R = Rpy(rpy[0], rpy[1], rpy[2]);
            }
        } else if (t.name == "mass" && in_inertial) {
            L->m = Num(t.attr["value"]);
        } else if (t.name == "inertia" && in_inertial) {
            const double ixx = Num(t.attr["ixx"]), ixy = Num(t.attr["ixy"]), ixz = Num(t.attr["ixz"]),
                         iyy = Num(t.attr["iyy"]), iyz = Num(t.attr["iyz"]), izz = Num(t.attr["izz"]);
            const M3 I = {{{ixx, ixy, ixz}, {ixy, iyy, iyz}, {ixz, iyz, izz}}};
            L->I = Mul(org_R, Mul(I, Tr(org_R)));
        } else if (in_joint && depth == 2 && t.name == "parent") {
            cur_joint.parent = t.attr["link"];
            joint_has_parent = true;
        } else if (in_joint && depth == 2 && t.name == "child") {
            cur_joint.child = t.attr["link"];
        } else if (in_joint && depth == 2 && t.name == "axis") {
            Vec3(t.attr["xyz"], cur_joint.axis);
        }
        if (!t.self_closing) stack.push_back(t.name);
    }
}
}  // namespace

bgg_robot RobotConstsFromURDF(const std::string& urdf_path, const std::map<std::string, double>& joint_cfg) {
    std::map<std::string, Link> links;
    std::vector<Joint> joints;
    ParseURDF(urdf_path, links, joints);
    // the inertial origin's rotation may come after <inertia> in the file: the A1 URDF lists origin first; nothing to fix up
    std::map<std::string, std::vector<const Joint*>> children;
    std::map<std::string, bool> is_child;
    for (const Joint& j : joints) {
        children[j.parent].push_back(&j);
        is_child[j.child] = true;
    }
    std::string root;
    for (const auto& kv : links)
        if (!is_child.count(kv.first)) root = kv.first;
    std::vector<Body> bodies;
    std::map<std::string, std::array<double, 3>> joint_pos;
    struct Frame {
        std::string link;
        double p[3];
        M3 R;
    };
    std::vector<Frame> todo;
    todo.push_back({root, {0, 0, 0}, Ident()});
    while (!todo.empty()) {
        const Frame fr = todo.back();
        todo.pop_back();
        const Link& l = links[fr.link];
        if (l.has_inertia) {
            Body b;
            b.m = l.m;
            double rc[3];
            MulV(fr.R, l.c, rc);
            for (int i = 0; i < 3; ++i) b.c[i] = fr.p[i] + rc[i];
            b.I = Mul(fr.R, Mul(l.I, Tr(fr.R)));
            bodies.push_back(b);
        }
        std::vector<const Joint*> ch = children[fr.link];
        std::sort(ch.begin(), ch.end(), [](const Joint* a, const Joint* b) { return a->name < b->name; });
        for (const Joint* j : ch) {
            Frame nf;
            nf.link = j->child;
            double rx[3];
            MulV(fr.R, j->xyz, rx);
            for (int i = 0; i < 3; ++i) nf.p[i] = fr.p[i] + rx[i];
            nf.R = Mul(fr.R, j->R);
            if (j->type == "revolute" || j->type == "continuous") {
                const auto it = joint_cfg.find(j->name);
                nf.R = Mul(nf.R, AxisAngle(j->axis, it == joint_cfg.end() ? 0.0 : it->second));
            }
            joint_pos[j->name] = {nf.p[0], nf.p[1], nf.p[2]};
            todo.push_back(nf);
        }
    }
    bgg_robot rb{};
    double com[3] = {0, 0, 0};
    for (const Body& b : bodies) {
        rb.mass += b.m;
        for (int i = 0; i < 3; ++i) com[i] += b.m * b.c[i];
    }
    for (double& c : com) c /= rb.mass;
    double Ir[3][3] = {};
    for (const Body& b : bodies) {
        const double d[3] = {b.c[0] - com[0], b.c[1] - com[1], b.c[2] - com[2]};
        const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Ir[i][j] += b.I.a[i][j] + b.m * ((i == j ? dd : 0.0) - d[i] * d[j]);
    }
    // Ir_.inverse() as Eigen evaluates it for a fixed 3 x 3 matrix (Eigen/src/LU/InverseImpl.h): cofactors times 1 / det,
    // the determinant expanded along the first column (single_rigid_body_model.cpp:37)
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return Ir[i1][j1] * Ir[i2][j2] - Ir[i1][j2] * Ir[i2][j1];
    };
    const double c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const double det = (c0 * Ir[0][0] + c1 * Ir[1][0]) + c2 * Ir[2][0];
    const double invdet = 1.0 / det;
    double inv[3][3];
    inv[0][0] = c0 * invdet; inv[0][1] = c1 * invdet; inv[0][2] = c2 * invdet;
    inv[1][0] = cof(0, 1) * invdet; inv[1][1] = cof(1, 1) * invdet; inv[2][2] = cof(2, 2) * invdet;
    inv[1][2] = cof(2, 1) * invdet; inv[2][1] = cof(1, 2) * invdet; inv[2][0] = cof(0, 2) * invdet;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            rb.Ir[3 * i + j] = Ir[i][j];
            rb.Ir_inv[3 * i + j] = inv[i][j];
        }
    const char* hips[4] = {"FL_hip_joint", "FR_hip_joint", "RL_hip_joint", "RR_hip_joint"};
    for (int e = 0; e < 4; ++e) {
        const auto it = joint_pos.find(hips[e]);
        if (it == joint_pos.end()) throw std::runtime_error(std::string("URDF has no joint ") + hips[e]);
        double x = it->second[0], y = it->second[1];
        y += (y >= 0) ? 0.1 : -0.1;   // single_rigid_body_model.cpp:291-297
        x += 0.025;                   // :299-305
        rb.hip_xy[2 * e] = x;
        rb.hip_xy[2 * e + 1] = y;
    }
    rb.gravity[0] = rb.gravity[1] = 0;
    rb.gravity[2] = -9.81;   // models/model.cpp:16
    return rb;
}

bgg_robot RobotConstsFromURDF(const std::string& urdf_path) {
    // apps/a1_configuration.yaml:init_config joint part, pinocchio order FL, FR, RL, RR
    std::map<std::string, double> cfg;
    const char* legs[4] = {"FL", "FR", "RL", "RR"};
    const double hip[4] = {-0.02, 0.02, 0.02, -0.02};
    for (int l = 0; l < 4; ++l) {
        cfg[std::string(legs[l]) + "_hip_joint"] = hip[l];
        cfg[std::string(legs[l]) + "_thigh_joint"] = 0.9;
        cfg[std::string(legs[l]) + "_calf_joint"] = -1.6;
    }
    return RobotConstsFromURDF(urdf_path, cfg);
}

// Leg chains for the inverse kinematics (single_rigid_body_model.cpp:314-455): every placement relative to the movable joint before it,
// fixed joints in between merged the way pinocchio's URDF parser does.
bgg_kinematics LegKinematicsFromURDF(const std::string& urdf_path) {
    std::map<std::string, Link> links;
    std::vector<Joint> joints;
    ParseURDF(urdf_path, links, joints);
    std::map<std::string, const Joint*> by_child, by_name;
    std::map<std::string, bool> is_child;
    for (const Joint& j : joints) {
        by_child[j.child] = &j;
        by_name[j.name] = &j;
        is_child[j.child] = true;
    }
    std::string root;
    for (const auto& kv : links)
        if (!is_child.count(kv.first)) root = kv.first;
    bgg_kinematics kin{};
    const char* legs[4] = {"FL", "FR", "RL", "RR"};
    const char* chain[4] = {"_hip_joint", "_thigh_joint", "_calf_joint", "_foot_fixed"};
    for (int e = 0; e < 4; ++e) {
        std::string ancestor = root;
        for (int k = 0; k < 4; ++k) {
            const std::string name = std::string(legs[e]) + chain[k];
            const auto it = by_name.find(name);
            if (it == by_name.end()) throw std::runtime_error("URDF has no joint " + name);
            const Joint* j = it->second;
            double t[3] = {j->xyz[0], j->xyz[1], j->xyz[2]};
            M3 R = j->R;
            std::string link = j->parent;
            while (link != ancestor) {
                const auto up = by_child.find(link);
                if (up == by_child.end() || up->second->type != "fixed") throw std::runtime_error("URDF: " + name + " does not hang off " + ancestor + " through fixed joints");
                double rt[3];
                MulV(up->second->R, t, rt);
                for (int i = 0; i < 3; ++i) t[i] = up->second->xyz[i] + rt[i];
                R = Mul(up->second->R, R);
                link = up->second->parent;
            }
            for (int i = 0; i < 3; ++i) kin.leg[e].t[k][i] = t[i];
            for (int i = 0; i < 3; ++i)
                for (int c = 0; c < 3; ++c) kin.leg[e].R[k][3 * i + c] = R.a[i][c];
            if (k < 3) {
                const double n = std::sqrt(j->axis[0] * j->axis[0] + j->axis[1] * j->axis[1] + j->axis[2] * j->axis[2]);
                for (int i = 0; i < 3; ++i) kin.leg[e].axis[k][i] = j->axis[i] / n;
            }
            ancestor = j->child;
        }
    }
    return kin;
}

}  // namespace mpc

// C entry point for bindings / tests: 0 on success, -1 on failure (message on stderr)
extern "C" int bgg_host_leg_kinematics_from_urdf(const char* urdf_path, bgg_kinematics* out) {
    try {
        *out = mpc::LegKinematicsFromURDF(urdf_path);
        return 0;
    } catch (...) {
        return -1;
    }
}

extern "C" int bgg_host_robot_consts_from_urdf(const char* urdf_path, bgg_robot* out) {
    try {
        *out = mpc::RobotConstsFromURDF(urdf_path);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "bgg_host_robot_consts_from_urdf: %s\n", e.what());
        return -1;
    }
}
