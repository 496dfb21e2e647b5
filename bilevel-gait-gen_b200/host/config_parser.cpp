// bilevel-gait-gen_b200 -- see config_parser.h.  C stdio only (no iostreams).
#include "config_parser.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>

namespace utils {

namespace {
std::string Trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
    while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
    return s.substr(a, b - a);
}
std::string Unquote(const std::string& s) {
    if (s.size() >= 2 && (s.front() == '"' || s.front() == '\'') && s.back() == s.front()) return s.substr(1, s.size() - 2);
    return s;
}
// strips a trailing comment (a '#' outside quotes)
std::string StripComment(const std::string& line) {
    char q = 0;
    for (size_t i = 0; i < line.size(); ++i) {
        const char c = line[i];
        if (q) {
            if (c == q) q = 0;
        } else if (c == '"' || c == '\'') {
            q = c;
        } else if (c == '#') {
            return line.substr(0, i);
        }
    }
    return line;
}
void SplitList(const std::string& body, std::vector<std::string>& out) {
    std::string cur;
    char q = 0;
    for (char c : body) {
        if (q) {
            cur += c;
            if (c == q) q = 0;
        } else if (c == '"' || c == '\'') {
            q = c;
            cur += c;
        } else if (c == ',') {
            if (!Trim(cur).empty()) out.push_back(Unquote(Trim(cur)));
            cur.clear();
        } else {
            cur += c;
        }
    }
    if (!Trim(cur).empty()) out.push_back(Unquote(Trim(cur)));
}
}  // namespace

ConfigParser::ConfigParser(const std::string& file_name) : file_name_(file_name) {
    std::FILE* f = std::fopen(file_name.c_str(), "rb");
    if (!f) throw std::runtime_error("bad file: " + file_name);   // YAML::BadFile in the reference
    std::string text;
    char buf[65536];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
    std::fclose(f);
    std::string key, open_list;
    bool in_list = false;
    size_t pos = 0;
    while (pos <= text.size()) {
        const size_t nl = text.find('\n', pos);
        std::string line = StripComment(text.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos));
        pos = (nl == std::string::npos) ? text.size() + 1 : nl + 1;
        if (in_list) {
            const size_t close = line.find(']');
            open_list += " " + (close == std::string::npos ? line : line.substr(0, close));
            if (close != std::string::npos) {
                SplitList(open_list, items_[key]);
                in_list = false;
            }
            continue;
        }
        if (Trim(line).empty()) continue;
        const size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        key = Trim(line.substr(0, colon));
        const std::string val = Trim(line.substr(colon + 1));
        items_[key].clear();
        if (!val.empty() && val.front() == '[') {
            const size_t close = val.find(']');
            if (close != std::string::npos) {
                SplitList(val.substr(1, close - 1), items_[key]);
            } else {
                open_list = val.substr(1);
                in_list = true;
            }
        } else if (!val.empty()) {
            items_[key].push_back(Unquote(val));
        }
    }
}

const std::vector<std::string>& ConfigParser::Items(const std::string& element) const {
    const auto it = items_.find(element);
    if (it == items_.end()) throw std::runtime_error("bad conversion: no element '" + element + "' in " + file_name_);
    return it->second;
}
static double ToNumber(const std::string& s, const std::string& element) {
    char* end = nullptr;
    const double x = std::strtod(s.c_str(), &end);
    if (end == s.c_str()) throw std::runtime_error("bad conversion: element '" + element + "' is not a number: " + s);
    return x;
}
double ConfigParser::Number(const std::string& element) const {
    const auto& v = Items(element);
    if (v.size() != 1) throw std::runtime_error("bad conversion: element '" + element + "' is not a scalar");
    return ToNumber(v[0], element);
}
void ConfigParser::Fill(const std::string& element, std::vector<double>& out) const {
    for (const auto& s : Items(element)) out.push_back(ToNumber(s, element));
}
void ConfigParser::Fill(const std::string& element, std::vector<int>& out) const {
    for (const auto& s : Items(element)) out.push_back(static_cast<int>(ToNumber(s, element)));
}
void ConfigParser::Fill(const std::string& element, std::vector<std::string>& out) const { out = Items(element); }
mpc::vector_t ConfigParser::ParseEigenVector(const std::string& element) const {
    const auto& v = Items(element);
    mpc::vector_t out(static_cast<int>(v.size()));
    for (size_t i = 0; i < v.size(); ++i) out(static_cast<int>(i)) = ToNumber(v[i], element);
    return out;
}
std::string ConfigParser::ParseString(const std::string& element) const {
    const auto& v = Items(element);
    if (v.size() != 1) throw std::runtime_error("bad conversion: element '" + element + "' is not a scalar");
    return v[0];
}
std::vector<std::string> ConfigParser::ParseStringVector(const std::string& element) const { return Items(element); }

}  // namespace utils

namespace mpc {

MPCInfo MPCInfoFromConfig(const utils::ConfigParser& config) {
    MPCInfo info;   // test/mpc_test.cpp:46-83
    info.discretization_steps = static_cast<int>(config.ParseNumber<double>("discretization_steps"));
    info.num_nodes = config.ParseNumber<int>("num_nodes");
    info.num_qp_iterations = config.ParseNumber<int>("num_qp");
    info.friction_coef = config.ParseNumber<double>("friction_coef");
    info.vel_bounds = config.ParseEigenVector("vel_bounds");
    info.joint_bounds_lb = config.ParseEigenVector("joint_bounds_lb");
    info.joint_bounds_ub = config.ParseEigenVector("joint_bounds_ub");
    info.ee_frames = config.ParseStdVector<std::string>("collision_frames");
    info.num_switches = config.ParseNumber<int>("num_switches");
    info.integrator_dt = config.ParseNumber<double>("integrator_dt");
    info.num_contacts = static_cast<int>(info.ee_frames.size());
    info.force_bound = config.ParseNumber<double>("force_bound");
    info.swing_height = config.ParseNumber<double>("swing_height");
    info.foot_offset = config.ParseNumber<double>("foot_offset");
    info.nom_state = config.ParseEigenVector("init_config");
    const vector_t box = config.ParseEigenVector("ee_box_size");
    if (box.size() != 2) throw std::runtime_error("ee_box_size must have two entries");
    info.ee_box_size = vector_2t(box(0), box(1));
    info.real_time_iters = config.ParseNumber<int>("run_time_iterations");
    info.force_cost = config.ParseNumber<double>("force_cost");
    if (config.Has("mpc_verbosity")) {
        switch (config.ParseNumber<int>("mpc_verbosity")) {
            case 0: info.verbose = Nothing; break;
            case 1: info.verbose = Timing; break;
            case 2: info.verbose = Optimization; break;
            case 3: info.verbose = All; break;
            default: throw std::runtime_error("Not a valid verbosity level for MPC.");
        }
    }
    return info;
}

}  // namespace mpc

// C entry point for bindings / tests: the MPCInfo fields that act on the live path plus the cost weights and states the
// drivers read next to it.  out = [num_nodes, integrator_dt, friction_coef, force_bound, swing_height, foot_offset,
// ee_box_x, ee_box_y, force_cost, Q_srbd_diag[12], srb_init[13], srb_target[13]] (47 doubles).  0 on success.
extern "C" int bgg_host_parse_config(const char* path, double* out) {
    try {
        const utils::ConfigParser cfg(path);
        const mpc::MPCInfo info = mpc::MPCInfoFromConfig(cfg);
        int k = 0;
        out[k++] = info.num_nodes;
        out[k++] = info.integrator_dt;
        out[k++] = info.friction_coef;
        out[k++] = info.force_bound;
        out[k++] = info.swing_height;
        out[k++] = info.foot_offset;
        out[k++] = info.ee_box_size(0);
        out[k++] = info.ee_box_size(1);
        out[k++] = info.force_cost;
        const mpc::vector_t q = cfg.ParseEigenVector("Q_srbd_diag"), s0 = cfg.ParseEigenVector("srb_init");
        const mpc::vector_t s1 = cfg.Has("srb_target") ? cfg.ParseEigenVector("srb_target") : s0;   // a1_gait_opt_config.yaml has none
        if (q.size() != 12 || s0.size() != 13 || s1.size() != 13) throw std::runtime_error("unexpected vector sizes");
        for (int i = 0; i < 12; ++i) out[k++] = q(i);
        for (int i = 0; i < 13; ++i) out[k++] = s0(i);
        for (int i = 0; i < 13; ++i) out[k++] = s1(i);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "bgg_host_parse_config: %s\n", e.what());
        return -1;
    }
}
