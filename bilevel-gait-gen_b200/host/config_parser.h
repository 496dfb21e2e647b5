// bilevel-gait-gen_b200 -- utils::ConfigParser of the reference (utils/include/config_parser.h:13-40) without yaml-cpp:
// the reference's configuration files (apps/*.yaml) only use `key: scalar`, `key: "string"` and flow sequences
// `key: [a, b, ...]` that may continue over several lines, with `#` comments.  Same member names, so
// test/mpc_test.cpp:43-89 / test/simulation_mpc.cpp:55-89 style code fills an MPCInfo unchanged.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "mpc_b200.h"

namespace utils {

class ConfigParser {
public:
    explicit ConfigParser(const std::string& file_name);

    mpc::vector_t ParseEigenVector(const std::string& element) const;
    std::string ParseString(const std::string& element) const;
    std::vector<std::string> ParseStringVector(const std::string& element) const;

    template <typename scalar>
    scalar ParseNumber(const std::string& element) const {
        return static_cast<scalar>(Number(element));
    }
    template <typename scalar>
    std::vector<scalar> ParseStdVector(const std::string& element) const {
        std::vector<scalar> out;
        Fill(element, out);
        return out;
    }
    bool Has(const std::string& element) const { return items_.count(element) != 0; }

private:
    const std::vector<std::string>& Items(const std::string& element) const;
    double Number(const std::string& element) const;
    void Fill(const std::string& element, std::vector<double>& out) const;
    void Fill(const std::string& element, std::vector<int>& out) const;
    void Fill(const std::string& element, std::vector<std::string>& out) const;
    std::string file_name_;
    std::map<std::string, std::vector<std::string>> items_;   // scalars are one-element lists
};

}  // namespace utils

namespace mpc {
// The MPCInfo every driver of the reference builds from its configuration file (test/mpc_test.cpp:46-83,
// test/simulation_mpc.cpp:55-89, apps/mpc_demo.cpp).
MPCInfo MPCInfoFromConfig(const utils::ConfigParser& config);
}  // namespace mpc
