// cellbp_script.h -- the cellBP script format (reference: cellBP/newparser.cpp:135-232, lexer.cpp:76-361,
// scriptinit.cpp:307-604): `name value`, `name ( ... )` branches, `[v v v]` vectors, `#` comments to end of line.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace cellbp {

struct ScriptTree {
    std::map<std::string, double> numbers;
    std::map<std::string, std::string> strings;
    std::map<std::string, std::vector<double>> vectors;
    std::map<std::string, ScriptTree> branches;
    bool get_number(const std::string &k, double &out) const;
    bool get_int(const std::string &k, int &out) const;
    bool get_string(const std::string &k, std::string &out) const;
    const ScriptTree *get_branch(const std::string &k) const;
};

// throws std::runtime_error on malformed input
ScriptTree parse_script_text(const std::string &text);
ScriptTree parse_script_file(const std::string &path);

// what scriptinit.cpp extracts for the cell / substrate runs (scriptinit.cpp:307-372, :535-580, :40-75)
struct ScriptInfo {
    double frame_dt = 0, sim_dt = 0, end_sim_t = 1e30, curr_t = 0;
    bool curr_t_specified = false;
    std::string output_path = "./";
    bool output_path_is_relative = false;
    std::string mesh_path;                 // ecomsolfile ( filepath ... )
    int perform_improvement = 1, allow_topology_changes = 1, collision_safety = 1;   // surftrack_parameters
};
ScriptInfo read_script(const std::string &path);

}   // namespace cellbp
