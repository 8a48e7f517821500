#include "cellbp_script.h"
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace cellbp {

namespace {
enum Kind { T_END, T_NAME, T_NUMBER, T_LPAREN, T_RPAREN, T_LBRACKET, T_RBRACKET };
struct Token { Kind kind; std::string text; double value; };

struct Lexer {
    const std::string &s; size_t i = 0;
    explicit Lexer(const std::string &text) : s(text) {}
    Token next() {
        for (;;) {
            while (i < s.size() && std::isspace((unsigned char)s[i])) i++;
            if (i < s.size() && s[i] == '#') { while (i < s.size() && s[i] != '\n') i++; continue; }
            break;
        }
        if (i >= s.size()) return {T_END, "", 0};
        const char ch = s[i];
        if (ch == '(') { i++; return {T_LPAREN, "(", 0}; }
        if (ch == ')') { i++; return {T_RPAREN, ")", 0}; }
        if (ch == '[') { i++; return {T_LBRACKET, "[", 0}; }
        if (ch == ']') { i++; return {T_RBRACKET, "]", 0}; }
        const size_t b = i;
        while (i < s.size() && !std::isspace((unsigned char)s[i]) && s[i] != '(' && s[i] != ')' && s[i] != '[' && s[i] != ']' && s[i] != '#') i++;
        const std::string w = s.substr(b, i - b);
        char *end = nullptr;
        const double v = std::strtod(w.c_str(), &end);
        if (end && *end == '\0' && end != w.c_str()) return {T_NUMBER, w, v};
        return {T_NAME, w, 0};
    }
};

void parse_block(Lexer &lx, ScriptTree &tree, bool top) {
    for (;;) {
        Token t = lx.next();
        if (t.kind == T_END) { if (!top) throw std::runtime_error("cellBP script: missing ')'"); return; }
        if (t.kind == T_RPAREN) { if (top) throw std::runtime_error("cellBP script: unmatched ')'"); return; }
        if (t.kind != T_NAME) throw std::runtime_error("cellBP script: expected a name, got '" + t.text + "'");
        const std::string name = t.text;
        Token v = lx.next();
        switch (v.kind) {
        case T_NUMBER: tree.numbers[name] = v.value; break;
        case T_NAME: tree.strings[name] = v.text; break;
        case T_LPAREN: parse_block(lx, tree.branches[name], false); break;
        case T_LBRACKET: {
            std::vector<double> vec;
            for (;;) {
                Token e = lx.next();
                if (e.kind == T_RBRACKET) break;
                if (e.kind != T_NUMBER) throw std::runtime_error("cellBP script: non-number inside [ ] of '" + name + "'");
                vec.push_back(e.value);
            }
            tree.vectors[name] = vec;
            break;
        }
        default: throw std::runtime_error("cellBP script: '" + name + "' has no value");
        }
    }
}
}   // namespace

bool ScriptTree::get_number(const std::string &k, double &out) const { auto it = numbers.find(k); if (it == numbers.end()) return false; out = it->second; return true; }
bool ScriptTree::get_int(const std::string &k, int &out) const { double d; if (!get_number(k, d)) return false; out = (int)d; return true; }
bool ScriptTree::get_string(const std::string &k, std::string &out) const {
    auto it = strings.find(k);
    if (it != strings.end()) { out = it->second; return true; }
    auto in = numbers.find(k);                       // a path made of digits lexes as a number
    if (in != numbers.end()) { std::ostringstream o; o << in->second; out = o.str(); return true; }
    return false;
}
const ScriptTree *ScriptTree::get_branch(const std::string &k) const { auto it = branches.find(k); return it == branches.end() ? nullptr : &it->second; }

ScriptTree parse_script_text(const std::string &text) { Lexer lx(text); ScriptTree t; parse_block(lx, t, true); return t; }
ScriptTree parse_script_file(const std::string &path) {
    std::ifstream f(path);
    if (!f.good()) throw std::runtime_error("Could not open script file " + path);      // scriptinit.cpp:311-315
    std::stringstream ss; ss << f.rdbuf();
    return parse_script_text(ss.str());
}

ScriptInfo read_script(const std::string &path) {
    const ScriptTree tree = parse_script_file(path);
    ScriptInfo s;
    if (!tree.get_number("frame_dt", s.frame_dt)) throw std::runtime_error(path + ": frame_dt missing");   // scriptinit.cpp:328
    int nsub = 0;
    const bool sub = tree.get_int("num_substeps", nsub);
    if (sub && nsub > 0) s.sim_dt = s.frame_dt / (double)nsub;
    double sd;
    if (tree.get_number("sim_dt", sd)) s.sim_dt = sd;                                     // sim_dt wins (scriptinit.cpp:339-347)
    if (!(s.sim_dt > 0)) throw std::runtime_error(path + ": neither sim_dt nor num_substeps given");
    tree.get_number("end_sim_t", s.end_sim_t);
    s.curr_t_specified = tree.get_number("curr_t", s.curr_t);
    if (tree.get_string("output_path", s.output_path)) s.output_path_is_relative = false;
    else if (tree.get_string("relative_output_path", s.output_path)) s.output_path_is_relative = true;
    else { s.output_path = "./"; s.output_path_is_relative = false; }
    if (const ScriptTree *b = tree.get_branch("surftrack_parameters")) {
        b->get_int("perform_improvement", s.perform_improvement);
        b->get_int("allow_topology_changes", s.allow_topology_changes);
        b->get_int("collision_safety", s.collision_safety);
    }
    if (const ScriptTree *b = tree.get_branch("ecomsolfile")) b->get_string("filepath", s.mesh_path);
    return s;
}

}   // namespace cellbp
