// IniParametersHandler.cpp -- see BipedalLocomotion/ParametersHandler.h
#include "BipedalLocomotion/ParametersHandler.h"

#include <cstdlib>
#include <fstream>
#include <sstream>

namespace BipedalLocomotion {
namespace ParametersHandler {

namespace {
std::string trim(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string stripComment(const std::string& line)
{
    bool quoted = false;
    for (size_t i = 0; i < line.size(); ++i) {
        if (line[i] == '"') quoted = !quoted;
        if (!quoted && (line[i] == '#' || (line[i] == '/' && i + 1 < line.size() && line[i + 1] == '/'))) return line.substr(0, i);
    }
    return line;
}
// tokens of a value: separators are blanks and commas; parentheses mark a list; quotes protect blanks
bool tokenize(const std::string& text, std::vector<std::string>& out, bool& isList)
{
    out.clear();
    isList = false;
    std::string cur;
    bool quoted = false, had = false;
    int depth = 0;
    for (char ch : text) {
        if (quoted) {
            if (ch == '"') { quoted = false; had = true; } else cur.push_back(ch);
            continue;
        }
        if (ch == '"') { quoted = true; continue; }
        if (ch == '(') { isList = true; ++depth; continue; }
        if (ch == ')') { --depth; if (depth < 0) return false; }
        if (ch == ' ' || ch == '\t' || ch == ',' || ch == ')') {
            if (!cur.empty() || had) out.push_back(cur);
            cur.clear(); had = false;
            continue;
        }
        cur.push_back(ch);
    }
    if (!cur.empty() || had) out.push_back(cur);
    return depth == 0 && !quoted;
}
std::string dirOf(const std::string& path)
{
    size_t p = path.find_last_of('/');
    return p == std::string::npos ? std::string(".") : path.substr(0, p);
}
}  // namespace

bool IniImplementation::setFromFile(const std::string& path)
{
    std::ifstream f(path);
    if (!f) { m_error = "cannot open " + path; return false; }
    std::stringstream ss;
    ss << f.rdbuf();
    return setFromString(ss.str(), dirOf(path));
}

bool IniImplementation::setFromString(const std::string& text, const std::string& baseDir)
{
    clear();
    std::istringstream in(text);
    std::string line;
    IniImplementation* cur = this;
    int lineNo = 0;
    while (std::getline(in, line)) {
        ++lineNo;
        line = trim(stripComment(line));
        if (line.empty()) continue;
        if (line.front() == '[') {
            if (line.back() != ']') { m_error = "line " + std::to_string(lineNo) + ": unterminated section"; return false; }
            std::vector<std::string> tok;
            bool dummy;
            tokenize(line.substr(1, line.size() - 2), tok, dummy);
            if (tok.empty()) { m_error = "line " + std::to_string(lineNo) + ": empty section"; return false; }
            if (tok[0] == "include") {
                // [include GROUP "./file.ini"]  or  [include "./file.ini"] (merged in place)
                if (tok.size() < 2) { m_error = "line " + std::to_string(lineNo) + ": include without file"; return false; }
                const std::string file = tok.back();
                const std::string full = (!file.empty() && file[0] == '/') ? file : baseDir + "/" + file;
                auto inc = std::make_shared<IniImplementation>();
                if (!inc->setFromFile(full)) { m_error = "line " + std::to_string(lineNo) + ": " + inc->m_error; return false; }
                if (tok.size() >= 3) { m_groups[tok[1]] = inc; }
                else { for (auto& kv : inc->m_values) m_values[kv.first] = kv.second; for (auto& kv : inc->m_groups) m_groups[kv.first] = kv.second; }
                cur = this;
                continue;
            }
            auto g = std::make_shared<IniImplementation>();
            m_groups[tok[0]] = g;
            cur = g.get();
            continue;
        }
        size_t sp = line.find_first_of(" \t");
        const std::string key = sp == std::string::npos ? line : line.substr(0, sp);
        const std::string rest = sp == std::string::npos ? std::string() : trim(line.substr(sp));
        Value v;
        if (!tokenize(rest, v.tokens, v.isList)) { m_error = "line " + std::to_string(lineNo) + ": unbalanced value"; return false; }
        cur->m_values[key] = v;
    }
    return true;
}

const IniImplementation::Value* IniImplementation::find(const std::string& name) const
{
    auto it = m_values.find(name);
    return it == m_values.end() ? nullptr : &it->second;
}

static bool toDouble(const std::string& s, double& v)
{
    if (s.empty()) return false;
    char* end = nullptr;
    v = std::strtod(s.c_str(), &end);
    return end && *end == '\0';
}

bool IniImplementation::getParameter(const std::string& name, double& v) const
{
    const Value* val = find(name);
    return val && val->tokens.size() == 1 && toDouble(val->tokens[0], v);
}
bool IniImplementation::getParameter(const std::string& name, int& v) const
{
    double d;
    if (!getParameter(name, d) || d != (double)(long long)d) return false;
    v = (int)d;
    return true;
}
bool IniImplementation::getParameter(const std::string& name, bool& v) const
{
    const Value* val = find(name);
    if (!val || val->tokens.size() != 1) return false;
    const std::string& s = val->tokens[0];
    if (s == "true" || s == "True" || s == "1") { v = true; return true; }
    if (s == "false" || s == "False" || s == "0") { v = false; return true; }
    return false;
}
bool IniImplementation::getParameter(const std::string& name, std::string& v) const
{
    const Value* val = find(name);
    if (!val || val->tokens.size() != 1) return false;
    v = val->tokens[0];
    return true;
}
bool IniImplementation::getParameter(const std::string& name, std::vector<double>& v) const
{
    const Value* val = find(name);
    if (!val) return false;
    std::vector<double> out;
    for (const auto& t : val->tokens) { double d; if (!toDouble(t, d)) return false; out.push_back(d); }
    v = out;
    return true;
}
bool IniImplementation::getParameter(const std::string& name, std::vector<int>& v) const
{
    std::vector<double> d;
    if (!getParameter(name, d)) return false;
    v.clear();
    for (double x : d) { if (x != (double)(long long)x) return false; v.push_back((int)x); }
    return true;
}
bool IniImplementation::getParameter(const std::string& name, std::vector<std::string>& v) const
{
    const Value* val = find(name);
    if (!val) return false;
    v = val->tokens;
    return true;
}

static std::string num(double d) { std::ostringstream o; o.precision(17); o << d; return o.str(); }
void IniImplementation::setParameter(const std::string& name, int v) { m_values[name] = Value{{std::to_string(v)}, false}; }
void IniImplementation::setParameter(const std::string& name, double v) { m_values[name] = Value{{num(v)}, false}; }
void IniImplementation::setParameter(const std::string& name, bool v) { m_values[name] = Value{{v ? "true" : "false"}, false}; }
void IniImplementation::setParameter(const std::string& name, const std::string& v) { m_values[name] = Value{{v}, false}; }
void IniImplementation::setParameter(const std::string& name, const std::vector<double>& v)
{
    Value val;
    val.isList = true;
    for (double d : v) val.tokens.push_back(num(d));
    m_values[name] = val;
}

IParametersHandler::weak_ptr IniImplementation::getGroup(const std::string& name) const
{
    auto it = m_groups.find(name);
    if (it == m_groups.end()) return weak_ptr();
    return std::static_pointer_cast<IParametersHandler>(it->second);
}
bool IniImplementation::setGroup(const std::string& name, shared_ptr group)
{
    auto g = std::dynamic_pointer_cast<IniImplementation>(group);
    if (!g) return false;
    m_groups[name] = g;
    return true;
}
IParametersHandler::shared_ptr IniImplementation::clone() const
{
    auto c = std::make_shared<IniImplementation>();
    c->m_values = m_values;
    for (const auto& kv : m_groups) c->m_groups[kv.first] = std::static_pointer_cast<IniImplementation>(kv.second->clone());
    return c;
}
std::string IniImplementation::toString() const
{
    std::ostringstream o;
    for (const auto& kv : m_values) {
        o << kv.first << " ";
        if (kv.second.isList) o << "(";
        for (size_t i = 0; i < kv.second.tokens.size(); ++i) o << (i ? ", " : "") << kv.second.tokens[i];
        if (kv.second.isList) o << ")";
        o << "\n";
    }
    for (const auto& kv : m_groups) o << "[" << kv.first << "]\n" << kv.second->toString();
    return o.str();
}

}  // namespace ParametersHandler
}  // namespace BipedalLocomotion
