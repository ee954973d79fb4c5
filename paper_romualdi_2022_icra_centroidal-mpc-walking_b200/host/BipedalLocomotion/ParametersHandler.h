// ParametersHandler.h -- BLF ParametersHandler::IParametersHandler as used by the reference to configure the MPC
// (src/centroidal-mpc-walking/src/Main.cpp:53-59, 91; src/CentroidalMPCBlock.cpp:112-161: getGroup / getParameter / lock()),
// plus a self-contained implementation over YARP .ini files (the reference uses YarpImplementation over
// yarp::os::ResourceFinder; YARP is not available here).  Format notes in SURVEY.md 5.6.
#pragma once

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace BipedalLocomotion {
namespace ParametersHandler {

class IParametersHandler {
public:
    using shared_ptr = std::shared_ptr<IParametersHandler>;
    using weak_ptr = std::weak_ptr<IParametersHandler>;
    virtual ~IParametersHandler() = default;

    virtual bool getParameter(const std::string& name, int& v) const = 0;
    virtual bool getParameter(const std::string& name, double& v) const = 0;
    virtual bool getParameter(const std::string& name, bool& v) const = 0;
    virtual bool getParameter(const std::string& name, std::string& v) const = 0;
    virtual bool getParameter(const std::string& name, std::vector<double>& v) const = 0;
    virtual bool getParameter(const std::string& name, std::vector<int>& v) const = 0;
    virtual bool getParameter(const std::string& name, std::vector<std::string>& v) const = 0;

    virtual void setParameter(const std::string& name, int v) = 0;
    virtual void setParameter(const std::string& name, double v) = 0;
    virtual void setParameter(const std::string& name, bool v) = 0;
    virtual void setParameter(const std::string& name, const std::string& v) = 0;
    virtual void setParameter(const std::string& name, const std::vector<double>& v) = 0;

    // empty weak_ptr when the group does not exist (BLF semantics)
    virtual weak_ptr getGroup(const std::string& name) const = 0;
    virtual bool setGroup(const std::string& name, shared_ptr group) = 0;
    virtual bool isEmpty() const = 0;
    virtual void clear() = 0;
    virtual shared_ptr clone() const = 0;
    virtual std::string toString() const = 0;
};

// YARP-ini backed handler.  Grammar: `key value...`, lists in parentheses with comma OR whitespace separators (the
// reference files contain `corner_3 (-0.08 0.03, 0.0)`), optional quotes, `#` / `//` comments, `[GROUP]` sections and
// `[include GROUP "./file.ini"]` (quotes optional), which nests the included file's top level under GROUP.
class IniImplementation : public IParametersHandler {
public:
    IniImplementation() = default;
    // false (and lastError()) when the file cannot be read or parsed
    bool setFromFile(const std::string& path);
    bool setFromString(const std::string& text, const std::string& baseDir = ".");
    const std::string& lastError() const { return m_error; }

    bool getParameter(const std::string& name, int& v) const override;
    bool getParameter(const std::string& name, double& v) const override;
    bool getParameter(const std::string& name, bool& v) const override;
    bool getParameter(const std::string& name, std::string& v) const override;
    bool getParameter(const std::string& name, std::vector<double>& v) const override;
    bool getParameter(const std::string& name, std::vector<int>& v) const override;
    bool getParameter(const std::string& name, std::vector<std::string>& v) const override;
    void setParameter(const std::string& name, int v) override;
    void setParameter(const std::string& name, double v) override;
    void setParameter(const std::string& name, bool v) override;
    void setParameter(const std::string& name, const std::string& v) override;
    void setParameter(const std::string& name, const std::vector<double>& v) override;
    weak_ptr getGroup(const std::string& name) const override;
    bool setGroup(const std::string& name, shared_ptr group) override;
    bool isEmpty() const override { return m_values.empty() && m_groups.empty(); }
    void clear() override { m_values.clear(); m_groups.clear(); }
    shared_ptr clone() const override;
    std::string toString() const override;

private:
    struct Value { std::vector<std::string> tokens; bool isList = false; };
    std::map<std::string, Value> m_values;
    std::map<std::string, std::shared_ptr<IniImplementation>> m_groups;
    std::string m_error;
    const Value* find(const std::string& name) const;
};

}  // namespace ParametersHandler
}  // namespace BipedalLocomotion
