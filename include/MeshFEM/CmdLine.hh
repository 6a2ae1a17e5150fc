// Minimal command-line parser for the CLIs under src/bin (the reference uses
// boost::program_options, which is not in this image).  Supports what those CLIs need:
// "--long value", "--long=value", "-s value", "-svalue", flags without a value, and positionals;
// count()/as-style access mirrors po::variables_map so the CLI bodies read like the reference's.
#ifndef MESHFEM_B200_CMDLINE_HH
#define MESHFEM_B200_CMDLINE_HH
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

class CmdLine {
public:
    struct Option {
        std::string longName;
        char shortName;       // 0 = none
        bool takesValue;
        std::string defaultValue;
        bool hasDefault;
        std::string help;
    };

    CmdLine &flag(const std::string &longName, char shortName, const std::string &help) {
        m_opts.push_back({longName, shortName, false, "", false, help});
        return *this;
    }
    CmdLine &value(const std::string &longName, char shortName, const std::string &help) {
        m_opts.push_back({longName, shortName, true, "", false, help});
        return *this;
    }
    CmdLine &value(const std::string &longName, char shortName, const std::string &help, const std::string &def) {
        m_opts.push_back({longName, shortName, true, def, true, help});
        return *this;
    }
    CmdLine &positional(const std::string &name) { m_positionalNames.push_back(name); return *this; }

    // throws std::runtime_error on unknown options / missing values
    void parse(int argc, const char *argv[]) {
        for (const auto &o : m_opts) if (o.hasDefault) m_values[o.longName] = o.defaultValue;
        size_t nextPositional = 0;
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            const Option *opt = nullptr;
            std::string inlineValue;
            bool haveInline = false;
            if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
                const size_t eq = a.find('=');
                const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
                opt = m_find(name);
                if (!opt) throw std::runtime_error("unrecognised option '" + a + "'");
                if (eq != std::string::npos) { inlineValue = a.substr(eq + 1); haveInline = true; }
            } else if (a.size() >= 2 && a[0] == '-' && !m_looksNumeric(a)) {
                opt = m_find(a[1]);
                if (!opt) throw std::runtime_error("unrecognised option '" + a + "'");
                if (a.size() > 2) { inlineValue = a.substr(2); haveInline = true; }
            } else {
                if (nextPositional >= m_positionalNames.size()) throw std::runtime_error("too many positional options have been specified on the command line");
                m_values[m_positionalNames[nextPositional++]] = a;
                continue;
            }
            if (opt->takesValue) {
                if (!haveInline) {
                    if (i + 1 >= argc) throw std::runtime_error("the required argument for option '--" + opt->longName + "' is missing");
                    inlineValue = argv[++i];
                }
                m_values[opt->longName] = inlineValue;
            } else {
                if (haveInline) throw std::runtime_error("option '--" + opt->longName + "' does not take any arguments");
                m_values[opt->longName] = "";
            }
        }
    }

    size_t count(const std::string &name) const { return m_values.count(name); }
    const std::string &str(const std::string &name) const {
        auto it = m_values.find(name);
        if (it == m_values.end()) throw std::runtime_error("option '" + name + "' not set");
        return it->second;
    }
    int integer(const std::string &name) const {
        size_t pos = 0;
        const std::string &s = str(name);
        int v = 0;
        try { v = std::stoi(s, &pos); } catch (...) { pos = 0; }
        if (pos != s.size() || s.empty()) throw std::runtime_error("the argument ('" + s + "') for option '--" + name + "' is invalid");
        return v;
    }

    void printOptions(std::ostream &os) const {
        for (const auto &o : m_opts) {
            std::ostringstream l;
            l << "  ";
            if (o.shortName) l << '-' << o.shortName << " [ --" << o.longName << " ]"; else l << "--" << o.longName;
            if (o.takesValue) { l << " arg"; if (o.hasDefault && !o.defaultValue.empty()) l << " (=" << o.defaultValue << ")"; }
            os << std::left << std::setw(40) << l.str() << ' ' << o.help << std::endl;
        }
    }

private:
    std::vector<Option> m_opts;
    std::vector<std::string> m_positionalNames;
    std::map<std::string, std::string> m_values;
    const Option *m_find(const std::string &longName) const {
        for (const auto &o : m_opts) if (o.longName == longName) return &o;
        return nullptr;
    }
    const Option *m_find(char s) const {
        for (const auto &o : m_opts) if (o.shortName == s) return &o;
        return nullptr;
    }
    static bool m_looksNumeric(const std::string &a) {   // "-1,0,0" style values are positionals, not options
        return a.size() >= 2 && (std::isdigit((unsigned char)a[1]) || a[1] == '.');
    }
};

inline std::string fileExtension(const std::string &path) {
    const size_t dot = path.rfind('.');
    const size_t slash = path.find_last_of("/\\");
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return "";
    return path.substr(dot);
}
#endif
