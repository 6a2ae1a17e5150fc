// Expression-valued boundary conditions ("sin(pi*x)"): a small recursive-descent evaluator
// with tinyexpr's grammar and function set (the reference wraps tinyexpr,
// ExpressionVector.hh:21-58; tinyexpr itself is not vendored): + - * / % ^, unary sign,
// parentheses, constants pi/e, functions abs acos asin atan atan2 ceil cos cosh exp fac floor
// ln log log10 ncr npr pow sin sinh sqrt tan tanh.  As in tinyexpr's default build, `^`
// associates left-to-right, unary minus binds tighter than `^`, and `log` is base 10.
#ifndef MESHFEM_B200_EXPRESSIONVECTOR_HH
#define MESHFEM_B200_EXPRESSIONVECTOR_HH
#include <MeshFEM/Types.hh>

#include <cctype>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

struct ExpressionEnvironment {
    std::map<std::string, Real> vars;
    void setValue(const std::string &name, Real v) { vars[name] = v; }
    template <class Vec>
    void setVectorValue(const std::string &prefix, const Vec &v) {   // mesh_size_0, mesh_size_1, ...
        for (size_t i = 0; i < Vec::size(); ++i) vars[prefix + std::to_string(i)] = v[i];
    }
    template <class Vec>
    void setXYZ(const Vec &p) {
        vars["x"] = p[0];
        vars["y"] = Vec::size() > 1 ? p[1] : 0.0;
        vars["z"] = Vec::size() > 2 ? p[2] : 0.0;
    }
};

class Expression {
public:
    Expression() {}
    explicit Expression(const std::string &e) : m_expr(e) {}
    const std::string &string() const { return m_expr; }
    Real eval(const ExpressionEnvironment &env) const {
        Parser p{m_expr, 0, env};
        const Real v = p.list();
        p.skip();
        if (p.pos != m_expr.size()) throw std::runtime_error("Error parsing expression '" + m_expr + "'");
        return v;
    }

private:
    std::string m_expr;
    struct Parser {
        const std::string &s;
        size_t pos;
        const ExpressionEnvironment &env;
        void skip() { while (pos < s.size() && std::isspace((unsigned char)s[pos])) ++pos; }
        bool eat(char c) { skip(); if (pos < s.size() && s[pos] == c) { ++pos; return true; } return false; }
        [[noreturn]] void fail() const { throw std::runtime_error("Error parsing expression '" + s + "' near offset " + std::to_string(pos)); }
        Real list() { Real v = expr(); while (eat(',')) v = expr(); return v; }
        Real expr() {
            Real v = term();
            while (true) {
                if (eat('+')) v += term();
                else if (eat('-')) v -= term();
                else return v;
            }
        }
        Real term() {
            Real v = factor();
            while (true) {
                if (eat('*')) v *= factor();
                else if (eat('/')) v /= factor();
                else if (eat('%')) v = std::fmod(v, factor());
                else return v;
            }
        }
        Real factor() {
            Real v = power();
            while (eat('^')) v = std::pow(v, power());
            return v;
        }
        Real power() {
            int sign = 1;
            while (true) {
                if (eat('-')) sign = -sign;
                else if (eat('+')) {}
                else break;
            }
            return sign * base();
        }
        static Real fac(Real a) {
            if (a < 0.0) return NAN;
            unsigned long result = 1;
            for (unsigned long i = 1; i <= (unsigned long)a; ++i) result *= i;
            return (Real)result;
        }
        static Real ncr(Real n, Real r) {
            if (n < 0.0 || r < 0.0 || n < r) return NAN;
            unsigned long un = (unsigned long)n, ur = (unsigned long)r, result = 1;
            if (ur > un / 2) ur = un - ur;
            for (unsigned long i = 1; i <= ur; ++i) { result *= un - ur + i; result /= i; }
            return (Real)result;
        }
        Real base() {
            skip();
            if (pos >= s.size()) fail();
            const char c = s[pos];
            if (std::isdigit((unsigned char)c) || c == '.') {
                char *end = nullptr;
                const Real v = std::strtod(s.c_str() + pos, &end);
                if (end == s.c_str() + pos) fail();
                pos += size_t(end - (s.c_str() + pos));
                return v;
            }
            if (c == '(') {
                ++pos;
                const Real v = list();
                if (!eat(')')) fail();
                return v;
            }
            if (std::isalpha((unsigned char)c) || c == '_') {
                size_t b = pos;
                while (pos < s.size() && (std::isalnum((unsigned char)s[pos]) || s[pos] == '_')) ++pos;
                const std::string name = s.substr(b, pos - b);
                auto it = env.vars.find(name);
                if (it != env.vars.end()) return it->second;
                if (name == "pi") return 3.14159265358979323846;
                if (name == "e") return 2.71828182845904523536;
                // two-argument functions need a parenthesised list
                if (name == "atan2" || name == "pow" || name == "ncr" || name == "npr") {
                    if (!eat('(')) fail();
                    const Real a = expr();
                    if (!eat(',')) fail();
                    const Real b2 = expr();
                    if (!eat(')')) fail();
                    if (name == "atan2") return std::atan2(a, b2);
                    if (name == "pow") return std::pow(a, b2);
                    if (name == "ncr") return ncr(a, b2);
                    return ncr(a, b2) * fac(b2);
                }
                typedef Real (*F1)(Real);
                static const std::map<std::string, F1> f1 = {
                    {"abs", [](Real a) { return std::fabs(a); }}, {"acos", [](Real a) { return std::acos(a); }},
                    {"asin", [](Real a) { return std::asin(a); }}, {"atan", [](Real a) { return std::atan(a); }},
                    {"ceil", [](Real a) { return std::ceil(a); }}, {"cos", [](Real a) { return std::cos(a); }},
                    {"cosh", [](Real a) { return std::cosh(a); }}, {"exp", [](Real a) { return std::exp(a); }},
                    {"fac", [](Real a) { return fac(a); }}, {"floor", [](Real a) { return std::floor(a); }},
                    {"ln", [](Real a) { return std::log(a); }}, {"log", [](Real a) { return std::log10(a); }},
                    {"log10", [](Real a) { return std::log10(a); }}, {"sin", [](Real a) { return std::sin(a); }},
                    {"sinh", [](Real a) { return std::sinh(a); }}, {"sqrt", [](Real a) { return std::sqrt(a); }},
                    {"tan", [](Real a) { return std::tan(a); }}, {"tanh", [](Real a) { return std::tanh(a); }}};
                auto fit = f1.find(name);
                if (fit == f1.end()) fail();
                return fit->second(power());     // tinyexpr: <function-1> <power>
            }
            fail();
        }
    };
};

class ExpressionVector {
public:
    void add(const std::string &expr) { m_exprs.emplace_back(expr); }
    size_t size() const { return m_exprs.size(); }
    const std::string &componentString(size_t i) const { return m_exprs.at(i).string(); }
    template <size_t N>
    VectorND<N> eval(const ExpressionEnvironment &env) const {
        if (m_exprs.size() != N) throw std::runtime_error("Invalid expression vector size.");
        VectorND<N> r;
        for (size_t i = 0; i < N; ++i) r[i] = m_exprs[i].eval(env);
        return r;
    }

private:
    std::vector<Expression> m_exprs;
};
#endif
