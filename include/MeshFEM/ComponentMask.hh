// xyz component masks (mirrors ComponentMask.hh).
#ifndef MESHFEM_B200_COMPONENTMASK_HH
#define MESHFEM_B200_COMPONENTMASK_HH
#include <MeshFEM/Types.hh>

#include <bitset>

struct ComponentMask {
    ComponentMask(const std::string &components = "") { setComponentString(components); }
    void setComponentString(const std::string &components) {
        m_active.reset();
        if (components.find("x") != std::string::npos) m_active.set(0);
        if (components.find("y") != std::string::npos) m_active.set(1);
        if (components.find("z") != std::string::npos) m_active.set(2);
        if (m_active.count() != components.size()) throw std::runtime_error("invalid component specifier: '" + components + "'");
    }
    bool has(size_t c) const { return m_active.test(c); }
    bool hasX() const { return m_active[0]; }
    bool hasY() const { return m_active[1]; }
    bool hasZ() const { return m_active[2]; }
    bool hasAny(size_t dim) const { return count(dim) > 0; }
    bool hasAll(size_t dim) const { return count(dim) == dim; }
    size_t count(size_t dim) const {
        if (dim == 3) return m_active.count();
        if (dim == 2) return m_active.count() - (hasZ() ? 1 : 0);
        throw std::runtime_error("Illegal dimension");
    }
    void set() { m_active.set(); }
    void set(size_t c) { m_active.set(c); }
    void clear() { m_active.reset(); }
    void clear(size_t c) { m_active.reset(c); }
    bool operator==(const ComponentMask &b) const { return m_active == b.m_active; }
    bool operator!=(const ComponentMask &b) const { return m_active != b.m_active; }
    std::string componentString() const {
        std::string r;
        if (hasX()) r += "x";
        if (hasY()) r += "y";
        if (hasZ()) r += "z";
        return r;
    }

private:
    std::bitset<3> m_active;
};
#endif
